"""Host-side scalar vocabulary shared by every layer of skrample_b200.

Everything here is float64 Python/NumPy arithmetic that never touches a
device tensor: schedule coordinates (``Point``/``DeltaPoint``/``Step``), list
merge policy, and the handful of numeric helpers the samplers use to derive
per-step coefficients.  The CUDA path only ever sees the *results* of these
helpers, as scalars baked into a step program.

Mirrors the public surface of the reference ``skrample/common.py``
(reference: skrample/common.py:11-213) so user code written against it keeps
working; the expression order of every helper is kept because the schedule /
coefficient values must be bit-identical to the reference.
"""

from __future__ import annotations

import enum
import math
from collections.abc import Callable
from functools import lru_cache
from typing import TYPE_CHECKING, Any, NamedTuple

import numpy as np
from numpy.typing import NDArray

if TYPE_CHECKING:
    from torch import Tensor

    type Sample = float | NDArray[np.floating] | Tensor
else:  # torch stays an optional import for the scalar layer
    type Sample = float | NDArray[np.floating]

type RNG[T: Sample] = Callable[["Step | None"], T]
"Noise source keyed by the step being taken (``None`` = initial latent)."


class Point(NamedTuple):
    "One coordinate on a noise schedule. reference: skrample/common.py:24-40"

    timestep: float
    sigma: float
    alpha: float

    def add_noise[T: Sample](self, sample: T, noise: T) -> T:
        "``sample*alpha + noise*sigma`` (device tensors run the fused axpby kernel)."
        from skrample_b200.sampling import program  # late: keeps the scalar layer torch-free

        return program.point_add_noise(self, sample, noise)

    def remove_noise[T: Sample](self, sample: T, noise: T) -> T:
        "``(sample - noise*sigma) / alpha``; a scalar alpha of 0 returns the scaled noise."
        from skrample_b200.sampling import program

        return program.point_remove_noise(self, sample, noise)


class DeltaPoint(NamedTuple):
    "Schedule coordinates at both ends of a step. reference: skrample/common.py:43-52"

    point_from: Point
    point_to: Point

    def difference(self) -> Point:
        a, b = self
        return Point(b.timestep - a.timestep, b.sigma - a.sigma, b.alpha - a.alpha)


class Step(NamedTuple):
    """A sampling step as a pair of normalised times in 0..=1.

    reference: skrample/common.py:55-97
    """

    time_from: float
    time_to: float

    @staticmethod
    def from_int(position: int, amount: int) -> "Step":
        return Step(position / amount, (position + 1) / amount)

    def distance(self) -> float:
        return self.time_to - self.time_from

    def offset(self, steps: int | float) -> "Step":
        shift = self.distance() * steps
        return Step(self.time_from + shift, self.time_to + shift)

    def clamp(self) -> "Step":
        width = self.distance()
        return Step(clamp(self.time_from, high=1 - width), clamp(self.time_to, low=width))

    def position(self) -> float:
        return self.time_from / self.distance()

    def amount(self) -> float:
        return 1 / self.distance()

    def normal(self) -> "Step":
        return Step(min(self), max(self))


@enum.unique
class MergeStrategy(enum.StrEnum):
    "How two option lists combine. reference: skrample/common.py:100-130"

    Ours = enum.auto()
    Theirs = enum.auto()
    After = enum.auto()
    Before = enum.auto()
    UniqueAfter = enum.auto()
    UniqueBefore = enum.auto()

    def merge[T](self, ours: list[T], theirs: list[T], cmp: Callable[[T, T], bool] = lambda a, b: a == b) -> list[T]:
        def absent_from(pool: list[T], item: T) -> bool:
            return not any(cmp(member, item) for member in pool)

        if self is MergeStrategy.Ours:
            return ours
        if self is MergeStrategy.Theirs:
            return theirs
        if self is MergeStrategy.After:
            return ours + theirs
        if self is MergeStrategy.Before:
            return theirs + ours
        if self is MergeStrategy.UniqueAfter:
            return ours + [item for item in theirs if absent_from(ours, item)]
        return theirs + [item for item in ours if absent_from(theirs, item)]


def divf(lhs: float, rhs: float) -> float:
    "Division that saturates to ±inf; 0/0 raises. reference: skrample/common.py:133-140"
    if rhs != 0:
        return lhs / rhs
    if lhs == 0:
        raise ZeroDivisionError
    return math.copysign(math.inf, lhs)


def ln(x: float) -> float:
    "Natural log with ln(0) = -inf. reference: skrample/common.py:143-150"
    if x > 0:
        return math.log(x)
    if x < 0:
        raise ValueError
    return -math.inf


def normalize[T: Sample](regular: T, start: float, end: float = 0) -> T:
    "start..end -> 1..0"
    return (regular - end) / (start - end)  # type: ignore[return-value]


def regularize[T: Sample](normal: T, start: float, end: float = 0) -> T:
    "1..0 -> start..end"
    return normal * (start - end) + end  # type: ignore[return-value]


def rescale_positive(x: float) -> float:
    "Monotone map of the real line onto (0, inf) with 0 -> 1: (|x| + 1) ** sign(x)."
    return (abs(x) + 1) ** math.copysign(1, x)


def rescale_subnormal(x: float) -> float:
    "Monotone map of the real line onto (-1, 1) with 0 -> 0."
    return math.copysign(1 - (abs(x) + 1) ** -1, x)


def exp[T: Sample](x: T) -> T:
    return math.e**x  # type: ignore[return-value]


def sigmoid[T: Sample](array: T) -> T:
    grown: Any = exp(array)
    return grown / (1 + grown)


def softmax[T: tuple[Sample, ...]](elems: T) -> T:
    grown = [exp(e) for e in elems]
    total = sum(grown)
    return tuple(g / total for g in grown)  # type: ignore[return-value]


def spowf[T: Sample](x: T, f: float) -> T:
    "Sign-preserving power ``|x|**f * sign(x)``. reference: skrample/common.py:187-190"
    return abs(x) ** f * (-1 * (x < 0) | 1)  # type: ignore[operator,return-value]


def mean(x: Sample) -> float:
    return x if isinstance(x, float | int) else x.mean().item()  # type: ignore[union-attr]


def clamp(x: float, low: float = 0, high: float = 1) -> float:
    return max(low, min(high, x))


@lru_cache
def bashforth(order: int) -> tuple[float, ...]:
    """Adams-Bashforth weights: solve ``sum_j (-j)^k b_j = 1/(k+1)``.

    reference: skrample/common.py:205-213 (same linear system, same solver,
    so the float64 weights are identical).
    """
    vandermonde = [[(-j) ** k for j in range(order)] for k in range(order)]
    moments = [1 / (k + 1) for k in range(order)]
    return tuple(np.linalg.solve(vandermonde, moments).tolist())
