"""Noise schedules: float64 host scalars, never device tensors.

The samplers only consume ``(timestep, sigma, alpha)`` triples; those are
produced here with NumPy float64 and baked into step programs before a kernel
launch.  The north star requires sigma/timestep values bit-identical to the
reference, so each curve keeps the reference's floating-point expression order
(reference: skrample/scheduling.py:22-664); what differs is how the classes are
put together (shared helpers for the "endcap" normalisation trick, a single
place for timestep mapping, plan-friendly caches).
"""

from __future__ import annotations

import functools
import math
from abc import ABC, abstractmethod
from collections.abc import Sequence
from dataclasses import dataclass, replace
from typing import Literal, Self

import numpy as np

from .common import DeltaPoint, Point, Step, normalize, regularize, rescale_positive, sigmoid

type NPPoints = np.ndarray[tuple[int, Literal[3]], np.dtype[np.float64]]
"rows of (timestep, sigma, alpha)"
type NPSequence = np.ndarray[tuple[int], np.dtype[np.float64]]
type Sigma = NPSequence | float


# --------------------------------------------------------------------------- sigma spaces


@dataclass(frozen=True)
class SigmaSpace(ABC):
    "Mapping between 'regular' sigmas (0..inf) and the normalised (sigma, alpha) pair."

    @abstractmethod
    def normalize(self, regular_sigmas: Sigma) -> tuple[NPSequence, NPSequence]: ...

    @abstractmethod
    def regularize(self, normal_sigmas: Sigma) -> NPSequence: ...


@dataclass(frozen=True)
class VariancePreserving(SigmaSpace):
    "sigma^2 + alpha^2 = 1. reference: skrample/scheduling.py:31-38"

    def normalize(self, regular_sigmas: Sigma) -> tuple[NPSequence, NPSequence]:
        angle = np.atan(regular_sigmas)
        return np.sin(angle), np.cos(angle)

    def regularize(self, normal_sigmas: Sigma) -> NPSequence:
        return np.tan(np.asin(normal_sigmas))


@dataclass(frozen=True)
class FlowMatching(SigmaSpace):
    "sigma + alpha = 1. reference: skrample/scheduling.py:41-48"

    def normalize(self, regular_sigmas: Sigma) -> tuple[NPSequence, NPSequence]:
        s = np.asarray(regular_sigmas)
        return s, 1 - s

    def regularize(self, normal_sigmas: Sigma) -> NPSequence:
        return np.asarray(normal_sigmas)


# --------------------------------------------------------------------------- caches


@functools.lru_cache
def np_schedule_lru(schedule: "SkrampleSchedule", steps: int) -> NPPoints:
    "Process-wide memo of ``schedule.schedule_np(steps)`` (schedules are hashable)."
    return schedule.schedule_np(steps)


@functools.lru_cache
def schedule_lru(schedule: "SkrampleSchedule", steps: int) -> Sequence[Point]:
    "Process-wide memo of ``schedule.schedule(steps)``."
    return tuple(Point(*row) for row in np_schedule_lru(schedule, steps).tolist())


def _rows_to_points(rows: NPPoints) -> list[Point]:
    return [Point(*row) for row in rows.tolist()]


# --------------------------------------------------------------------------- base protocol


@dataclass(frozen=True)
class SkrampleSchedule(ABC):
    """A continuously variable noise curve.

    ``_points(t)`` takes *noise time* (1 = all noise, 0 = clean); the ``i``-prefixed
    accessors take *inference time* (0 = all noise).  reference:
    skrample/scheduling.py:65-135
    """

    @property
    @abstractmethod
    def space(self) -> SigmaSpace: ...

    @abstractmethod
    def _points(self, t: NPSequence) -> NPPoints: ...

    # -- vector accessors
    def points_np(self, t: Sequence[float] | NPSequence) -> NPPoints:
        return self._points(np.asarray(t, dtype=np.float64).clip(0, 1))

    def ipoints_np(self, t: Sequence[float] | NPSequence) -> NPPoints:
        return self._points(1 - np.asarray(t, dtype=np.float64).clip(0, 1))

    def points(self, t: Sequence[float] | NPSequence) -> Sequence[Point]:
        return _rows_to_points(self.points_np(t))

    def ipoints(self, t: Sequence[float] | NPSequence) -> Sequence[Point]:
        return _rows_to_points(self.ipoints_np(t))

    # -- scalar accessors
    def point(self, t: float) -> Point:
        at = np.expand_dims(np.float64(t).clip(0, 1), 0)
        return Point(*self._points(at)[0].tolist())

    def ipoint(self, t: float) -> Point:
        at = np.expand_dims(1 - np.float64(t).clip(0, 1), 0)
        return Point(*self._points(at)[0].tolist())

    @functools.cached_property
    def point_0(self) -> Point:
        "The clean end of the curve."
        return self.point(0)

    @functools.cached_property
    def point_1(self) -> Point:
        "The all-noise end of the curve."
        return self.point(1)

    def _ipoints_memo(self, times: tuple[float, ...]) -> tuple[Point, ...]:
        """``ipoints`` memoised on the instance (schedules are immutable, so this is a pure function).

        The samplers ask for the same handful of step boundaries again and again while building step
        programs; the reference recomputes the NumPy curve every time (reference: structured.py:33-34).
        """
        try:
            memo = self.__dict__["_skr_ipoints"]
        except KeyError:
            memo = {}
            object.__setattr__(self, "_skr_ipoints", memo)
        found = memo.get(times)
        if found is None:
            if len(memo) > 4096:
                memo.clear()
            found = memo[times] = tuple(self.ipoints(times))
        return found

    def _ipoint_memo(self, time: float) -> Point:
        "``ipoint`` memoised on the instance (scalar evaluation path, kept separate from the vector one)."
        try:
            memo = self.__dict__["_skr_ipoint"]
        except KeyError:
            memo = {}
            object.__setattr__(self, "_skr_ipoint", memo)
        found = memo.get(time)
        if found is None:
            if len(memo) > 4096:
                memo.clear()
            found = memo[time] = self.ipoint(time)
        return found

    def step(self, step: Step) -> DeltaPoint:
        return DeltaPoint(*self.points(step))

    def istep(self, step: Step) -> DeltaPoint:
        return DeltaPoint(*self.ipoints(step))

    # -- whole trajectories (no trailing zero)
    def schedule_np(self, steps: int) -> NPPoints:
        return self._points(np.linspace(1, 0, steps, endpoint=False))

    def schedule(self, steps: int) -> Sequence[Point]:
        return tuple(Point(*row) for row in self.schedule_np(steps).tolist())


def _timestep_axis(t: NPSequence, base_timesteps: int) -> NPSequence:
    "Noise time -> model timestep; a negative base flips the direction."
    return ((1 - t) if base_timesteps < 0 else t) * abs(base_timesteps)


@dataclass(frozen=True)
class ScheduleCommon(SkrampleSchedule):
    "Standalone base curves. reference: skrample/scheduling.py:138-157"

    base_timesteps: int = 1000

    @functools.cached_property
    def all_points(self) -> NPPoints:
        count = abs(self.base_timesteps)
        if count <= 1:  # a 0..1 float schedule: sample it densely instead
            count = 10_000
        return self.points_np(np.linspace(0, 1, count))

    @abstractmethod
    def _sigmas_to_points(self, sigmas: NPSequence, alphas: NPSequence) -> NPPoints: ...


@dataclass(frozen=True)
class FixedSchedule(SkrampleSchedule):
    "Piecewise-linear curve through user points. reference: skrample/scheduling.py:160-177"

    fixed_schedule: Sequence[Point] | NPPoints
    sigma_space: SigmaSpace

    @classmethod
    def from_regular(cls, timesteps: NPSequence, regular_sigmas: NPSequence, sigma_space: SigmaSpace) -> Self:
        return cls(np.stack([timesteps, *sigma_space.normalize(regular_sigmas)], axis=1), sigma_space)

    def _points(self, t: NPSequence) -> NPPoints:
        from scipy.interpolate import make_interp_spline

        knots = np.concatenate([np.asarray(self.fixed_schedule, dtype=np.float64), [[0, 0, 1]]])
        return make_interp_spline(np.linspace(0, 1, len(knots)), knots, k=1, axis=0)(1 - t)

    @property
    def space(self) -> SigmaSpace:
        return self.sigma_space


# --------------------------------------------------------------------------- base curves


@dataclass(frozen=True)
class Scaled(ScheduleCommon):
    """Stable-Diffusion style beta schedule, in closed continuous form.

    reference: skrample/scheduling.py:180-251.  alphas_cumprod(t) is
    ``exp(-T * (int beta + int beta^2 / 2))`` with ``beta(u) = (a + (b-a) u)^k``.
    """

    beta_start: float = 0.00085
    beta_end: float = 0.012
    beta_scale: float = 2

    @property
    def space(self) -> SigmaSpace:
        return VariancePreserving()

    def continuous_alphas_cumprod(self, t: NPSequence) -> NPSequence:
        k = self.beta_scale
        horizon = abs(self.base_timesteps)
        root_start = self.beta_start ** (1 / k)
        root_end = self.beta_end ** (1 / k)
        slope = root_end - root_start

        if abs(slope) < 1e-8:  # constant beta
            beta_val = root_start**k
            integral_beta = beta_val * t
            integral_beta2 = (beta_val**2) * t
        else:
            integral_beta = ((root_start + slope * t) ** (k + 1) - root_start ** (k + 1)) / (slope * (k + 1))
            integral_beta2 = ((root_start + slope * t) ** (2 * k + 1) - root_start ** (2 * k + 1)) / (
                slope * (2 * k + 1)
            )

        return np.exp(-(horizon * (integral_beta + integral_beta2 / 2)))

    def _points(self, t: NPSequence) -> NPPoints:
        acp = self.continuous_alphas_cumprod(t)
        with np.errstate(divide="ignore"):  # ZSNR: acp -> 0 gives sigma = inf -> (1, 0) after atan
            sigmas = np.sqrt((1 - acp) / acp)
        return np.stack([_timestep_axis(t, self.base_timesteps), *self.space.normalize(sigmas)], 1)

    def _sigmas_to_points(self, sigmas: NPSequence, alphas: NPSequence) -> NPPoints:
        table = self.all_points
        return np.stack([np.interp(sigmas, table[:, 1], table[:, 0]), sigmas, alphas], axis=1)


@dataclass(frozen=True)
class ZSNR(Scaled):
    "Zero-terminal-SNR rescale of ``Scaled`` (arXiv 2305.08891 alg. 1). reference: skrample/scheduling.py:254-278"

    def continuous_alphas_cumprod(self, t: NPSequence) -> NPSequence:
        padded = np.sqrt(super().continuous_alphas_cumprod(np.concatenate([[0], t, [1]])))
        first = padded[0].item()
        last = padded[-1].item()
        body = padded[1:-1]
        body -= last
        body *= first / (first - last)
        return body**2


@dataclass(frozen=True)
class Linear(ScheduleCommon):
    "sigma falls linearly from ``sigma_start`` to 0. reference: skrample/scheduling.py:281-318"

    sigma_start: float = 1
    custom_space: SigmaSpace | None = None

    @property
    def space(self) -> SigmaSpace:
        if self.custom_space is not None:
            return self.custom_space
        return FlowMatching() if self.sigma_start <= 1 else VariancePreserving()

    def _points(self, t: NPSequence) -> NPPoints:
        return np.stack(
            [_timestep_axis(t, self.base_timesteps), *self.space.normalize(t * self.sigma_start)],
            axis=1,
        )

    def _sigmas_to_points(self, sigmas: NPSequence, alphas: NPSequence) -> NPPoints:
        progress = (self.sigma_start - sigmas) if self.base_timesteps < 0 else sigmas
        return np.stack([progress * (abs(self.base_timesteps) / self.sigma_start), sigmas, alphas], axis=1)


# --------------------------------------------------------------------------- wrappers over other schedules


@dataclass(frozen=True)
class _PartialSchedule[T: SkrampleSchedule](SkrampleSchedule):
    "A schedule defined in terms of another one."

    base: T

    @property
    @abstractmethod
    def lowest(self) -> T: ...

    @property
    @abstractmethod
    def all(self) -> Sequence[SkrampleSchedule]: ...

    @property
    def space(self) -> SigmaSpace:
        return self.base.space


@dataclass(frozen=True)
class SubSchedule(_PartialSchedule[ScheduleCommon]):
    "Replaces the base curve outright but needs it for scale. reference: skrample/scheduling.py:349-370"

    base: ScheduleCommon

    @property
    def all(self) -> tuple["SubSchedule", ScheduleCommon]:
        return (self, self.base)

    @property
    def lowest(self) -> ScheduleCommon:
        return self.base

    @property
    def base_timesteps(self) -> int:
        return self.base.base_timesteps


class SubSigmas(SubSchedule):
    "Sub-schedules that supply their own regular sigmas. reference: skrample/scheduling.py:373-389"

    @functools.cached_property
    def _base_regular_0(self) -> float:
        return self.base.space.regularize(self.base.point_0.sigma).item()

    @functools.cached_property
    def _base_regular_1(self) -> float:
        return self.base.space.regularize(self.base.point_1.sigma).item()

    @abstractmethod
    def _sub_sigmas(self, t: NPSequence) -> NPSequence: ...

    def _points(self, t: NPSequence) -> NPPoints:
        return self.base._sigmas_to_points(*self.space.normalize(self._sub_sigmas(t)))


@dataclass(frozen=True)
class ScheduleModifier(_PartialSchedule[SkrampleSchedule]):
    "Warps the time axis of another schedule. reference: skrample/scheduling.py:392-474"

    base: SkrampleSchedule

    @abstractmethod
    def _modify(self, t: NPSequence) -> NPSequence: ...

    def _points(self, t: NPSequence) -> NPPoints:
        return self.base._points(self._modify(t))

    @property
    def all_split(self) -> tuple[list["ScheduleModifier"], SubSchedule | None, SkrampleSchedule]:
        "(modifier chain outermost-first, optional sub-schedule, base curve)."
        chain: list[ScheduleModifier] = []
        cursor: SkrampleSchedule = self
        while isinstance(cursor, ScheduleModifier):
            chain.append(cursor)
            cursor = cursor.base
        sub: SubSchedule | None = None
        if isinstance(cursor, SubSchedule):
            sub, cursor = cursor, cursor.base
        return chain, sub, cursor

    @property
    def all(self) -> list["SkrampleSchedule | ScheduleModifier | SubSchedule"]:
        chain, sub, bottom = self.all_split
        return [*chain, *(() if sub is None else (sub,)), bottom]

    @property
    def lowest(self) -> SkrampleSchedule:
        return self.all_split[2]

    @staticmethod
    def stack(
        modifiers: list["ScheduleModifier"],
        sub: SubSchedule | None,
        base: ScheduleCommon | SkrampleSchedule,
    ) -> "ScheduleModifier | SubSchedule | SkrampleSchedule":
        "Inverse of ``all_split``: rebuild the chain around ``base``."
        built: SkrampleSchedule = base
        if sub is not None:
            assert isinstance(base, ScheduleCommon)
            built = replace(sub, base=built)
        for modifier in reversed(modifiers):
            built = replace(modifier, base=built)
        return built

    @staticmethod
    def _matches(candidate: "ScheduleModifier", kind: type, exact: bool) -> bool:
        return type(candidate) is kind or (not exact and isinstance(candidate, kind))

    def find[T: "ScheduleModifier"](self, skrample_schedule: type[T], exact: bool = False) -> T | None:
        for candidate in self.all_split[0]:
            if self._matches(candidate, skrample_schedule, exact):
                return candidate  # type: ignore[return-value]
        return None

    def find_split[T: "ScheduleModifier"](
        self,
        skrample_schedule: type[T],
        exact: bool = False,
    ) -> tuple[list["ScheduleModifier"], T, list["ScheduleModifier"], SubSchedule | None, SkrampleSchedule] | None:
        "(before, match, after, sub, base); the *last* match wins, as in the reference."
        chain, sub, bottom = self.all_split
        hit: T | None = None
        before: list[ScheduleModifier] = []
        after: list[ScheduleModifier] = []
        for candidate in chain:
            if self._matches(candidate, skrample_schedule, exact):
                hit = candidate  # type: ignore[assignment]
            elif hit is None:
                before.append(candidate)
            else:
                after.append(candidate)
        return None if not hit else (before, hit, after, sub, bottom)


@dataclass(frozen=True)
class NoSub(SubSchedule):
    def _points(self, t: NPSequence) -> NPPoints:
        return self.base._points(t)


@dataclass(frozen=True)
class NoMod(ScheduleModifier):
    def _modify(self, t: NPSequence) -> NPSequence:
        return t


# --------------------------------------------------------------------------- sub-sigma curves


def _with_endcaps(t: NPSequence) -> NPSequence:
    "Prefix the 1 and 0 endpoints so a curve can be normalised by its own extremes."
    return np.concatenate([[1, 0], t])


@dataclass(frozen=True)
class Karras(SubSigmas):
    "EDM rho-ramp. reference: skrample/scheduling.py:493-514"

    rho: float = 7.0
    steps: float = 20

    @functools.cached_property
    def _base_regular_s(self) -> float:
        return self.base.space.regularize(self.base.point(1 / self.steps).sigma).item()

    def _sub_sigmas(self, t: NPSequence) -> NPSequence:
        lo, hi = self._base_regular_s, self._base_regular_1
        u = _with_endcaps(t)
        ramp = ((lo ** (1.0 / self.rho)) * (1 - u) + (hi ** (1.0 / self.rho)) * u) ** self.rho
        return normalize(ramp[2:], ramp[0], ramp[1]) * hi


@dataclass(frozen=True)
class Exponential(SubSigmas):
    "Log-linear ('polyexponential' for rho != 1). reference: skrample/scheduling.py:517-538"

    rho: float = 1.0
    steps: float = 20

    @functools.cached_property
    def _base_regular_s(self) -> float:
        return self.base.space.regularize(self.base.point(1 / self.steps).sigma).item()

    def _sub_sigmas(self, t: NPSequence) -> NPSequence:
        lo, hi = self._base_regular_s, self._base_regular_1
        u = _with_endcaps(t) ** self.rho
        ramp = np.exp(np.log(lo) * (1 - u) + np.log(hi) * u)
        return normalize(ramp[2:], ramp[0], ramp[1]) * hi


@dataclass(frozen=True)
class Beta(SubSigmas):
    "Beta-distribution quantiles (arXiv 2407.12173). reference: skrample/scheduling.py:541-558"

    alpha: float = 0.6
    beta: float = 0.6

    def _sub_sigmas(self, t: NPSequence) -> NPSequence:
        from scipy.stats import beta as beta_dist

        quantiles = beta_dist.ppf(np.concatenate([[1], t]), self.alpha, self.beta)
        return normalize(quantiles, quantiles[0])[1:] * self._base_regular_1


@dataclass(frozen=True)
class Probit(SubSigmas):
    "Sigmoid of the normal quantile function. reference: skrample/scheduling.py:561-580"

    scale: float = 3

    def _sub_sigmas(self, t: NPSequence) -> NPSequence:
        from scipy.stats import norm

        probabilities = regularize(_with_endcaps(t), 1 - 1e-8, 0)  # ppf(1) is inf
        curve = sigmoid(norm.ppf(probabilities, scale=self.scale))
        return normalize(curve[2:], *curve[:2]) * self._base_regular_1


# --------------------------------------------------------------------------- time modifiers


@dataclass(frozen=True)
class FlowShift(ScheduleModifier):
    "SD3/Flux time shift ``s t / (1 + (s-1) t)``. reference: skrample/scheduling.py:583-592"

    shift: float = 3.0

    def _modify(self, t: NPSequence) -> NPSequence:
        return self.shift * t / (1 + (self.shift - 1) * t)


@dataclass(frozen=True)
class Hyper(ScheduleModifier):
    "tanh (scale > 0) or sinh (scale < 0) warp. reference: skrample/scheduling.py:595-614"

    scale: float = 2
    tail: bool = True

    def _modify(self, t: NPSequence) -> NPSequence:
        if abs(self.scale) <= 1e-8:
            return t
        spread = regularize(np.concatenate([[1], t]), self.scale, -self.scale * self.tail)
        curve = np.sinh(spread) if self.scale < 0 else np.tanh(spread / math.sqrt(2))
        return normalize(curve[1:], curve[0], -curve[0] * self.tail)


@dataclass(frozen=True)
class Sinner(ScheduleModifier):
    "Monotone sine-wave warp ``sin(x) + x*scale``. reference: skrample/scheduling.py:617-664"

    count: float = -2
    scale: float = 2

    def _modify(self, t: NPSequence) -> NPSequence:
        if abs(self.scale) <= 1e-8 or self.count == math.inf:
            return t
        half_cycles = rescale_positive(self.count * 2 ** math.copysign(1, self.count)) + 1
        phase = np.concatenate([[0, 1], 1 - t]) * (math.pi * half_cycles)
        if self.scale >= 0:
            phase += math.pi
        lift = abs(self.scale) ** -1 + 1
        curve = np.sin(phase) + phase * lift
        return normalize(curve[2:], *curve[:2])
