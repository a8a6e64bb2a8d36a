"""Butcher tableau containers (explicit Runge-Kutta methods).

Same public types as reference: skrample/sampling/tableaux/common.py:7-156 - plain hashable tuples, because the
functional samplers and the RK wrappers key caches on them.
"""

from __future__ import annotations

import dataclasses
import math
from collections.abc import MutableSequence, Sequence
from typing import NamedTuple, Self


class Stage(NamedTuple):
    "One row of the Butcher matrix: node ``c`` and the (strictly lower triangular) coupling row ``a``."

    c: float
    a: tuple[float, ...]


class Tableau(NamedTuple):
    stages: tuple[Stage, ...]
    weights: tuple[float, ...]


class EmbeddedTableau(NamedTuple):
    stages: tuple[Stage, ...]
    weights: tuple[float, ...]
    error_weights: tuple[float, ...]

    def unembed(self) -> Tableau:
        return Tableau(self.stages, self.weights)


type TableauType = Tableau | EmbeddedTableau


@dataclasses.dataclass(frozen=True)
class ButcherCoeffs:
    "Mutable work form of a tableau (optionally 1-indexed like most papers print them)."

    one_index: bool
    c: MutableSequence[float]
    a: Sequence[MutableSequence[float]]
    b: MutableSequence[float]

    @classmethod
    def empty(cls, stages: int, fill: float = -math.inf, one_index: bool = False) -> Self:
        size = stages + one_index
        coeffs = cls(one_index, c=[fill] * size, a=[[fill] * n for n in range(size)], b=[fill] * size)
        coeffs.c[one_index] = 0  # the first stage is always evaluated at the start of the step
        return coeffs

    def compute_c(self) -> None:
        "Row-sum condition c_i = sum_j a_ij."
        self.c[:] = [math.fsum(row) for row in self.a]

    def compose(self) -> Tableau:
        skip = int(self.one_index)
        rows = zip(self.c[skip:], self.a[skip:], strict=True)
        return Tableau(tuple(Stage(c, tuple(a[skip:])) for c, a in rows), tuple(self.b[skip:]))

    @classmethod
    def decompose(cls, tableau: Tableau) -> Self:
        return cls(False, c=[s.c for s in tableau.stages], a=[list(s.a) for s in tableau.stages], b=list(tableau.weights))

    @classmethod
    def deserialize(cls, coeffs: list[float], stages: int, compute_c: bool = False, b_last: bool = True) -> Self:
        "Read a flat coefficient list laid out as [c...] [b... if not b_last] a-rows [b... if b_last]."
        t = cls.empty(stages)
        expected = len(t.c) * (not compute_c) + len(t.b) + sum(len(row) for row in t.a)
        assert len(coeffs) == expected
        feed = iter(coeffs)
        if not compute_c:
            t.c[:] = [next(feed) for _ in t.c]
        if not b_last:
            t.b[:] = [next(feed) for _ in t.b]
        for row in t.a[1:]:
            row[:] = [next(feed) for _ in row]
        if compute_c:
            t.compute_c()
        if b_last:
            t.b[:] = [next(feed) for _ in t.b]
        return t

    def serialize(self) -> Sequence[float]:
        return [*self.c, *(x for row in self.a for x in row), *self.b]

    @classmethod
    def from_shu_osher(cls, alphas: Sequence[Sequence[float]], betas: Sequence[Sequence[float]]) -> Self:
        "Shu-Osher (alpha, beta) form -> Butcher form, for strong-stability-preserving methods."
        stages = len(alphas)
        t = cls.empty(stages)

        def entry(row: int, col: int, upto: int) -> float:
            return math.fsum((betas[row][col], *(alphas[row][k] * t.a[k][col] for k in range(col + 1, upto))))

        for i in range(1, stages):
            for j in range(i):
                t.a[i][j] = entry(i - 1, j, i)
        for j in range(stages):
            t.b[j] = entry(stages - 1, j, stages)
        t.compute_c()
        return t


def pretty_tableau(tableau: TableauType, label: str | None = None) -> str:
    def cell(x: float) -> str:
        return f"{'+' if x >= 0 else '-'}{float(round(abs(x), 4)): <6}"

    rows = [f"{cell(c)} | {' '.join(cell(x) for x in a)}" for c, a in tableau[0]]
    sums = ["        | " + " ".join(cell(x) for x in w) for w in tableau[1:]]
    width = max(len(line) for line in (*sums, *rows))
    head = [label.rjust((width + len(label)) // 2)] if label is not None else []
    return "\n".join([*head, *rows, "-" * width, *sums])


def validate_tableau(tab: TableauType, tolerance: float = 1e-12) -> IndexError | ValueError | None:
    "Structural (lower-triangular) and consistency (row sums, weight sums) checks."
    for index, stage in enumerate(tab.stages):
        if index != (stage_len := len(stage.a)):
            return IndexError(f"{index=}, {stage_len=}, {stage=}")
        if tolerance < (stage_err := abs(stage.c - math.fsum(stage[1]))):
            return ValueError(f"{tolerance=}, {stage_err=}, {stage=}")
    for weight in tab[1:]:
        if (stage_count := len(tab.stages)) != (weight_len := len(weight)):
            return IndexError(f"{stage_count=}, {weight_len=}, {weight=}")
        if tolerance < (weight_err := abs(1 - math.fsum(weight))):
            return ValueError(f"{tolerance=}, {weight_err=}, {weight=}")
    return None
