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
        """Read a flat coefficient list: nodes (unless ``compute_c``), then the weights either before or after the
        coupling rows, rows in order without the empty first one."""
        table = cls.empty(stages)
        pending = list(reversed(coeffs))

        def take(target: MutableSequence[float]) -> None:
            for slot in range(len(target)):
                target[slot] = pending.pop()

        if not compute_c:
            take(table.c)
        if not b_last:
            take(table.b)
        for row in table.a[1:]:
            take(row)
        if b_last:
            take(table.b)
        assert not pending, f"{len(pending)} coefficients left over for a {stages}-stage tableau"
        if compute_c:
            table.compute_c()
        return table

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
    "The tableau as text: one line per stage (node | couplings), a rule, one line per weight row."

    def cell(value: float) -> str:
        sign = "-" if value < 0 else "+"
        return f"{sign}{float(round(abs(value), 4)):<6}"

    body = [f"{cell(stage.c)} | {' '.join(map(cell, stage.a))}" for stage in tableau[0]]
    foot = [" " * 8 + "| " + " ".join(map(cell, row)) for row in tableau[1:]]
    width = max(map(len, body + foot))
    title = [] if label is None else [label.rjust((width + len(label)) // 2)]
    return "\n".join(title + body + ["-" * width] + foot)


def validate_tableau(tab: TableauType, tolerance: float = 1e-12) -> IndexError | ValueError | None:
    """None for a well-formed explicit tableau, else the problem as an exception instance (returned, not raised, like
    the reference): IndexError for a shape problem (stage i needs i couplings, every weight row one entry per stage),
    ValueError for an inconsistent one (node != sum of its row, weights not summing to 1)."""
    count = len(tab.stages)
    for position, (node, couplings) in enumerate(tab.stages):
        if len(couplings) != position:
            return IndexError(f"stage {position} has {len(couplings)} couplings, an explicit method needs {position}: {couplings}")
        defect = abs(node - math.fsum(couplings))
        if defect > tolerance:
            return ValueError(f"stage {position}: node {node} differs from its row sum by {defect} (tolerance {tolerance})")
    for row_number, row in enumerate(tab[1:]):
        if len(row) != count:
            return IndexError(f"weight row {row_number} has {len(row)} entries for {count} stages: {row}")
        defect = abs(1 - math.fsum(row))
        if defect > tolerance:
            return ValueError(f"weight row {row_number} sums to 1 with defect {defect} (tolerance {tolerance}): {row}")
    return None
