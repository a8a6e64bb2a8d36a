"""Mix-in traits shared by structured and functional samplers.

reference: skrample/sampling/traits.py:9-61 (field names, defaults and MRO are
part of the plugin API - samplers are frozen, hashable dataclasses compared
with ``==`` - so they are kept as is).
"""

from __future__ import annotations

import abc
import dataclasses

from skrample_b200 import common

from . import models


@dataclasses.dataclass(frozen=True)
class SamplingCommon:
    def add_noise[T: common.Sample](self, sample: T, noise: T, point: common.Point) -> T:
        "``sample*alpha + noise*sigma`` at ``point`` (one fused launch for device tensors)."
        return point.add_noise(sample, noise)

    def remove_noise[T: common.Sample](self, sample: T, noise: T, point: common.Point) -> T:
        "Inverse of :meth:`add_noise`."
        return point.remove_noise(sample, noise)


@dataclasses.dataclass(frozen=True)
class HigherOrder(abc.ABC):
    order: int = 2
    "Requested solver order; the order actually used at a step can be lower."

    @staticmethod
    def min_order() -> int:
        return 1

    @staticmethod
    @abc.abstractmethod
    def max_order() -> int: ...


@dataclasses.dataclass(frozen=True)
class Stochastic:
    stochasticity: float = 0
    "0 = deterministic ODE, 1 = fully stochastic SDE"


@dataclasses.dataclass(frozen=True)
class DerivativeTransform:
    derivative_transform: models.DiffusionModel | None = models.DataModel()  # noqa: RUF009 - immutable
    "Space the solver combines predictions in."


@dataclasses.dataclass(frozen=True)
class UnifiedModelling(DerivativeTransform, Stochastic, HigherOrder):
    "Order + stochasticity + derivative space, in the reference's field order."
