"""Mix-in traits shared by structured and functional samplers.

The names, field order, defaults and base-class order below are the plugin API of the reference
(skrample/sampling/traits.py:9-61): samplers are frozen, hashable dataclasses that users construct by keyword
and compare with ``==``, so those must not drift.  Everything else about this module is local.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

from skrample_b200.common import Point, Sample

from .models import DataModel, DiffusionModel

frozen = dataclass(frozen=True)


@frozen
class SamplingCommon:
    "Noise mixing at a schedule point; device tensors take one fused launch (skr_axpby)."

    def add_noise[T: Sample](self, sample: T, noise: T, point: Point) -> T:
        "``sample * alpha + noise * sigma``"
        return point.add_noise(sample, noise)

    def remove_noise[T: Sample](self, sample: T, noise: T, point: Point) -> T:
        "``(sample - noise * sigma) / alpha``"
        return point.remove_noise(sample, noise)


@frozen
class HigherOrder(ABC):
    "Samplers with a selectable order; the order effective at a step can be lower (warm-up, end of schedule)."

    order: int = 2

    @staticmethod
    def min_order() -> int:
        return 1

    @staticmethod
    @abstractmethod
    def max_order() -> int: ...


@frozen
class Stochastic:
    "``stochasticity`` 0 solves the ODE, 1 the fully stochastic SDE."

    stochasticity: float = 0


@frozen
class DerivativeTransform:
    "Samplers that may combine predictions in a model space other than the network's own."

    derivative_transform: DiffusionModel | None = DataModel()  # noqa: RUF009 - frozen instance, safe to share


@frozen
class UnifiedModelling(DerivativeTransform, Stochastic, HigherOrder):
    "Derivative space + stochasticity + order: the usual combination, with the reference's resolution order."
