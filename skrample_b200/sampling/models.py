"""Model-output algebra: what a network predicts, and how a step uses it.

Every diffusion parameterisation is reduced to the same update

    next = sample * Gamma + output * Delta + noise * Zeta

with Gamma/Delta/Zeta float64 host scalars derived from the two schedule
points of the step and the stochasticity ``eta``.  Conversions between
parameterisations (x-hat / epsilon / v / flow) are affine in ``(sample,
output)`` and are described by :class:`~skrample_b200.sampling.program.ConvSpec`
so the fused kernel can apply them in registers.

Public surface and numerics follow reference: skrample/sampling/models.py:10-239.
Tensor arguments on a CUDA device run as one fused launch; floats, NumPy arrays
and CPU tensors are evaluated with ordinary Python arithmetic.
"""

from __future__ import annotations

import abc
import dataclasses
import math
from collections.abc import Callable
from functools import wraps
from typing import Any

from skrample_b200.common import DeltaPoint, Point, Sample

from . import program as pg
from .program import CONV_DIV, CONV_MUL_X, CONV_MUL_Y, CONV_USE_X, ConvSpec

_EPS = 1e-8


def _apply(spec: ConvSpec | None, sample: Any, value: Any) -> Any:
    if spec is None:
        return value
    return pg._apply_conv(spec.flags, (spec.c0, spec.c1, spec.c2), sample, value)


def _run_conv(specs: tuple[ConvSpec | None, ...], sample: Any, value: Any) -> Any:
    "Apply a chain of conversions to ``value`` (fused on device)."
    specs = tuple(s for s in specs if s is not None)
    if not specs:
        return value
    if pg.is_cuda_tensor(value) and pg.is_cuda_tensor(sample) and pg._fusable((sample, value)):
        prog = pg.Program()
        prog.load(pg.X, sample)
        prog.conv(specs[0], value)
        for spec in specs[1:]:
            prog.conv(spec)
        prog.store(pg.P)
        return prog.run()[0]
    for spec in specs:
        value = _apply(spec, sample, value)
    return value


@dataclasses.dataclass(frozen=True)
class DiffusionModel(abc.ABC):
    "Base parameterisation. reference: skrample/sampling/models.py:10-83"

    # ---- conversions -------------------------------------------------------------------------
    def conv_to_x(self, point: Point) -> ConvSpec | None:
        "Affine description of ``to_x`` for the fused kernel; NotImplemented for user models."
        return NotImplemented  # type: ignore[return-value]

    def conv_from_x(self, point: Point) -> ConvSpec | None:
        return NotImplemented  # type: ignore[return-value]

    @abc.abstractmethod
    def to_x[T: Sample](self, sample: T, output: T, point: Point) -> T:
        "output -> x-hat"

    @abc.abstractmethod
    def from_x[T: Sample](self, sample: T, x: T, point: Point) -> T:
        "x-hat -> output"

    # ---- step scalars ------------------------------------------------------------------------
    @abc.abstractmethod
    def gamma(self, delta_point: DeltaPoint, eta: float = 0) -> float: ...

    @abc.abstractmethod
    def delta(self, delta_point: DeltaPoint, eta: float = 0) -> float: ...

    def zeta_ts(self, delta: DeltaPoint, eta: float = 1.0, epsilon: float = _EPS) -> float:
        "Std of the fresh noise injected by an SDE step. reference: models.py:30-38"
        src, dst = delta
        if abs(eta) < epsilon or abs(dst.sigma) < epsilon:
            return 0
        ratio = (src.alpha * dst.sigma) / (dst.alpha * src.sigma)
        variance = (dst.sigma**2) * (1.0 - ratio**2)
        return eta * math.sqrt(max(0.0, variance))

    def zeta(self, delta_point: DeltaPoint, eta: float = 1.0) -> float:
        return self.zeta_ts(delta_point, eta)

    def eta_transform(self, delta_point: DeltaPoint, eta: float = 0) -> DeltaPoint:
        "Shrink the deterministic target sigma to make room for the injected noise. reference: models.py:44-51"
        src, dst = delta_point
        zeta = self.zeta_ts(delta_point, eta)
        if zeta != 0:
            dst = Point(dst.timestep, math.sqrt(max(0.0, dst.sigma**2 - zeta**2)), dst.alpha)
        return DeltaPoint(src, dst)

    def step_scalars(self, delta_point: DeltaPoint, eta: float, with_noise: bool) -> tuple[float, float, float]:
        "(Gamma, Delta, Zeta) with Zeta forced to 0 when no noise term will be applied."
        zeta = self.zeta(delta_point, eta) if with_noise else 0
        return self.gamma(delta_point, eta), self.delta(delta_point, eta), zeta

    # ---- the update and its inverse ----------------------------------------------------------
    def forward[T: Sample](
        self, sample: T, output: T, delta_point: DeltaPoint, noise: T | None = None, eta: float = 0
    ) -> T:
        "sample*Gamma + output*Delta + noise*Zeta. reference: models.py:53-67"
        gamma, delta, zeta = self.step_scalars(delta_point, eta, noise is not None)
        prog = pg.Program()
        prog.load(pg.X, sample)
        prog.load(pg.P, output)
        prog.fwd(gamma, delta, pg.P, noise if zeta != 0 else None, zeta)
        prog.store(pg.R)
        return prog.run()[0]

    def backward[T: Sample](
        self, sample: T, result: T, delta_point: DeltaPoint, noise: T | None = None, eta: float = 0
    ) -> T:
        "(result - sample*Gamma - noise*Zeta) / Delta. reference: models.py:69-83"
        gamma, delta, zeta = self.step_scalars(delta_point, eta, noise is not None)
        prog = pg.Program()
        prog.load(pg.X, sample)
        prog.load(pg.R, result)
        prog.back(gamma, delta, noise if zeta != 0 else None, zeta)
        prog.store(pg.P)
        return prog.run()[0]


class _BuiltinModel(DiffusionModel):
    "Built-in parameterisations route to_x/from_x through their ConvSpec."

    def to_x[T: Sample](self, sample: T, output: T, point: Point) -> T:
        return _run_conv((self.conv_to_x(point),), sample, output)

    def from_x[T: Sample](self, sample: T, x: T, point: Point) -> T:
        return _run_conv((self.conv_from_x(point),), sample, x)


@dataclasses.dataclass(frozen=True)
class DataModel(_BuiltinModel):
    "x-prediction: the network returns the clean sample. reference: models.py:86-106"

    def conv_to_x(self, point: Point) -> ConvSpec | None:
        return None

    def conv_from_x(self, point: Point) -> ConvSpec | None:
        return None

    def gamma(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return dst.sigma / src.sigma

    def delta(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return dst.alpha - src.alpha * dst.sigma / src.sigma


@dataclasses.dataclass(frozen=True)
class NoiseModel(_BuiltinModel):
    "epsilon-prediction. reference: models.py:109-128"

    def conv_to_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_USE_X | CONV_MUL_Y | CONV_DIV, c1=point.sigma, c2=point.alpha)

    def conv_from_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_USE_X | CONV_MUL_Y | CONV_DIV, c1=point.alpha, c2=point.sigma)

    def gamma(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        return delta_point.point_to.alpha / delta_point.point_from.alpha

    def delta(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return dst.sigma - (dst.alpha * src.sigma) / src.alpha


@dataclasses.dataclass(frozen=True)
class FlowModel(_BuiltinModel):
    "flow / u-prediction (SD3, FLUX). reference: models.py:131-152"

    def conv_to_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_USE_X | CONV_MUL_Y | CONV_DIV, c1=point.sigma, c2=point.alpha + point.sigma)

    def conv_from_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_USE_X | CONV_MUL_Y | CONV_DIV, c1=point.alpha + point.sigma, c2=point.sigma)

    def gamma(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return (dst.sigma + dst.alpha) / (src.sigma + src.alpha)

    def delta(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return (src.alpha * dst.sigma - dst.alpha * src.sigma) / (src.alpha + src.sigma)


@dataclasses.dataclass(frozen=True)
class VelocityModel(_BuiltinModel):
    "v-prediction. reference: models.py:155-176"

    def conv_to_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_USE_X | CONV_MUL_X | CONV_MUL_Y, c0=point.alpha, c1=point.sigma)

    def conv_from_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_USE_X | CONV_MUL_X | CONV_DIV, c0=point.alpha, c2=point.sigma)

    def gamma(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return (dst.sigma / src.sigma) * (1 - src.alpha * src.alpha) + dst.alpha * src.alpha

    def delta(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return src.alpha * dst.sigma - dst.alpha * src.sigma


@dataclasses.dataclass(frozen=True)
class FakeModel(DiffusionModel):
    "Marker: spaces that exist only to sample other models in."


@dataclasses.dataclass(frozen=True)
class ScaleX(FakeModel, _BuiltinModel):
    "x-prediction with an exponential bias along the schedule. reference: models.py:184-212"

    bias: float = 3

    def x_scale(self, point: Point) -> float:
        along = point.sigma if self.bias < 0 else point.alpha
        return math.exp(-math.log10(abs(self.bias) + 1) * along)

    def conv_to_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_MUL_Y, c1=self.x_scale(point))

    def conv_from_x(self, point: Point) -> ConvSpec | None:
        return ConvSpec(CONV_DIV, c2=self.x_scale(point))

    def gamma(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return dst.sigma / src.sigma

    def delta(self, delta_point: DeltaPoint, eta: float = 0) -> float:
        src, dst = self.eta_transform(delta_point, eta)
        return (dst.alpha - src.alpha * dst.sigma / src.sigma) * self.x_scale(src)


@dataclasses.dataclass(frozen=True)
class ModelConvert:
    """Re-express a network output in another parameterisation.

    Skipped only when both ends are the *same object*, exactly like
    reference: skrample/sampling/models.py:215-239.
    """

    transform_from: DiffusionModel
    transform_to: DiffusionModel

    @property
    def is_identity(self) -> bool:
        return self.transform_to is self.transform_from

    def specs_to(self, point: Point) -> tuple[ConvSpec | None, ...] | None:
        "ConvSpec chain from -> to, or None when a user-defined model is involved."
        if self.is_identity:
            return ()
        first, second = self.transform_from.conv_to_x(point), self.transform_to.conv_from_x(point)
        if first is NotImplemented or second is NotImplemented:
            return None
        return (first, second)

    def specs_from(self, point: Point) -> tuple[ConvSpec | None, ...] | None:
        if self.is_identity:
            return ()
        first, second = self.transform_to.conv_to_x(point), self.transform_from.conv_from_x(point)
        if first is NotImplemented or second is NotImplemented:
            return None
        return (first, second)

    def output_to[T: Sample](self, sample: T, output_from: T, point: Point) -> T:
        specs = self.specs_to(point)
        if specs is None:
            return self.transform_to.from_x(sample, self.transform_from.to_x(sample, output_from, point), point)
        return _run_conv(specs, sample, output_from)

    def output_from[T: Sample](self, sample: T, output_to: T, point: Point) -> T:
        specs = self.specs_from(point)
        if specs is None:
            return self.transform_from.from_x(sample, self.transform_to.to_x(sample, output_to, point), point)
        return _run_conv(specs, sample, output_to)

    def wrap_model_call[T: Sample](
        self, model: Callable[[T, float, float, float], T]
    ) -> Callable[[T, float, float, float], T]:
        @wraps(model)
        def converted(x: T, t: float, s: float, a: float) -> T:
            return self.output_to(x, model(x, t, s, a), Point(t, s, a))

        return converted
