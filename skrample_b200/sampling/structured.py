"""Structured (stateless, one-call-per-step) samplers as fused step programs.

Each sampler *emits* its step into a :class:`~.program.Program` - convert the
network output to the derivative space, combine it with the history, apply
``sample*Gamma + pred*Delta + noise*Zeta`` - and the whole program runs as ONE
kernel launch on CUDA tensors (or through the generic executor for floats /
ndarrays / CPU tensors).  Predictor-corrector samplers (UniPC, SPC) inline their
sub-samplers into the same program, so even a UniPC-9 SDE step that costs the
reference ~105 device ops is a single pass over memory here.

All scalar coefficient math (lambda ratios, Bashforth weights, the UniP linear
solve, effective order) is host float64 and follows the reference bit for bit:
reference: skrample/sampling/structured.py:16-577.

History handling differs by design: converted predictions (x-hat) of a step are
written once by the fused kernel and remembered on the returned ``SKSamples``
(a non-field attribute), so later steps read ``k`` cached tensors instead of
re-converting ``2k`` raw ones.  ``SKSamples`` built by the caller fall back to
converting on demand.  Inputs are aliased, not deep-copied.
"""

from __future__ import annotations

import contextlib
import math
import threading
from abc import ABC, abstractmethod
from collections.abc import Iterator, Sequence
from dataclasses import dataclass, replace
from typing import Any

import numpy as np

from skrample_b200 import common
from skrample_b200.common import DeltaPoint, Point, Sample, Step, divf, ln, softmax
from skrample_b200.scheduling import SkrampleSchedule

from . import models, plan, traits
from . import program as pg
from .program import A, B, P, R, S, X, Program


@dataclass(frozen=True)
class SampleInput[T: Sample]:
    "What a sampler consumes for one step. reference: structured.py:16-34"

    sample: T
    prediction: T
    step: Step
    noise: T | None

    def delta_point(self, schedule: SkrampleSchedule) -> DeltaPoint:
        return DeltaPoint(*schedule.ipoints(self.step))


@dataclass(frozen=True)
class SKSamples[T: Sample](SampleInput[T]):
    "A finished step; keep these in ``previous`` for multistep samplers. reference: structured.py:37-40"

    final: T


# ------------------------------------------------------------------------------------------------
# emission plumbing


class CannotFuse(Exception):
    "Raised by an emitter when a step cannot be expressed as one program."


class _InRegister:
    __slots__ = ("reg",)

    def __init__(self, reg: int) -> None:
        self.reg = reg


IN_X = _InRegister(X)
"The step's sample is already in register X (produced by earlier ops of the same program)."
IN_P = _InRegister(P)
"The step's (converted) prediction is already in register P."

COMPUTE = "compute"
"Output dtype marker: the kernel's compute type (fp32, or fp64 for fp64 inputs)."


@dataclass
class _View:
    "Emission-time view of a SampleInput whose fields may live in registers."

    sample: Any
    prediction: Any
    step: Step
    noise: Any


class _StepOptions(threading.local):
    "Per-thread overrides the diffusers wrapper sets around ``sample_packed``."

    final_dtype: Any = None  # dtype of `final` (default: dtype of the step's sample)


_OPTIONS = _StepOptions()


@contextlib.contextmanager
def step_options(final_dtype: Any = None) -> Iterator[None]:
    "Write `final` (and a copy of x-hat for predictor-corrector samplers) directly in ``final_dtype``."
    saved = _OPTIONS.final_dtype
    _OPTIONS.final_dtype = final_dtype
    try:
        yield
    finally:
        _OPTIONS.final_dtype = saved


def low_precision_prediction(result: "SKSamples") -> Any:
    "The x-hat copy in the wrapper's dtype that a fused UniPC/SPC step wrote next to its fp32 state (or None)."
    return getattr(result, "_skr_pred_lowp", None)


class _Ctx:
    "One program under construction plus the outputs the caller wants back."

    __slots__ = ("depth", "lowp_slot", "out_dtype", "preserve_p", "prog", "xhat_key", "xhat_slot")

    def __init__(self, like: Any = None) -> None:
        self.prog = Program()
        self.depth = 0
        self.out_dtype = getattr(like, "dtype", None) if pg.is_cuda_tensor(like) else None  # results follow the sample
        if self.out_dtype is not None and _OPTIONS.final_dtype is not None:
            self.out_dtype = _OPTIONS.final_dtype
        self.preserve_p = False  # a later block of the same program still needs P as it is
        self.lowp_slot: int | None = None
        self.xhat_slot: int | None = None
        self.xhat_key: Any = None


_XHAT_ATTR = "_skr_xhat"


def _HALF_DTYPES() -> tuple[Any, ...]:
    import torch

    return (torch.bfloat16, torch.float16)


def _assemble(packed: SampleInput, outs: list[Any], spec: tuple) -> "SKSamples":
    "SKSamples of a fused step from its outputs; spec = (final, sample, prediction, xhat cache, xhat key, low-precision x-hat)."
    final_slot, sample_slot, pred_slot, cache_slot, cache_key, lowp_slot = spec
    result = SKSamples(
        packed.sample if sample_slot is None else outs[sample_slot],
        packed.prediction if pred_slot is None else outs[pred_slot],
        packed.step,
        packed.noise,
        outs[final_slot],
    )
    if cache_slot is not None:
        object.__setattr__(result, _XHAT_ATTR, (cache_key, outs[cache_slot]))
    if lowp_slot is not None:
        object.__setattr__(result, "_skr_pred_lowp", outs[lowp_slot])
    return result


_native: Any = None


def _drive(sampler: Any, packed: SampleInput, model_transform: Any, schedule: Any, previous: Any, build: Any) -> "SKSamples":
    """Run one fused step: replay a cached plan when this exact step was taken before (device tensors only),
    otherwise emit the program with ``build(ctx) -> spec``, run it, and remember the plan."""
    global _native
    sample = packed.sample
    key = None
    if pg.is_cuda_tensor(sample):
        if _native is None:
            from skrample_b200 import native

            _native = native
        # an explicit final_dtype also asks predictor-correctors for a low-precision x-hat copy: a different program
        # from the one a sample of that same dtype gets, so the flag is part of the key
        forced = _OPTIONS.final_dtype
        key = plan.key_for(sampler, packed, model_transform, schedule, previous, (sample.dtype, False) if forced is None else (forced, True))
        hit = plan.lookup(key, sampler, model_transform, schedule)
        if hit is not None:
            compiled = hit.compiled
            outs = None
            fast = _native._fast if _native._fast is not _native._MISSING else _native._fast_module()
            if fast is not None and compiled.fast and not _native.ACCOUNT["on"]:
                # bind by role, fill the Philox key tables, allocate the outputs and launch: one C++ call
                outs = fast.hit(hit.roles, compiled.n_inputs, compiled.n_philox, packed, previous, compiled.fast, False)
                if outs.__class__ is not list:
                    if outs.__class__ is tuple:
                        _native.check(outs[0], "skr_plan_launch")
                    outs = None  # an unseen dtype combination or a tensor that does not qualify: the Python path decides
            if outs is None:
                bound = plan.bind(hit, packed, previous)
                if bound is not None:
                    count = compiled.n_inputs
                    outs = _native.launch_compiled(compiled, bound, None) if len(bound) == count else _native.launch_compiled(compiled, bound[:count], bound[count:])
            if outs is not None:
                # _assemble, inlined: (final, sample, prediction, x-hat cache, x-hat key, low-precision x-hat)
                final_slot, sample_slot, pred_slot, cache_slot, cache_key, lowp_slot = hit.result
                result = SKSamples(
                    sample if sample_slot is None else outs[sample_slot],
                    packed.prediction if pred_slot is None else outs[pred_slot],
                    packed.step,
                    packed.noise,
                    outs[final_slot],
                )
                if cache_slot is not None:
                    result.__dict__[_XHAT_ATTR] = (cache_key, outs[cache_slot])
                if lowp_slot is not None:
                    result.__dict__["_skr_pred_lowp"] = outs[lowp_slot]
                return result
    ctx = _Ctx(sample)
    spec = build(ctx)
    if key is not None:
        ctx.prog.settle_noise()
        if len(ctx.prog.ops) > pg.MAX_OPS or len(ctx.prog.inputs) > pg.MAX_INPUTS or len(ctx.prog.outputs) > pg.MAX_OUTPUTS:
            # deeper compositions than one launch can describe (e.g. SPC over two high-order UniPCs on a long
            # schedule): predictor-correctors fall back to separately fused launches, in the reference's order
            raise CannotFuse(f"step program too large for one launch ({len(ctx.prog.ops)} ops, {len(ctx.prog.inputs)} inputs)")
    outs = ctx.prog.run()
    if key is not None and pg._fusable(ctx.prog.inputs):
        roles = plan.roles_of([*ctx.prog.inputs, *ctx.prog.philox], packed, previous)
        if roles is not None:
            plan.store(key, (sampler, model_transform, schedule), _native.CompiledProgram(ctx.prog), roles, spec)
    return _assemble(packed, outs, spec)


def _convert_for(sampler: Any, model_transform: Any) -> "models.ModelConvert | None":
    """``ModelConvert(model_transform, sampler.derivative_transform)`` (None without a derivative space), remembered on
    the frozen sampler object per model instance: predictor-correctors build it on every call otherwise."""
    held = sampler.__dict__.get("_skr_convert")
    if held is not None and held[0] is model_transform:
        return held[1]
    convert = models.ModelConvert(model_transform, sampler.derivative_transform) if sampler.derivative_transform else None
    object.__setattr__(sampler, "_skr_convert", (model_transform, convert))
    return convert


def _remember_xhat(entry: SKSamples, key: Any, value: Any) -> None:
    object.__setattr__(entry, _XHAT_ATTR, (key, value))


def _recall_xhat(entry: SKSamples, key: Any) -> Any:
    held = getattr(entry, _XHAT_ATTR, None)
    if held is not None and held[0] == key:
        return held[1]
    return None


def _trivial(specs: tuple | None) -> bool:
    return specs is not None and all(s is None for s in specs)


def _point_from(step: Step, schedule: SkrampleSchedule) -> Point:
    "Origin of a step, evaluated the way ``SampleInput.delta_point`` does (both ends in one call)."
    return schedule._ipoints_memo(tuple(step))[0]


# ------------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class StructuredSampler(ABC, traits.SamplingCommon):
    "Stateless sampler protocol. reference: structured.py:43-91"

    @property
    def require_noise(self) -> bool:
        return False

    @property
    def require_previous(self) -> int:
        return 0

    @abstractmethod
    def sample_packed[T: Sample](
        self,
        packed: SampleInput[T],
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples[T]] = (),
    ) -> SKSamples[T]: ...

    def sample[T: Sample](
        self,
        sample: T,
        prediction: T,
        step: Step | tuple[float, float],
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        noise: T | None = None,
        previous: Sequence[SKSamples[T]] = (),
    ) -> SKSamples[T]:
        "Shorthand for :meth:`sample_packed`."
        return self.sample_packed(
            SampleInput(sample, prediction, step if step.__class__ is Step else Step(*step), noise),
            model_transform=model_transform,
            schedule=schedule,
            previous=previous,
        )

    def scale_input[T: Sample](self, sample: T, point: Point) -> T:
        return sample

    # -- fused emission protocol (internal)
    def _emit(
        self,
        ctx: _Ctx,
        view: _View,
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples],
    ) -> None:
        "Append ops that leave this step's ``final`` in register R."
        raise CannotFuse(type(self).__name__)


@dataclass(frozen=True)
class StatedSampler(StructuredSampler):
    "Samplers whose result is only ``final``. reference: structured.py:94-125"

    def _sample_packed[T: Sample](
        self,
        packed: SampleInput[T],
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples[T]],
    ) -> T:
        "Just the final sample.  User subclasses may override this instead of ``_emit``."
        return self.sample_packed(packed, model_transform, schedule, previous).final

    def sample_packed[T: Sample](
        self,
        packed: SampleInput[T],
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples[T]] = (),
    ) -> SKSamples[T]:
        def build(ctx: _Ctx) -> tuple:
            self._emit(ctx, _View(packed.sample, packed.prediction, packed.step, packed.noise), model_transform, schedule, previous)
            final_slot = ctx.prog.store(R, ctx.out_dtype)
            return (final_slot, None, None, ctx.xhat_slot, ctx.xhat_key, None)

        try:
            return _drive(self, packed, model_transform, schedule, previous, build)
        except CannotFuse:
            if type(self)._sample_packed is StatedSampler._sample_packed:
                raise
            final = self._sample_packed(packed, model_transform, schedule, previous)
            return SKSamples(packed.sample, packed.prediction, packed.step, packed.noise, final)


@dataclass(frozen=True)
class StructuredMultistep(traits.HigherOrder, StructuredSampler):
    "Order > 1 support; the caller keeps ``previous``. reference: structured.py:128-149"

    @property
    def require_previous(self) -> int:
        return max(min(self.order, self.max_order()), self.min_order()) - 1

    def effective_order(self, step: Step, previous: Sequence[SKSamples]) -> int:
        "Order usable at this step: ramps up with history and down towards the end."
        position = step.position()
        return max(
            1,
            min(
                self.max_order(),
                round(position + 1),
                self.order,
                len(previous) + 1,
                round(step.amount() - position),
            ),
        )


@dataclass(frozen=True)
class StructuredStochastic(traits.Stochastic, StructuredSampler):
    @property
    def require_noise(self) -> bool:
        return abs(self.stochasticity) > 1e-8


@dataclass(frozen=True)
class StructuredUnified(traits.UnifiedModelling, StructuredStochastic, StructuredMultistep):
    "Shared head of DPM / Adams / UniP: derivative-space conversion of current and past predictions."

    def _convert(self, model_transform: models.DiffusionModel) -> models.ModelConvert | None:
        if self.derivative_transform:
            return models.ModelConvert(model_transform, self.derivative_transform)
        return None

    def _history(
        self,
        convert: models.ModelConvert | None,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples],
        count: int,
    ) -> list[Any]:
        "x-hat of the ``count`` most recent entries, newest first (cached, else converted on demand)."
        found: list[Any] = []
        for entry in reversed(previous[len(previous) - count :] if count else ()):
            if convert is None or convert.is_identity:
                found.append(entry.prediction)
                continue
            origin = _point_from(entry.step, schedule)
            if _trivial(convert.specs_to(origin)):  # e.g. Data -> Data between distinct instances: x-hat IS the prediction
                found.append(entry.prediction)
                continue
            key = (convert.transform_from, convert.transform_to, origin)
            value = _recall_xhat(entry, key)
            if value is None:
                value = convert.output_to(entry.sample, entry.prediction, origin)
                _remember_xhat(entry, key, value)
            found.append(value)
        return found

    def _head(
        self,
        ctx: _Ctx,
        view: _View,
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples],
        order: int,
        origin: Point,
    ) -> tuple[models.DiffusionModel, list[Any]]:
        "Load X, leave the converted current prediction in P, return (forward model, history x-hats)."
        prog = ctx.prog
        convert = self._convert(model_transform)
        specs = () if convert is None else convert.specs_to(origin)
        if specs is None:  # user-defined model space: its own to_x/from_x run first, the rest is fused
            if view.sample is IN_X or view.prediction is IN_P:
                raise CannotFuse("user-defined model space")
            view = replace(view, prediction=convert.output_to(view.sample, view.prediction, origin))  # type: ignore[union-attr]
            specs = ()

        if view.sample is not IN_X:
            prog.load(X, view.sample)
        if view.prediction is IN_P:
            if ctx.preserve_p and not _trivial(specs):
                raise CannotFuse("in-place conversion would clobber a prediction that is still needed")
            for spec in specs:
                prog.conv(spec)
        else:
            live = [s for s in specs if s is not None]
            prog.conv(live[0] if live else None, view.prediction)
            for spec in live[1:]:
                prog.conv(spec)
            if live and ctx.depth == 0 and self.require_previous > 0 and ctx.xhat_slot is None:
                # write x-hat once so later steps do not re-convert this entry
                ctx.xhat_slot = prog.store(P, COMPUTE)
                ctx.xhat_key = (convert.transform_from, convert.transform_to, origin)  # type: ignore[union-attr]

        history = self._history(convert, schedule, previous, order - 1)
        return (convert.transform_to if convert is not None else model_transform), history


def _lambda(point: Point) -> float:
    "Half log-SNR ``ln(alpha / sigma)``."
    return ln(divf(point.alpha, point.sigma))


def _finish(
    prog: Program,
    model: models.DiffusionModel,
    delta: DeltaPoint,
    eta: float,
    pred: int,
    noise: Any,
) -> None:
    gamma, dlt, zeta = model.step_scalars(delta, eta, noise is not None)
    prog.fwd(gamma, dlt, pred, noise if zeta != 0 else None, zeta)


@dataclass(frozen=True)
class Euler(StructuredStochastic, StatedSampler):
    "First order, in the model's own space. reference: structured.py:163-180"

    def _emit(self, ctx, view, model_transform, schedule, previous) -> None:  # noqa: ANN001
        prog = ctx.prog
        if view.sample is not IN_X:
            prog.load(X, view.sample)
        if view.prediction is not IN_P:
            prog.load(P, view.prediction)
        delta = DeltaPoint(*schedule._ipoints_memo(tuple(view.step)))
        _finish(prog, model_transform, delta, self.stochasticity, P, view.noise)


@dataclass(frozen=True)
class DPM(StructuredUnified, StatedSampler):
    """DPM-Solver++ multistep, orders 1-3, optional SDE (arXiv 2211.01095).

    reference: structured.py:183-283
    """

    @staticmethod
    def max_order() -> int:
        return 3

    def _emit(self, ctx, view, model_transform, schedule, previous) -> None:  # noqa: ANN001
        prog = ctx.prog
        delta = DeltaPoint(*schedule._ipoints_memo(tuple(view.step)))
        order = self.effective_order(view.step, previous)
        model, history = self._head(ctx, view, model_transform, schedule, previous, order, delta.point_from)

        pred = P
        if order >= 2:
            lam = _lambda(delta.point_from)
            h = abs(_lambda(delta.point_to) - lam)
            lam_prev = _lambda(schedule._ipoint_memo(previous[-1].step.time_from))
            r = (lam - lam_prev) / h
            prog.mov(B, P)
            if order >= 3:
                lam_prev2 = _lambda(schedule._ipoint_memo(previous[-2].step.time_from))
                r2 = (lam_prev - lam_prev2) / h
                hh = -h
                e = math.expm1(hh)
                c1 = (e / hh - 1.0) / e if e != 0 else 0
                c2 = ((e - hh) / hh**2 - 0.5) / e if e != 0 else 0
                prog.dpm3(history[0], history[1], 1.0 / r, 1.0 / r2, r / (r + r2), 1.0 / (r + r2), c1, c2)
            else:
                prog.dpm2(history[0], 1.0 / r, 0.5)
            pred = A
        _finish(prog, model, delta, self.stochasticity, pred, view.noise)


@dataclass(frozen=True)
class Adams(StructuredUnified, StatedSampler):
    "Adams-Bashforth on the prediction (IPNDM family), orders 1-9. reference: structured.py:286-330"

    @staticmethod
    def max_order() -> int:
        return 9

    def _emit(self, ctx, view, model_transform, schedule, previous) -> None:  # noqa: ANN001
        prog = ctx.prog
        order = self.effective_order(view.step, previous)
        delta = DeltaPoint(*schedule._ipoints_memo(tuple(view.step)))
        model, history = self._head(ctx, view, model_transform, schedule, previous, order, delta.point_from)

        weights = common.bashforth(order)
        prog.acc(weights[0], reg=P, first=True)
        for weight, past in zip(weights[1:], history, strict=False):
            prog.acc(weight, past)
        _finish(prog, model, delta, self.stochasticity, A, view.noise)


@dataclass(frozen=True)
class UniP(StructuredUnified, StatedSampler):
    "The UniPC predictor alone (B(h) = expm1(-h)). reference: structured.py:333-445"

    fast_solve: bool = False
    "Skip the matrix solve for UniP-2 / UniC-1"

    @staticmethod
    def max_order() -> int:
        return 9

    def _rhos(self, rks: list[float], h: float, order: int, corrector: bool) -> list[float]:
        "UniPC coefficient solve. reference: structured.py:407-424"
        if not rks or (order == (1 if corrector else 2) and self.fast_solve):
            return [0.5]
        hh = -h
        big_b = math.expm1(hh)
        phi = big_b / hh - 1
        rows: list[list[float]] = []
        rhs: list[float] = []
        for n in range(1, len(rks) + 1):
            rows.append([math.pow(v, n - 1) for v in rks])
            rhs.append(phi * math.factorial(n) / big_b)
            phi = phi / hh - 1 / math.factorial(n + 1)
        return np.linalg.solve(rows, rhs).tolist()

    def _emit_uni(
        self,
        ctx: _Ctx,
        view: _View,
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples],
        prediction_next: Any = None,
    ) -> None:
        """UniP (``prediction_next`` None) or UniC (``prediction_next`` = IN_P or a value) into R."""
        prog = ctx.prog
        delta = DeltaPoint(*schedule._ipoints_memo(tuple(view.step)))
        order = self.effective_order(view.step, previous)
        corrector = prediction_next is not None

        if not corrector:
            model, history = self._head(ctx, view, model_transform, schedule, previous, order, delta.point_from)
            prog.mov(B, P)
        else:
            convert = self._convert(model_transform)
            specs = () if convert is None else convert.specs_to(delta.point_from)
            model = convert.transform_to if convert is not None else model_transform
            if view.sample is IN_X or view.prediction is IN_P:
                raise CannotFuse("corrector inputs must be materialised")
            if _trivial(specs):
                base = view.prediction
            elif prediction_next is IN_P:
                raise CannotFuse("corrector needs a conversion while P is occupied")
            else:
                base = convert.output_to(view.sample, view.prediction, delta.point_from)  # type: ignore[union-attr]
            prog.load(X, view.sample)
            if prediction_next is not IN_P:
                if _trivial(specs):
                    prog.load(P, prediction_next)
                elif specs is None:
                    prog.load(P, convert.output_to(view.sample, prediction_next, delta.point_from))  # type: ignore[union-attr]
                else:
                    live = [s for s in specs if s is not None]
                    prog.conv(live[0], prediction_next)
                    for spec in live[1:]:
                        prog.conv(spec)
            prog.load(B, base)
            history = self._history(convert, schedule, previous, order - 1)

        lam = _lambda(delta.point_from)
        h = abs(_lambda(delta.point_to) - lam)

        rks: list[float] = []
        raw_rks: list[float] = []
        for n in range(1, order):
            lam_n = _lambda(_point_from(previous[-n].step, schedule))
            rk = (lam_n - lam) / h
            raw_rks.append(rk)
            rks.append(rk if math.isfinite(rk) else 0)
        if corrector:
            rks.append(1.0)

        rhos = self._rhos(rks, h, order, corrector)

        terms = 0
        for past, rk, rho in zip(history, raw_rks, rhos, strict=False):
            prog.uni(past, rk, rho, first=terms == 0)
            terms += 1
        if corrector:
            prog.unic(rhos[terms], first=terms == 0)
            terms += 1
        prog.addb(empty=terms == 0)
        _finish(prog, model, delta, self.stochasticity, A, view.noise)

    def unisolve[T: Sample](
        self,
        packed: SampleInput[T],
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples[T]],
        prediction_next: Sample | None = None,
    ) -> T:
        "Passing ``prediction_next`` makes this UniC, otherwise UniP. reference: structured.py:344-436"
        ctx = _Ctx(packed.sample)
        ctx.depth = 1  # standalone solve: no x-hat cache side output
        self._emit_uni(
            ctx,
            _View(packed.sample, packed.prediction, packed.step, packed.noise),
            model_transform,
            schedule,
            previous,
            prediction_next,
        )
        slot = ctx.prog.store(R, ctx.out_dtype)
        return ctx.prog.run()[slot]

    def _emit(self, ctx, view, model_transform, schedule, previous) -> None:  # noqa: ANN001
        self._emit_uni(ctx, view, model_transform, schedule, previous)


def _converted_current(
    ctx: _Ctx,
    packed: SampleInput,
    convert: models.ModelConvert | None,
    origin: Point,
) -> int | None:
    "Load X = sample and P = x-hat of the current prediction; returns the output slot of x-hat (None = alias)."
    prog = ctx.prog
    specs = () if convert is None else convert.specs_to(origin)
    if specs is None:
        raise CannotFuse("user-defined model space")
    prog.load(X, packed.sample)
    live = [s for s in specs if s is not None]
    prog.conv(live[0] if live else None, packed.prediction)
    for spec in live[1:]:
        prog.conv(spec)
    if not live:
        return None
    slot = prog.store(P, COMPUTE)
    if _OPTIONS.final_dtype is not None and ctx.out_dtype in _HALF_DTYPES():
        ctx.lowp_slot = prog.store(P, ctx.out_dtype)  # what the pipeline gets back as pred_original_sample
    return slot


@dataclass(frozen=True)
class UniPC(UniP):
    """UniP predictor + UniC corrector of the previous step, fused into one program (arXiv 2302.04867).

    reference: structured.py:448-497
    """

    predictor: StructuredSampler | None = None
    "Defaults to UniP with this sampler's settings"

    @staticmethod
    def max_order() -> int:
        return 9

    @property
    def require_noise(self) -> bool:
        return super().require_noise or (self.predictor.require_noise if self.predictor else False)

    @property
    def require_previous(self) -> int:
        return max(super().require_previous + 1, self.predictor.require_previous if self.predictor else 0)

    def sample_packed[T: Sample](
        self,
        packed: SampleInput[T],
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples[T]] = (),
    ) -> SKSamples[T]:
        convert = _convert_for(self, model_transform)
        inner_model = convert.transform_to if convert is not None else model_transform
        def build(ctx: _Ctx) -> tuple:
            ctx.depth = 1
            prog = ctx.prog
            origin = _point_from(packed.step, schedule)
            xhat_slot = _converted_current(ctx, packed, convert, origin)
            sample_slot = None
            if previous:
                last = previous[-1]
                self._emit_uni(
                    ctx, _View(last.sample, last.prediction, last.step, last.noise), inner_model, schedule, previous[:-1], IN_P
                )
                sample_slot = prog.store(R, COMPUTE)  # solver state stays in compute precision
                prog.mov(X, R)
            view = _View(IN_X, IN_P, packed.step, packed.noise)
            if self.predictor is None:
                UniP._emit(self, ctx, view, inner_model, schedule, previous)
            else:
                self.predictor._emit(ctx, view, inner_model, schedule, previous)
            final_slot = prog.store(R, ctx.out_dtype)
            return (final_slot, sample_slot, xhat_slot, None, None, ctx.lowp_slot)

        try:
            return _drive(self, packed, model_transform, schedule, previous, build)
        except CannotFuse:
            pass

        # Composition of separately fused launches, in the reference's order of operations.
        if convert is not None:
            origin = _point_from(packed.step, schedule)
            packed = replace(packed, prediction=convert.output_to(packed.sample, packed.prediction, origin))
        if previous:
            corrected = self.unisolve(previous[-1], inner_model, schedule, previous[:-1], prediction_next=packed.prediction)
            packed = replace(packed, sample=corrected)
        if self.predictor is None:
            return StatedSampler.sample_packed(self, packed, inner_model, schedule, previous)
        return self.predictor.sample_packed(packed, inner_model, schedule, previous)


@dataclass(frozen=True)
class SPC(traits.DerivativeTransform, StructuredSampler):
    """Simple predictor-corrector: blend the sample with a correction of the previous step.

    reference: structured.py:500-577
    """

    predictor: StructuredSampler = Euler()
    corrector: StructuredSampler = Adams(order=4)
    bias: float = 0
    power: float = 1
    adaptive: bool = True
    invert: bool = False

    @property
    def require_noise(self) -> bool:
        return self.predictor.require_noise or self.corrector.require_noise

    @property
    def require_previous(self) -> int:
        return max(self.predictor.require_previous, self.corrector.require_previous + 1)

    def _weights(self, origin: Point) -> tuple[float, float]:
        p, c = (origin.sigma, origin.alpha) if self.adaptive else (0, 0)
        p, c = softmax((p - self.bias, c + self.bias))
        return (c, p) if self.invert else (p, c)

    def sample_packed[T: Sample](
        self,
        packed: SampleInput[T],
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        previous: Sequence[SKSamples[T]] = (),
    ) -> SKSamples[T]:
        convert = _convert_for(self, model_transform)
        inner_model = convert.transform_to if convert is not None else model_transform
        origin = _point_from(packed.step, schedule)
        def build(ctx: _Ctx) -> tuple:
            ctx.depth = 1
            prog = ctx.prog
            xhat_slot = _converted_current(ctx, packed, convert, origin)
            sample_slot = None
            if previous:
                prog.mov(S, X)
                shifted = [replace(p, prediction=n.prediction) for p, n in zip(previous[:-1], previous[1:], strict=True)]
                last = previous[-1]
                ctx.preserve_p = True
                self.corrector._emit(ctx, _View(last.sample, IN_P, last.step, last.noise), inner_model, schedule, shifted)
                ctx.preserve_p = False
                prog.blend(*self._weights(origin), self.power)
                sample_slot = prog.store(X, COMPUTE)  # solver state stays in compute precision
            self.predictor._emit(ctx, _View(IN_X, IN_P, packed.step, packed.noise), inner_model, schedule, previous)
            final_slot = prog.store(R, ctx.out_dtype)
            return (final_slot, sample_slot, xhat_slot, None, None, ctx.lowp_slot)

        try:
            return _drive(self, packed, model_transform, schedule, previous, build)
        except CannotFuse:
            pass

        # Composition of separately fused launches, in the reference's order of operations.
        if convert is not None:
            packed = replace(packed, prediction=convert.output_to(packed.sample, packed.prediction, origin))
        if previous:
            shifted = [
                replace(p, prediction=pred)
                for p, pred in zip(previous, (*(p.prediction for p in previous[1:]), packed.prediction), strict=True)
            ]
            corrected = self.corrector.sample_packed(shifted.pop(), inner_model, schedule, shifted).final
            p, c = self._weights(origin)
            blend = Program()
            blend.load(S, packed.sample)
            blend.load(R, corrected)
            blend.blend(p, c, self.power)
            blend.store(X)
            packed = replace(packed, sample=blend.run()[0])
        return self.predictor.sample_packed(packed, inner_model, schedule, previous)
