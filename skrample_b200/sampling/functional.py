"""Functional (closure-driven) samplers: explicit Runge-Kutta steps as fused stage programs.

``step_tableau`` is the generic explicit RK step in derivative space.  The reference evaluates every stage
input and the final combination as a chain of elementwise ops (51 device ops for a 4-stage step, reference:
skrample/sampling/functional.py:55-105).  Here each *model-call boundary* is ONE fused launch: the kernel
that converts the fresh network output to a derivative ``k_j`` also forms the next stage input
``X_{j+1} = forward(sample, sum_i a_i k_i / sum(a), ...)`` (or the final result with its noise term), so a
4-stage step is 4 launches.  Zero tableau coefficients are skipped on the host (only the sign of an exact
zero could differ).

Sampler classes, provider tables and step bookkeeping mirror the reference's public API
(reference: functional.py:13-472) so user closures keep working unchanged.
"""

from __future__ import annotations

import dataclasses
import math
import threading
from abc import ABC, abstractmethod
from collections.abc import Callable, Mapping
from types import MappingProxyType
from typing import Any

from skrample_b200 import common, scheduling
from skrample_b200.common import RNG, DeltaPoint, Point, Sample, Step

from . import models, tableaux, traits
from . import program as pg
from .program import A, P, R, X, Program

type SampleCallback[T: Sample] = Callable[[T, int, DeltaPoint], Any]
"Return is ignored"
type SampleableModel[T: Sample] = Callable[[T, float, float, float], T]
"sample, timestep, sigma, alpha"

DEFAULT_PROVIDERS: Mapping[int, tableaux.TableauProvider[tableaux.TableauType]] = {
    1: tableaux.RK1.Euler,
    2: tableaux.RK2.Mid,
    3: tableaux.RK2.EES5_MIN,
    4: tableaux.RK2.EES7_MIN,
    5: tableaux.SSP.RK4_5,
    6: tableaux.RKE5.CashKarp,
    7: tableaux.RKZ.Butcher6,
    8: tableaux.SSP.RK3_8,
    10: tableaux.SSP.RK5_10,
    11: tableaux.RKZ.CV8,
    15: tableaux.RKZ.Stepanov10,
}
"Default tableau per *stage count* (not mathematical order). reference: functional.py:18-33"
STABLE_PROVIDERS: Mapping[int, tableaux.TableauProvider[tableaux.TableauType]] = {
    2: tableaux.RKE2.Heun,
    3: tableaux.SSP.RK3_3,
    4: tableaux.RKE3.SSPRK3_4,
    5: tableaux.SSP.RK3_5,
    6: tableaux.SSP.RK3_6,
    7: tableaux.SSP.RK3_7,
}
"Strong-stability-preserving choices. reference: functional.py:34-44"
DEFAULT_EMBEDDED_PROVIDERS: Mapping[int, tableaux.TableauProvider[tableaux.EmbeddedTableau]] = {
    2: tableaux.RKE2.Heun,
    4: tableaux.RKE3.BogackiShampine,
    6: tableaux.RKE5.Fehlberg,
}
"Embedded pairs for the adaptive solver. reference: functional.py:46-52"


# ------------------------------------------------------------------------------------------------------
# the fused explicit RK step


_CHUNK = 24
"Most tensor terms one launch accumulates (the kernel binds at most 32 tensors and 64 ops per program)."


def _emit_combination(prog: Program, derivatives: list[Any], coefficients: tuple[float, ...], in_register: int | None) -> None:
    """A = sum_i k_i * c_i in tableau order, skipping exact-zero coefficients.  ``derivatives[in_register]``
    (if any) is the one still sitting in register P of this program.

    Tableaux with more stages than one launch can bind (Feagin's 25/35-stage methods) accumulate a leading run of
    terms in extra launches; the partial sum round-trips through memory in compute precision, which is exact."""
    terms = [(index, k, c) for index, (k, c) in enumerate(zip(derivatives, coefficients, strict=True)) if c != 0]
    first = True
    tensors = [t for t in terms if t[0] != in_register]
    if len(tensors) > _CHUNK and all(pg.is_cuda_tensor(t[1]) for t in tensors):
        partial = None
        while sum(1 for t in terms if t[0] != in_register) > _CHUNK:
            head = []
            while terms and terms[0][0] != in_register and len(head) < _CHUNK:
                head.append(terms.pop(0))
            if not head:
                break
            sub = Program()
            if partial is not None:
                sub.load(A, partial)
            for n, (_, k, c) in enumerate(head):
                sub.acc(c, k, first=partial is None and n == 0)
            slot = sub.store(A, "compute")
            partial = sub.run()[slot]
        if partial is not None:
            prog.load(A, partial)
            first = False
    for index, k, c in terms:
        if index == in_register:
            prog.acc(c, reg=P, first=first)
        else:
            prog.acc(c, k, first=first)
        first = False
    if first:  # every coefficient was an exact zero: keep a (zero-weight) term so A is defined
        if in_register == 0 or not derivatives:
            prog.acc(0.0, reg=P, first=True)
        else:
            prog.acc(0.0, derivatives[0], first=True)


class _Script:
    """What one explicit RK step launched, recorded the first time a (tableau, models, schedule, step, dtype) is
    stepped on device tensors so later steps only bind pointers (the RK counterpart of sampling/plan.py; emitting
    four stage programs costs ~150 us of Python against ~45 us of kernels on a Flux-sized latent).

    Entries: ("model", input role, point) | ("raw-derivative",) | ("launch", compiled program, input roles,
    patched derivative (index, slot) or None, ((slot, "k" | "x" | "out"), ...)).  A role says where a tensor comes
    from: ("sample",), ("noise",), ("k", i) derivative i, ("x", j) stage input j, ("raw", j) network output j."""

    __slots__ = ("alive", "anchors", "entries", "ok", "where")

    def __init__(self, anchors: tuple) -> None:
        self.anchors = anchors
        self.entries: list[tuple] = []
        self.ok = True
        self.where: dict[int, tuple] = {}
        self.alive: list[Any] = []  # while recording: ids must not be recycled by tensors freed mid-step

    def note(self, value: Any, role: tuple) -> None:
        if pg.is_cuda_tensor(value) and id(value) not in self.where:
            self.where[id(value)] = role
            self.alive.append(value)

    def finish(self) -> None:
        self.where.clear()
        self.alive.clear()

    def roles(self, inputs: list[Any]) -> tuple | None:
        found = tuple(self.where.get(id(value)) for value in inputs)
        return None if any(role is None for role in found) else found


class _Scripts(threading.local):
    def __init__(self) -> None:
        self.known: dict[tuple, _Script] = {}


_scripts = _Scripts()
_MAX_SCRIPTS = 512


def _replay(script: _Script, sample: Any, model: Any, noise: Any) -> tuple | None:
    "Run a recorded RK step on new tensors; None when a tensor does not qualify (the caller emits afresh)."
    from skrample_b200 import native

    ks: list[Any] = []
    xs: list[Any] = []
    raws: list[Any] = []
    results: list[Any] = []

    def resolve(role: tuple) -> Any:
        kind = role[0]
        if kind == "sample":
            return sample
        if kind == "noise":
            return noise
        return (ks if kind == "k" else xs if kind == "x" else raws)[role[1]]

    for entry in script.entries:
        kind = entry[0]
        if kind == "model":
            raws.append(model(resolve(entry[1]), *entry[2]))
        elif kind == "raw-derivative":
            ks.append(raws[-1])
        else:
            _, compiled, roles, patch, produces = entry
            if patch is not None:
                ks.append(None)
            outs = native.launch_compiled(compiled, [resolve(role) for role in roles])
            if outs is None:
                return None
            if patch is not None:
                ks[patch[0]] = outs[patch[1]]
            for slot, what in produces:
                (ks if what == "k" else xs if what == "x" else results).append(outs[slot])
    return tuple(results)


def step_tableau[T: Sample](
    tableau: tableaux.Tableau | tableaux.EmbeddedTableau,
    sample: T,
    model: SampleableModel[T],
    model_transform: models.DiffusionModel,
    schedule: scheduling.SkrampleSchedule,
    step: Step,
    derivative_transform: models.DiffusionModel | None = None,
    noise: T | None = None,
    stochasticity: float = 0,
    epsilon: float = 1e-8,
) -> tuple[T, ...]:
    "One explicit RK step; returns one result per weight row. reference: functional.py:55-105"
    nodes, weights = tableau[0], tableau[1:]

    convert = models.ModelConvert(model_transform, derivative_transform) if derivative_transform else None
    solver_space = derivative_transform if derivative_transform else model_transform

    times = (step[0], step[1], *(step[0] + node[0] * (step[1] - step[0]) for node in nodes))
    S0, S1, *fractions = schedule._ipoints_memo(tuple(float(t) for t in times))
    delta = DeltaPoint(S0, S1)
    out_dtype = sample.dtype if pg.is_cuda_tensor(sample) else None

    # device tensors: replay the recorded launches of this exact step when there are any, else record them
    script: _Script | None = None
    key: tuple | None = None
    if out_dtype is not None and len(nodes) <= _CHUNK and not pg.is_lazy_noise(noise):
        key = (tableau, id(model_transform), id(schedule), id(derivative_transform), tuple(step), out_dtype, noise is not None, stochasticity, epsilon)
        known = _scripts.known.get(key)
        if known is not None:
            if known.anchors[0] is model_transform and known.anchors[1] is schedule and known.anchors[2] is derivative_transform:
                replayed = _replay(known, sample, model, noise)
                if replayed is not None:
                    return replayed
            else:
                del _scripts.known[key]  # an id was recycled by a different object
        script = _Script((model_transform, schedule, derivative_transform))
        script.note(sample, ("sample",))
        script.note(noise, ("noise",))
    stage_inputs = network_outputs = 0

    derivatives: list[Any] = []  # k_0 .. k_{i-1}, already in the solver's space
    fresh: tuple[Any, Any, Point] | None = None  # (stage input, raw network output, point) awaiting conversion

    def open_program() -> tuple[Program, int | None]:
        """Start the launch that follows a model call: convert the fresh output (head of the program)."""
        nonlocal fresh
        prog = Program()
        if fresh is None:
            return prog, None
        stage_input, raw, point = fresh
        fresh = None
        specs = () if convert is None else convert.specs_to(point)
        if specs is None:  # user-defined space: its own code converts, the rest is still fused
            derivatives.append(convert.output_to(stage_input, raw, point))  # type: ignore[union-attr]
            if script is not None:
                script.ok = False
            return prog, None
        live = [s for s in specs if s is not None]
        if not live:
            derivatives.append(raw)
            if script is not None:
                script.entries.append(("raw-derivative",))
            return prog, None
        prog.load(X, stage_input)
        prog.conv(live[0], raw)
        for spec in live[1:]:
            prog.conv(spec)
        slot = prog.store(P, "compute")
        derivatives.append(("pending", slot))
        return prog, len(derivatives) - 1

    def close_program(prog: Program, in_register: int | None, wanted: list[int], what: str) -> list[Any]:
        """Run the launch; patch the derivative computed by its head into the list; return the wanted outputs
        (`what` they are: "k" a derivative, "x" a stage input, "out" results - for the recorded script)."""
        nonlocal stage_inputs
        if not prog.ops:
            return []
        outs = prog.run()
        patch = None
        if in_register is not None:
            patch = (in_register, derivatives[in_register][1])
            derivatives[in_register] = outs[patch[1]]
        if script is not None and script.ok:
            roles = script.roles(prog.inputs)
            if roles is None or prog.philox or not all(pg.is_cuda_tensor(out) for out in outs):
                script.ok = False
            else:
                from skrample_b200 import native

                script.entries.append(("launch", native.CompiledProgram(prog), roles, patch, tuple((slot, what) for slot in wanted)))
                if patch is not None:
                    script.note(outs[patch[1]], ("k", patch[0]))
                for slot in wanted:
                    if what == "k":
                        script.note(outs[slot], ("k", len(derivatives)))
                    elif what == "x":
                        script.note(outs[slot], ("x", stage_inputs))
                        stage_inputs += 1
        return [outs[slot] for slot in wanted]

    def call_model(stage_input: Any, point: Point) -> Any:
        nonlocal network_outputs
        raw = model(stage_input, *point)
        if script is not None and script.ok:
            source = script.where.get(id(stage_input))
            if source is None or id(raw) in script.where or not pg.is_cuda_tensor(raw):
                script.ok = False  # e.g. a model that returns its input or one tensor twice: roles would be ambiguous
            else:
                script.entries.append(("model", source, tuple(point)))
                script.note(raw, ("raw", network_outputs))
                network_outputs += 1
        return raw

    for point, (_, couplings) in zip(fractions, nodes, strict=True):
        backward_stage = abs(point.timestep) < epsilon or abs(point.sigma) < epsilon
        if couplings:
            prog, in_register = open_program()
            prog.load(X, sample)
            _emit_combination(prog, derivatives, couplings, in_register)
            prog.diva(math.fsum(couplings))
            stage_delta = DeltaPoint(S0, point)
            prog.fwd(solver_space.gamma(stage_delta, 0), solver_space.delta(stage_delta, 0), A)
            if backward_stage:
                # the derivative at a sigma = 0 / t = 0 node is reconstructed, not asked of the network
                prog.back(solver_space.gamma(delta, 0), solver_space.delta(delta, 0))
                (k,) = close_program(prog, in_register, [prog.store(P, "compute")], "k")
                derivatives.append(k)
                continue
            (stage_input,) = close_program(prog, in_register, [prog.store(R, out_dtype)], "x")
        else:
            stage_input = sample
            if backward_stage:
                derivatives.append(solver_space.backward(sample, stage_input, delta))
                if script is not None:
                    script.ok = False
                continue
        fresh = (stage_input, call_model(stage_input, point), point)

    prog, in_register = open_program()
    prog.load(X, sample)
    slots: list[int] = []
    for row in weights:
        _emit_combination(prog, derivatives, row, in_register)
        gamma, dlt, zeta = solver_space.step_scalars(delta, stochasticity, noise is not None)
        prog.fwd(gamma, dlt, A, noise if zeta != 0 else None, zeta)
        slots.append(prog.store(R, out_dtype))
    results = tuple(close_program(prog, in_register, slots, "out"))
    if script is not None and script.ok and key is not None:
        if len(_scripts.known) >= _MAX_SCRIPTS:
            _scripts.known.clear()
        script.finish()
        _scripts.known[key] = script
    return results


# ------------------------------------------------------------------------------------------------------
# sampler classes


@dataclasses.dataclass(frozen=True)
class FunctionalSampler(ABC, traits.SamplingCommon):
    "reference: functional.py:108-149"

    @abstractmethod
    def sample_model[T: Sample](
        self,
        sample: T,
        model: SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        steps: int,
        include: slice = slice(None),
        rng: RNG[T] | None = None,
        callback: SampleCallback | None = None,
    ) -> T:
        "Run the noisy ``sample`` through ``model`` over the ``include`` range of ``steps``."

    def generate_model[T: Sample](
        self,
        model: SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        rng: RNG[T],
        steps: int,
        include: slice = slice(None),
        initial: T | None = None,
        callback: SampleCallback | None = None,
    ) -> T:
        "``sample_model`` with the starting noise drawn (and mixed into ``initial``) for you."
        if initial is None and include.start is None:
            sample: T = rng(None)
        else:
            start = schedule.ipoint((include.start or 0) / steps)
            noisy = self.add_noise(0 if initial is None else initial, rng(None), start)  # type: ignore[arg-type]
            sample = noisy / self.add_noise(0.0, 1.0, schedule.point_1)  # rescale by the initial sigma
        return self.sample_model(sample, model, model_transform, schedule, steps, include, rng, callback)


@dataclasses.dataclass(frozen=True)
class FunctionalHigher(traits.HigherOrder, FunctionalSampler):
    def adjust_steps(self, steps: int) -> int:
        "Steps that give roughly the same number of model calls."
        return round(steps / self.order)


@dataclasses.dataclass(frozen=True)
class FunctionalUnified(traits.UnifiedModelling, FunctionalHigher): ...


@dataclasses.dataclass(frozen=True)
class FunctionalSinglestep(FunctionalSampler):
    "One ``step`` per schedule interval. reference: functional.py:163-194"

    @abstractmethod
    def step[T: Sample](
        self,
        sample: T,
        model: SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        step: Step,
        rng: RNG[T] | None = None,
    ) -> T: ...

    def sample_model[T: Sample](
        self,
        sample: T,
        model: SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        steps: int,
        include: slice = slice(None),
        rng: RNG[T] | None = None,
        callback: SampleCallback | None = None,
    ) -> T:
        for n in list(range(steps))[include]:
            step = Step.from_int(n, steps)
            sample = self.step(sample, model, model_transform, schedule, step, rng)
            if callback:
                callback(sample, n, schedule.istep(step))
        return sample


@dataclasses.dataclass(frozen=True)
class FunctionalAdaptive(FunctionalSampler):
    "Error-controlled samplers. reference: functional.py:197-214"

    type Evaluator[T: Sample] = Callable[[T, T], float]

    @staticmethod
    def mae[T: Sample](a: T, b: T) -> float:
        return common.mean(abs(a - b))  # type: ignore[operator]

    @staticmethod
    def mse[T: Sample](a: T, b: T) -> float:
        return common.mean(abs(a - b) ** 2)  # type: ignore[operator]

    evaluator: Evaluator = mse
    threshold: float = 1e-2


def _largest_provider(providers: Mapping[int, Any], order: int) -> Any | None:
    "Provider with the largest key <= order, or None when order is below every key."
    if order < min(providers.keys()):
        return None
    best = max(o for o in providers.keys() if o <= order)
    return providers[best] if best else None


@dataclasses.dataclass(frozen=True)
class RKUltra(FunctionalUnified, FunctionalSinglestep):
    "Explicit Runge-Kutta with a tableau picked by stage budget. reference: functional.py:217-268"

    providers: Mapping[int, tableaux.TableauProvider[tableaux.Tableau | tableaux.EmbeddedTableau]] = MappingProxyType(
        DEFAULT_PROVIDERS
    )

    @staticmethod
    def max_order() -> int:
        return 99

    def tableau(self, order: int | None = None) -> tableaux.Tableau:
        provider = _largest_provider(self.providers, self.order if order is None else order)
        if provider is None:
            return tableaux.RK1.Euler.value
        picked = provider.tableau()
        return tableaux.Tableau(picked.stages, picked.weights)

    def adjust_steps(self, steps: int) -> int:
        stages = self.tableau()[0]
        calls = len(stages)
        # stages that land on the end of the interval re-use the next step's first evaluation point
        adjusted = steps / calls + sum(abs(1 - node[0]) < 1e-8 for node in stages) / calls
        return max(round(adjusted), 1)

    def step[T: Sample](
        self,
        sample: T,
        model: SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        step: Step,
        rng: RNG[T] | None = None,
    ) -> T:
        return step_tableau(
            self.tableau(),
            sample,
            model,
            model_transform,
            schedule,
            step,
            self.derivative_transform,
            rng(step) if rng else None,
            self.stochasticity,
        )[0]


@dataclasses.dataclass(frozen=True)
class DynasauRK(FunctionalUnified, FunctionalSinglestep):
    """RK whose tableau slides from a stable member of a family to a convergent one as sampling proceeds:
    weight ``exp(-(S*amount + s*position) * stages)``.  reference: functional.py:271-349"""

    per_step_decay: float = math.log(0.5) / -2
    total_step_decay: float = math.log(0.5) / -20
    invert: bool = False

    @staticmethod
    def min_order() -> int:
        return 2

    @staticmethod
    def max_order() -> int:
        return 4

    def adjust_steps(self, steps: int) -> int:
        return max(round(steps / self.order), 1)

    def gradient(self, step: Step, stages: int) -> float:
        "1.0 = most stable member, 0.0 = most convergent."
        step = step.normal().clamp()
        decay = math.exp((-self.total_step_decay * step.amount() - self.per_step_decay * step.position()) * stages)
        return abs(self.invert - min(max(decay, 0), 1))

    def _family(self) -> tuple[float, float, Callable[[float], tableaux.Tableau]]:
        if self.order >= 4:
            return 1 / 4 * (2 - math.sqrt(2)), 1 / 14 * (5 - 3 * math.sqrt(2)), tableaux.providers.ees27_tableau
        if self.order >= 3:
            return 0.25, 0.1, tableaux.providers.ees25_tableau
        return 1, 0.5, tableaux.providers.rk2_tableau

    def tableau(self, step: Step) -> tableaux.Tableau:
        high, low, family = self._family()
        weight = self.gradient(step, len(family((high + low) / 2).stages))
        return family(weight * high + (1 - weight) * low)

    def step[T: Sample](
        self,
        sample: T,
        model: SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        step: Step,
        rng: RNG[T] | None = None,
    ) -> T:
        return step_tableau(
            self.tableau(step),
            sample,
            model,
            model_transform,
            schedule,
            step,
            self.derivative_transform,
            rng(step) if rng else None,
            self.stochasticity,
        )[0]


@dataclasses.dataclass(frozen=True)
class RKMoire(traits.DerivativeTransform, FunctionalAdaptive, FunctionalHigher):
    "Adaptive step-size RK on embedded pairs (experimental in the reference too). reference: functional.py:352-472"

    providers: Mapping[int, tableaux.TableauProvider[tableaux.EmbeddedTableau]] = MappingProxyType(
        DEFAULT_EMBEDDED_PROVIDERS
    )
    threshold: float = 1e-4
    initial: float = 1 / 50
    maximum: float = 1 / 4
    adaption: float = 0.3
    discard: float = float("inf")
    rescale_init: bool = True
    rescale_max: bool = False

    @staticmethod
    def min_order() -> int:
        return 2

    @staticmethod
    def max_order() -> int:
        return 99

    def adjust_steps(self, steps: int) -> int:
        return steps

    def tableau(self, order: int | None = None) -> tableaux.EmbeddedTableau:
        provider = _largest_provider(self.providers, self.order if order is None else order)
        return (provider or tableaux.RKE2.Heun).tableau()

    def _relative_error[T: Sample](self, low: T, high: T, tiny: float) -> float:
        "evaluator(low, high) / max(evaluator(0, high), tiny); one fused reduction for device tensors and stock norms."
        stock = getattr(self.evaluator, "__func__", self.evaluator)
        power = 2 if stock is FunctionalAdaptive.mse else 1 if stock is FunctionalAdaptive.mae else 0
        if power and pg.is_cuda_tensor(low) and pg.is_cuda_tensor(high) and low.shape == high.shape and low.dtype == high.dtype:  # type: ignore[union-attr]
            from skrample_b200 import native

            if high.dtype in native.DTYPE_CODE:  # type: ignore[union-attr]
                difference, magnitude = native.error_norms(low, high, power)
                return difference / max(magnitude, tiny)
        return self.evaluator(low, high) / max(self.evaluator(0, high), tiny)

    def sample_model[T: Sample](
        self,
        sample: T,
        model: SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        steps: int,
        include: slice = slice(None),
        rng: RNG[T] | None = None,
        callback: SampleCallback | None = None,
    ) -> T:
        tab = self.tableau()
        calls_vs_heun = len(tab[0]) / 2
        first_stride = self.initial * (calls_vs_heun if self.rescale_init else 1)
        longest = self.maximum * (calls_vs_heun if self.rescale_max else 1)

        stride: int = max(round(steps * first_stride), 1)
        tiny: float = 1e-16
        indices: list[int] = list(range(steps))[include]
        at: int = indices[0]

        while at <= indices[-1]:
            upto = min(at + stride, indices[-1] + 1)
            if upto < steps:
                high, low = step_tableau(
                    tab, sample, model, model_transform, schedule, Step(at / steps, upto / steps), self.derivative_transform
                )
                sigma0, sigma1, sigma2 = schedule.ipoints_np([at / steps, upto / steps, (upto + stride) / steps])[:, 1].tolist()
                slope = abs(sigma0 - sigma1) / abs(sigma1 - sigma2)  # the next interval already differs by this much
                error = self._relative_error(low, high, tiny)
                adjustment: float = (self.threshold / max(error, tiny)) ** self.adaption / slope
                stride = max(round(min(stride * adjustment, steps * longest)), 1)
                if upto - at > stride and 1 / max(adjustment, tiny) > self.discard:
                    continue  # the step just taken was far too long: throw it away
            else:  # last stretch: the error estimate would not be used
                high = step_tableau(
                    tab.unembed(), sample, model, model_transform, schedule, Step(at / steps, 1), self.derivative_transform
                )[0]
            sample = high
            if callback:
                callback(sample, upto - 1, schedule.istep(Step.from_int(at, steps)))
            at = upto
        return sample
