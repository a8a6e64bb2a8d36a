"""Hugging Face ``diffusers`` scheduler facade over the fused samplers.

A ``DiffusionPipeline`` drives these objects exactly like its own schedulers
(``set_timesteps`` / ``timesteps`` / ``step`` / ``scale_noise`` / ``add_noise`` / ``config``); the public surface
and bookkeeping follow reference: skrample/diffusers.py:28-1041.

What is different is the tensor path of ``step``.  The reference casts sample and model output to
``compute_scale`` (two passes), runs the sampler as dozens of elementwise ops, and casts two results back (two
more passes).  Here the step is ONE kernel: it reads the pipeline's tensors in their storage dtype (bf16/fp16/
fp32), computes in fp32 in the reference's operation order, writes the next sample directly in the pipeline's
dtype and keeps the solver state (x-hat history, corrected samples) in fp32 - value for value what the
reference's ``compute_scale=float32`` path produces.

Config parsing/emission is host-side dictionary plumbing kept compatible with the reference.
"""

from __future__ import annotations

import abc
import contextlib
import dataclasses
import functools
import math
from collections import OrderedDict
from collections.abc import Hashable, Mapping, Sequence
from types import MappingProxyType
from typing import TYPE_CHECKING, Any

import numpy as np
import torch
from numpy.typing import NDArray
from torch import Tensor

import skrample_b200.sampling.structured as sampling
from skrample_b200 import scheduling
from skrample_b200.common import DeltaPoint, MergeStrategy, Point, Sample, Step
from skrample_b200.pytorch.noise import BatchTensorNoise, Random, TensorNoiseCommon, TensorNoiseProps
from skrample_b200.sampling import functional, interface, models, tableaux, traits
from skrample_b200.sampling import program as pg
from skrample_b200.sampling.models import DataModel, DiffusionModel, FlowModel, NoiseModel, VelocityModel
from skrample_b200.sampling.structured import SampleInput, SKSamples, StructuredSampler
from skrample_b200.scheduling import ScheduleCommon, ScheduleModifier, SkrampleSchedule, SubSchedule

if TYPE_CHECKING:
    from diffusers.configuration_utils import ConfigMixin

type BuiltinSkrampleWrapper = SkrampleWrapperScheduler | RKUltraWrapperScheduler | DynasauRKWrapperScheduler
"The concrete wrappers shipped here (custom wrappers may exist besides these)."

# ------------------------------------------------------------------------------------------------------------
# diffusers config <-> skrample objects (reference: diffusers.py:33-230)

DIFFUSERS_CLASS_MAP: dict[str, tuple[type[StructuredSampler], dict[str, Any]]] = {
    "DDIMScheduler": (sampling.Euler, {}),
    "DDPMScheduler": (sampling.Euler, {"stochasticity": True}),
    "DPMSolverMultistepScheduler": (sampling.DPM, {}),
    "DPMSolverSDEScheduler": (sampling.DPM, {"stochasticity": True, "order": 1}),
    "EulerAncestralDiscreteScheduler": (sampling.Euler, {"stochasticity": True}),
    "EulerDiscreteScheduler": (sampling.Euler, {}),
    "FlowMatchEulerDiscreteScheduler": (sampling.Euler, {}),
    "IPNDMScheduler": (sampling.Adams, {"order": 4}),
    "MiniMaxH3Scheduler": (sampling.Euler, {}),
    "UniPCMultistepScheduler": (sampling.UniPC, {}),
}

DIFFUSERS_KEY_MAP: dict[str, str] = {
    "shift": "shift",
    "flow_shift": "shift",
    "solver_order": "order",
    "num_train_timesteps": "base_timesteps",
}
"diffusers key -> skrample key, value untouched"
DIFFUSERS_KEY_MAP_REV: dict[str, str] = {v: k for k, v in DIFFUSERS_KEY_MAP.items()}

DIFFUSERS_VALUE_MAP: dict[tuple[str, Any], tuple[str, Any]] = {
    ("beta_schedule", "linear"): ("beta_scale", 1),
    ("beta_schedule", "scaled_linear"): ("beta_scale", 2),
    ("algorithm_type", "dpmsolver"): ("stochasticity", False),
    ("algorithm_type", "dpmsolver++"): ("stochasticity", False),
    ("algorithm_type", "sde-dpmsolver"): ("stochasticity", True),
    ("algorithm_type", "sde-dpmsolver++"): ("stochasticity", True),
    ("prediction_type", "epsilon"): ("skrample_predictor", NoiseModel()),
    ("prediction_type", "flow"): ("skrample_predictor", FlowModel()),
    ("prediction_type", "sample"): ("skrample_predictor", DataModel()),
    ("prediction_type", "v_prediction"): ("skrample_predictor", VelocityModel()),
    # later entries win
    ("use_flow_sigmas", True): ("skrample_subschedule", None),
    ("use_beta_sigmas", True): ("skrample_subschedule", scheduling.Beta),
    ("use_exponential_sigmas", True): ("skrample_subschedule", scheduling.Exponential),
    ("use_karras_sigmas", True): ("skrample_subschedule", scheduling.Karras),
}
"(diffusers key, value) -> (skrample key, value)"
DIFFUSERS_VALUE_MAP_REV: dict[tuple[str, Any], tuple[str, Any]] = {v: k for k, v in DIFFUSERS_VALUE_MAP.items()}

DEFAULT_FAKE_CONFIG = {
    "base_image_seq_len": 256,
    "base_shift": 0.5,
    "max_image_seq_len": 4096,
    "max_shift": 1.15,
    "use_dynamic_shifting": True,
}
"What pipelines expect to find in ``scheduler.config``"


@dataclasses.dataclass(frozen=True)
class ParsedDiffusersConfig:
    sampler: type[StructuredSampler]
    sampler_props: dict[str, Any]
    schedule: type[SkrampleSchedule]
    schedule_props: dict[str, Any]
    subschedule: type[SubSchedule] | None
    subschedule_props: dict[str, Any]
    schedule_modifiers: list[tuple[type[ScheduleModifier], dict[str, Any]]]
    model: DiffusionModel
    invert_prediction: bool


def _field_names(cls: type) -> list[str]:
    return [f.name for f in dataclasses.fields(cls)]


def _only(keys: list[str], source: dict[str, Any]) -> dict[str, Any]:
    return {k: v for k, v in source.items() if k in keys}


def parse_diffusers_config(
    config: "dict[str, Any] | ConfigMixin",
    sampler: type[StructuredSampler] | None = None,
    schedule: type[SkrampleSchedule] | None = None,
) -> ParsedDiffusersConfig:
    "Read a diffusers scheduler (or its config dict) as skrample classes + constructor arguments."
    class_name = config.get("_class_name", "") if isinstance(config, dict) else type(config).__name__
    if not isinstance(config, dict):
        config = dict(config.config)

    found = {new: config[old] for old, new in DIFFUSERS_KEY_MAP.items() if old in config}
    found |= {
        new_key: new_value
        for (old_key, old_value), (new_key, new_value) in DIFFUSERS_VALUE_MAP.items()
        if old_key in config and config[old_key] == old_value
    }

    if "skrample_predictor" in found:
        model: DiffusionModel = found.pop("skrample_predictor")
    elif "shift" in found:
        model = FlowModel()
    else:
        model = NoiseModel()

    if sampler:
        sampler_props: dict[str, Any] = {}
    else:
        sampler, sampler_props = DIFFUSERS_CLASS_MAP.get(class_name, (sampling.DPM, {}))

    if not schedule:
        if isinstance(model, FlowModel):
            schedule = scheduling.Linear
        elif config.get("rescale_betas_zero_snr", False):
            schedule = scheduling.ZSNR
        else:
            schedule = scheduling.Scaled

    # a Linear schedule for a non-flow model starts where the scaled-beta schedule would
    if "sigma_start" not in found and not isinstance(model, FlowModel) and issubclass(schedule, scheduling.Linear):
        beta = scheduling.Scaled(**_only(_field_names(scheduling.Scaled), found))
        found["sigma_start"] = beta.space.regularize(beta.point_1.sigma).item()

    modifiers: list[tuple[type[ScheduleModifier], dict[str, Any]]] = []
    subschedule: type[SubSchedule] | None = None
    subschedule_props: dict[str, Any] = {}
    if "skrample_subschedule" in found:
        subschedule = found.pop("skrample_subschedule")
        if config.get("use_flow_sigmas", False) is True and subschedule in (scheduling.Karras, scheduling.Exponential):
            subschedule = None  # flow sigmas win over karras/exponential
        subschedule_props = _only(_field_names(subschedule), found) if subschedule else {}

    if isinstance(model, FlowModel) and not subschedule:
        modifiers.append((scheduling.FlowShift, _only(_field_names(scheduling.FlowShift), found)))

    invert = class_name == "MiniMaxH3Scheduler"
    if invert and "base_timesteps" not in found:
        found["base_timesteps"] = -1

    return ParsedDiffusersConfig(
        sampler=sampler,
        sampler_props=sampler_props | _only(_field_names(sampler), found),
        schedule=schedule,
        schedule_props=_only(_field_names(schedule), found),
        subschedule=subschedule,
        subschedule_props=subschedule_props,
        schedule_modifiers=modifiers,
        model=model,
        invert_prediction=invert,
    )


def attr_dict[T: Any](**kwargs: T) -> OrderedDict[str, T]:
    "An OrderedDict whose items are also attributes (what pipelines do with ``scheduler.config``)."
    od = OrderedDict(**kwargs)
    for k, v in od.items():
        setattr(od, k, v)
    return od


def as_diffusers_config(sampler: StructuredSampler, schedule: SkrampleSchedule, model: DiffusionModel) -> dict[str, Any]:
    "Best-effort inverse of :func:`parse_diffusers_config`."
    ours = dataclasses.asdict(sampler)
    ours["skrample_predictor"] = model
    if isinstance(schedule, ScheduleModifier):
        _, sub, _ = schedule.all_split
        if sub is not None:
            ours["skrample_subschedule"] = type(sub)
    else:
        ours |= dataclasses.asdict(schedule)
    renamed = {DIFFUSERS_KEY_MAP_REV[k]: v for k, v in ours.items() if k in DIFFUSERS_KEY_MAP_REV}
    revalued = {
        DIFFUSERS_VALUE_MAP_REV[(k, v)][0]: DIFFUSERS_VALUE_MAP_REV[(k, v)][1]
        for k, v in ours.items()
        if isinstance(v, Hashable) and (k, v) in DIFFUSERS_VALUE_MAP_REV
    }
    return ours | renamed | revalued


def _build_schedule(
    parsed: ParsedDiffusersConfig,
    schedule: type[SkrampleSchedule] | None,
    subschedule: type[SubSchedule] | None,
    schedule_modifiers: list[tuple[type[ScheduleModifier], dict[str, Any]]],
    schedule_props: dict[str, Any],
    subschedule_props: dict[str, Any],
    strategy: MergeStrategy,
) -> SkrampleSchedule:
    "Base schedule -> optional sub-schedule -> merged modifier chain (shared by every from_diffusers_config)."
    built: SkrampleSchedule = (schedule or parsed.schedule)(**parsed.schedule_props | schedule_props)
    sub = subschedule or parsed.subschedule
    if sub is not None and isinstance(built, ScheduleCommon):
        built = sub(built, **parsed.subschedule_props | subschedule_props)
    if isinstance(built, ScheduleCommon | SubSchedule | ScheduleModifier):
        merged = strategy.merge(ours=schedule_modifiers, theirs=parsed.schedule_modifiers, cmp=lambda a, b: a[0] is b[0])
        for modifier, props in merged:
            built = modifier(base=built, **props)
    return built


def _dynamic_schedule(schedule: SkrampleSchedule, steps: int, mu: float | None) -> SkrampleSchedule:
    "Per-run schedule overrides: FlowShift(mu) and Karras/Exponential step counts. reference: diffusers.py:511-532"
    if mu is not None and isinstance(schedule, ScheduleModifier) and (hit := schedule.find_split(scheduling.FlowShift)) is not None:
        before, flow, after, sub, base = hit
        schedule = schedule.stack([*before, dataclasses.replace(flow, shift=math.exp(mu)), *after], sub, base)
    if isinstance(schedule, scheduling.Karras | scheduling.Exponential):
        schedule = dataclasses.replace(schedule, steps=steps)
    elif isinstance(schedule, ScheduleModifier):
        mods, sub, base = schedule.all_split
        if isinstance(sub, scheduling.Karras | scheduling.Exponential):
            schedule = schedule.stack(mods, dataclasses.replace(sub, steps=steps), base)
    return schedule


def _as_float(timestep: float | Tensor) -> float:
    "A Python number for the timestep.  A tensor costs one device->host read, like the reference's ``.item()``."
    return timestep if isinstance(timestep, int | float) else timestep.item()


_HALF = (torch.bfloat16, torch.float16)


def _kernel_upcasts(dtype: torch.dtype, compute_scale: torch.dtype | None) -> bool:
    "True when feeding ``dtype`` storage straight to the kernel equals casting to ``compute_scale`` first."
    return compute_scale is None or dtype == compute_scale or (dtype in _HALF and compute_scale == torch.float32)


def _cast_inputs(compute_scale: torch.dtype | None, *tensors: Tensor) -> tuple[Tensor, ...]:
    """Device tensors: fp32 compute over 16-bit/fp32 storage happens inside the kernel (it reads the storage dtype
    and computes fp32), value-identical to the reference's cast-first; any other combination is cast for real."""
    return tuple(t if _kernel_upcasts(t.dtype, compute_scale) else t.to(compute_scale) for t in tensors)


# ------------------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class SkrampleWrapperCore(abc.ABC):
    "Common scheduler facade. reference: diffusers.py:233-387"

    fused_noise = False
    """Let the step kernel draw plain ``Random`` noise itself (CUDA generators on the sample's device) instead of reading
    the tensor the fill kernel wrote.  Same values either way.  It removes a launch, one write and up to two reads of
    the noise, but adds Philox + Box-Muller work to the step kernel: a win for fp32 latents that are HBM-bound
    (measured free on B200), a loss for small or 16-bit latents where the step kernel is issue-bound."""

    def __post_init__(self) -> None:
        self._steps: int = 50
        self._index: int = 0
        self._device: torch.device = torch.device("cpu")
        self._noise_generator: BatchTensorNoise | None = None
        self._handed_out: list[tuple[Tensor, int, int, int, list[float]]] = []  # timesteps tensors given to the caller: (tensor, base address, byte pitch, version, values)

    @property
    @abc.abstractmethod
    def sigma_space(self) -> scheduling.SigmaSpace: ...

    @property
    @abc.abstractmethod
    def schedule_np(self) -> NDArray[np.float64]: ...

    @property
    @abc.abstractmethod
    def config(self) -> OrderedDict[str, Any]: ...

    @property
    def schedule_pt(self) -> Tensor:
        return torch.from_numpy(self.schedule_np).to(self._device)

    @property
    def timesteps(self) -> Tensor:
        values = self.schedule_np[:, 0]
        tensor = torch.from_numpy(values).to(self._device)
        # Remembered so that `for t in scheduler.timesteps: scheduler.step(out, t, x)` resolves `t` by its address
        # inside this tensor instead of a device->host read (reference: diffusers.py:262-270 does `.item()`, a stream
        # drain on every step).  Holding the tensor keeps its memory from being recycled while the entry lives.
        pitch = 8 * (tensor.stride(0) if tensor.numel() > 1 else 1)  # bytes between consecutive timesteps
        self._handed_out = [*self._handed_out[-3:], (tensor, tensor.data_ptr(), pitch, tensor._version, values.tolist())]
        return tensor

    def _adopt_step_count(self, steps: int | None, timesteps: Any, sigmas: Any, mu: float | None) -> bool:
        """Shared part of ``set_timesteps``: the step count is given or is the length of a custom timestep / sigma
        list (their values are ignored, as in the reference: diffusers.py:494-520); then the schedule is re-derived
        for that count when it is allowed to depend on it.  False when nothing was given."""
        if steps is None:
            given = timesteps if timesteps is not None else sigmas
            if given is None:
                return False
            steps = len(given)
        self._steps = steps
        if self.allow_dynamic:
            self.schedule = _dynamic_schedule(self.schedule, steps, mu)  # type: ignore[attr-defined]
        return True

    def _index_without_sync(self, timestep: Tensor) -> int | None:
        "Index of a one-element view of a timesteps tensor this wrapper handed out, or None."
        if timestep.numel() != 1 or timestep.dtype != torch.float64:
            return None
        address = timestep.data_ptr()
        for tensor, base, pitch, version, values in reversed(self._handed_out):
            offset = address - base
            if 0 <= offset < len(values) * pitch and offset % pitch == 0 and timestep.device == tensor.device and tensor._version == version:
                if len(values) != len(self.schedule_np) or values != self.schedule_np[:, 0].tolist():
                    return None  # the schedule changed since (set_timesteps): fall back to the value
                return values.index(values[offset // pitch])  # first occurrence, like list.index on the value
        return None

    @property
    def sigmas(self) -> Tensor:
        sigmas = torch.from_numpy(self.sigma_space.regularize(self.schedule_np[:, 1])).to(self._device)
        return torch.cat([sigmas, torch.zeros([1], device=sigmas.device, dtype=sigmas.dtype)])  # diffusers' trailing zero

    @property
    def init_noise_sigma(self) -> float:
        return 1

    @property
    def order(self) -> int:
        return 1

    @abc.abstractmethod
    def functional_interface(self) -> tuple[functional.FunctionalSampler, scheduling.SkrampleSchedule, models.DiffusionModel]:
        "The equivalent functional sampler, schedule and model."

    def functional_sample_model[T: Sample](
        self,
        sample: T,
        model: functional.SampleableModel[T],
        steps: int,
        include: slice = slice(None),
        rng: functional.RNG[T] | None = None,
        callback: functional.SampleCallback | None = None,
    ) -> T:
        sampler, schedule, transform = self.functional_interface()
        return sampler.sample_model(sample, model, transform, schedule, steps, include, rng, callback)

    def functional_generate_model[T: Sample](
        self,
        model: functional.SampleableModel[T],
        rng: functional.RNG[T],
        steps: int,
        include: slice = slice(None),
        initial: T | None = None,
        callback: functional.SampleCallback | None = None,
    ) -> T:
        sampler, schedule, transform = self.functional_interface()
        return sampler.generate_model(model, transform, schedule, rng, steps, include, initial, callback)

    def get_step_noise[T: TensorNoiseProps | None](
        self,
        step: Step,
        sample: torch.Tensor,
        noise_type: type[TensorNoiseCommon[T]],
        noise_props: T | None,
        generator: torch.Generator | list[torch.Generator] | None = None,
        dtype: torch.dtype | None = None,
    ) -> torch.Tensor:
        "Per-item seeded noise for this step. reference: diffusers.py:312-346"
        if self._noise_generator is None:
            if isinstance(generator, list) and len(generator) == sample.shape[0]:
                seeds = generator
            elif isinstance(generator, torch.Generator) and sample.shape[0] == 1:
                seeds = [generator]
            else:
                # no usable generator: derive a per-item seed from the item's median element (one host read)
                seeds = [
                    torch.Generator(torch.get_default_device()).manual_seed(
                        int(item.reshape(item.numel())[item.numel() // 2].item() * 1e4 * (step.position() + 1))
                    )
                    for item in sample
                ]
            self._noise_generator = BatchTensorNoise.from_batch_inputs(
                noise_type,
                unit_shape=sample.shape[1:],
                seeds=seeds,  # type: ignore[arg-type]
                props=noise_props,
                dtype=torch.float32 if any(s.device.type == "cpu" for s in seeds) else sample.dtype,
            )
        if sample.is_cuda and self.fused_noise:
            noise = self._noise_generator.lazy(step)
            if pg.is_lazy_noise(noise):
                if noise.device == sample.device:
                    return noise  # drawn inside the step kernel, never written to memory
                noise = noise.materialize()
        else:
            noise = self._noise_generator.generate(step)
        if noise.device != sample.device:
            noise = noise.to(device=sample.device, non_blocking=True)
        want = dtype or sample.dtype
        if noise.dtype != want and not (noise.is_cuda and want == torch.float32 and noise.dtype in (torch.bfloat16, torch.float16)):
            noise = noise.to(dtype=want)  # 16-bit -> fp32 is value preserving, the kernel reads it directly instead
        return noise

    @abc.abstractmethod
    def scale_noise(self, sample: Tensor, timestep: Tensor, noise: Tensor) -> Tensor: ...

    @abc.abstractmethod
    def set_timesteps(
        self,
        num_inference_steps: int | None = None,
        device: torch.device | str | None = None,
        timesteps: Tensor | list[int] | None = None,
        sigmas: Tensor | list[float] | None = None,
        mu: float | None = None,
    ) -> None: ...

    @abc.abstractmethod
    def step(
        self,
        model_output: Tensor,
        timestep: float | Tensor,
        sample: Tensor,
        s_churn: float = 0.0,
        s_tmin: float = 0.0,
        s_tmax: float = float("inf"),
        s_noise: float = 1.0,
        generator: torch.Generator | list[torch.Generator] | None = None,
        return_dict: bool = True,
    ) -> tuple[Tensor, Tensor] | OrderedDict[str, Tensor]: ...

    def set_begin_index(self, begin_index: int = 0) -> None:
        self._index = begin_index

    def add_noise(self, original_samples: Tensor, noise: Tensor, timesteps: Tensor) -> Tensor:
        if len(timesteps) == 0:
            return original_samples
        return self.scale_noise(original_samples, timesteps[0], noise)

    def scale_model_input(self, sample: Tensor, timestep: float | Tensor) -> Tensor:
        return sample

    def time_shift(self, mu: float, sigma: float, t: Tensor) -> Tensor:
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def _index_of(self, timestep: float | Tensor) -> int:
        "Exact-match lookup of a timestep in the current schedule (bit-exact, like the reference's list.index)."
        if isinstance(timestep, Tensor):
            found = self._index_without_sync(timestep)
            if found is not None:
                return found
        return self.schedule_np[:, 0].tolist().index(_as_float(timestep))

    @staticmethod
    def _result(return_dict: bool, prev_sample: Tensor, pred_original: Tensor) -> tuple[Tensor, Tensor] | OrderedDict[str, Tensor]:
        if return_dict:
            return attr_dict(prev_sample=prev_sample, pred_original_sample=pred_original)
        return (prev_sample, pred_original)


@dataclasses.dataclass
class SkrampleWrapperScheduler[T: TensorNoiseProps | None](SkrampleWrapperCore):
    "Structured sampler behind the diffusers scheduler interface. reference: diffusers.py:390-599"

    sampler: StructuredSampler
    schedule: SkrampleSchedule
    model: DiffusionModel = NoiseModel()  # noqa: RUF009 - immutable
    noise_type: type[TensorNoiseCommon[T]] = Random  # type: ignore[assignment]
    noise_props: T | None = None
    compute_scale: torch.dtype | None = torch.float32
    allow_dynamic: bool = True
    "Allow set_timesteps to override FlowShift (``mu``) and Karras/Exponential step counts"
    invert_prediction: bool = False
    "Negate the model output before sampling (additive-derivative models such as MiniMax H3)"
    fake_config: dict[str, Any] = dataclasses.field(default_factory=DEFAULT_FAKE_CONFIG.copy)

    def __post_init__(self) -> None:
        super().__post_init__()
        self._previous: list[SKSamples[Tensor]] = []
        self._schedule = self.schedule  # pristine copy, restored by set_timesteps

    @classmethod
    def from_diffusers_config[N: TensorNoiseProps | None](
        cls,
        config: "dict[str, Any] | ConfigMixin",
        sampler: type[StructuredSampler] | None = None,
        schedule: type[SkrampleSchedule] | None = None,
        subschedule: type[SubSchedule] | None = None,
        schedule_modifiers: list[tuple[type[ScheduleModifier], dict[str, Any]]] = [],
        model: DiffusionModel | None = None,
        noise_type: type[TensorNoiseCommon[N]] = Random,  # type: ignore[assignment]
        compute_scale: torch.dtype | None = torch.float32,
        sampler_props: dict[str, Any] = {},
        noise_props: N | None = None,
        schedule_props: dict[str, Any] = {},
        subschedule_props: dict[str, Any] = {},
        modifier_merge_strategy: MergeStrategy = MergeStrategy.UniqueBefore,
        allow_dynamic: bool = True,
        invert_prediction: bool | None = None,
    ) -> "SkrampleWrapperScheduler[N]":
        parsed = parse_diffusers_config(config=config, sampler=sampler, schedule=schedule)
        return cls(  # type: ignore[return-value]
            (sampler or parsed.sampler)(**parsed.sampler_props | sampler_props),
            _build_schedule(parsed, schedule, subschedule, schedule_modifiers, schedule_props, subschedule_props, modifier_merge_strategy),
            model or parsed.model,
            noise_type=noise_type,  # type: ignore[arg-type]
            noise_props=noise_props,  # type: ignore[arg-type]
            compute_scale=compute_scale,
            fake_config=config.copy() if isinstance(config, dict) else dict(config.config),
            allow_dynamic=allow_dynamic,
            invert_prediction=parsed.invert_prediction if invert_prediction is None else invert_prediction,
        )

    def functional_interface(self) -> tuple[interface.StructuredFunctionalAdapter, scheduling.SkrampleSchedule, models.DiffusionModel]:
        return interface.StructuredFunctionalAdapter(self.sampler), self._schedule, self.model

    @property
    def sigma_space(self) -> scheduling.SigmaSpace:
        return self.schedule.space

    @property
    def schedule_np(self) -> NDArray[np.float64]:
        return scheduling.np_schedule_lru(self.schedule, self._steps)

    @property
    def init_noise_sigma(self) -> float:
        return self.sampler.scale_input(1, Point(*self.schedule_np[0]))

    @property
    def order(self) -> int:
        return 1  # multistep samplers call the model once per step

    @property
    def config(self) -> OrderedDict[str, Any]:
        return attr_dict(**(self.fake_config | as_diffusers_config(self.sampler, self._schedule, self.model)))

    def set_begin_index(self, begin_index: int = 0) -> None:
        super().set_begin_index(begin_index)
        self.fake_config["begin_index"] = begin_index

    def set_timesteps(
        self,
        num_inference_steps: int | None = None,
        device: torch.device | str | None = None,
        timesteps: Tensor | list[int] | None = None,
        sigmas: Tensor | list[float] | None = None,
        mu: float | None = None,
    ) -> None:
        self._index = 0
        self.schedule = self._schedule
        if not self._adopt_step_count(num_inference_steps, timesteps, sigmas, mu):
            return
        self._previous = []
        self._noise_generator = None
        if device is not None:
            self._device = torch.device(device)

    def scale_noise(self, sample: Tensor, timestep: Tensor, noise: Tensor) -> Tensor:
        return self.sampler.add_noise(sample, noise, Point(*self.schedule_np[self._index_of(timestep)]))

    def scale_model_input(self, sample: Tensor, timestep: float | Tensor) -> Tensor:
        return self.sampler.scale_input(sample, Point(*self.schedule_np[self._index_of(timestep)]))

    def step(
        self,
        model_output: Tensor,
        timestep: float | Tensor,
        sample: Tensor,
        s_churn: float = 0.0,
        s_tmin: float = 0.0,
        s_tmax: float = float("inf"),
        s_noise: float = 1.0,
        generator: torch.Generator | list[torch.Generator] | None = None,
        return_dict: bool = True,
    ) -> tuple[Tensor, Tensor] | OrderedDict[str, Tensor]:
        schedule_np = self.schedule_np
        step = Step.from_int(self._index_of(timestep), len(schedule_np))

        noise = None
        if self.sampler.require_noise:
            noise = self.get_step_noise(step, sample, self.noise_type, self.noise_props, generator, self.compute_scale)

        if self.invert_prediction:
            model_output = -model_output
        out_dtype, out_device = model_output.dtype, model_output.device
        fused = model_output.is_cuda and sample.is_cuda
        if fused:
            sample_in, output_in = _cast_inputs(self.compute_scale, sample, model_output)
            packed = SampleInput(sample=sample_in, prediction=output_in, step=step, noise=noise)
            with sampling.step_options(final_dtype=out_dtype if self.compute_scale != torch.float64 else None):
                sampled = self.sampler.sample_packed(packed, self.model, self.schedule, self._previous)
        else:
            packed = SampleInput(
                sample=sample.to(dtype=self.compute_scale), prediction=model_output.to(dtype=self.compute_scale), step=step, noise=noise
            )
            sampled = self.sampler.sample_packed(packed, self.model, self.schedule, self._previous)
        self._previous.append(sampled)
        self._previous = self._previous[max(len(self._previous) - self.sampler.require_previous, 0) :]

        prediction = sampling.low_precision_prediction(sampled)
        if prediction is None:
            prediction = sampled.prediction
        return self._result(
            return_dict,
            sampled.final.to(device=out_device, dtype=out_dtype),
            prediction.to(device=out_device, dtype=out_dtype),
        )


# ------------------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class RKWrapperCore[T: TensorNoiseProps | None, U: functional.FunctionalUnified](SkrampleWrapperCore):
    """Runge-Kutta samplers turned inside out: every ``step()`` call is one stage of the tableau.

    reference: diffusers.py:602-873
    """

    schedule: SkrampleSchedule
    sampler_order: int = traits.UnifiedModelling.order
    stochasticity: float = 0
    model: DiffusionModel = NoiseModel()  # noqa: RUF009 - immutable
    derivative_transform: DiffusionModel | None = traits.UnifiedModelling.derivative_transform
    noise_type: type[TensorNoiseCommon[T]] = Random  # type: ignore[assignment]
    noise_props: T | None = None
    compute_scale: torch.dtype | None = torch.float32
    allow_dynamic: bool = True
    invert_prediction: bool = False
    fake_config: dict[str, Any] = dataclasses.field(default_factory=DEFAULT_FAKE_CONFIG.copy)

    def __post_init__(self) -> None:
        super().__post_init__()
        self._index = 0
        self._derivatives: list[Tensor] = []
        self._sample: Tensor | None = None
        self._schedule = self.schedule

    @abc.abstractmethod
    def functional_sampler(self) -> U: ...

    def functional_interface(self) -> tuple[U, scheduling.SkrampleSchedule, models.DiffusionModel]:
        return (self.functional_sampler(), self._schedule, self.model)

    @abc.abstractmethod
    def tableau(self) -> tableaux.Tableau: ...

    def adjust_steps(self, steps: int) -> int:
        return self.functional_interface()[0].adjust_steps(steps)

    @abc.abstractmethod
    def _schedule_full(self, steps: int) -> Sequence[Point]: ...

    @functools.cached_property
    def all_points(self) -> Sequence[Point]:
        "Every point the network (or the backward reconstruction) is evaluated at, T=1 stages included."
        return self._schedule_full(self._steps)

    @functools.cached_property
    def schedule_np_trim(self) -> scheduling.NPPoints:
        "``all_points`` without the stages that sit on the clean end of the schedule."
        clean = self.schedule.point_0
        kept = [p for p in self.all_points if abs(p.timestep - clean.timestep) > 1e-8 and abs(p.sigma - clean.sigma) > 1e-8]
        trimmed = np.asarray(kept, dtype=np.float64)
        if trimmed.size == 0:
            trimmed = np.asarray(self.all_points, dtype=np.float64)
        return trimmed  # type: ignore[return-value]

    @property
    def sigma_space(self) -> scheduling.SigmaSpace:
        return self.schedule.space

    @property
    def schedule_np(self) -> NDArray[np.float64]:
        return self.schedule_np_trim

    @property
    def order(self) -> int:
        return len(self.tableau().stages)

    @property
    def config(self) -> OrderedDict[str, Any]:
        return attr_dict(**self.fake_config)

    def set_begin_index(self, begin_index: int = 0) -> None:
        assert begin_index % self.order == 0, f"Expected {begin_index=} to be multiple of {self.order=}!"
        super().set_begin_index(begin_index)
        self.fake_config["begin_index"] = begin_index

    def set_timesteps(
        self,
        num_inference_steps: int | None = None,
        device: torch.device | str | None = None,
        timesteps: Tensor | list[int] | None = None,
        sigmas: Tensor | list[float] | None = None,
        mu: float | None = None,
    ) -> None:
        self._index = 0
        self._derivatives.clear()
        with contextlib.suppress(AttributeError):
            del self.all_points
            del self.schedule_np_trim
        self.schedule = self._schedule
        if not self._adopt_step_count(num_inference_steps, timesteps, sigmas, mu):
            return
        self._noise_generator = None
        if device is not None:
            self._device = torch.device(device)

    def scale_noise(self, sample: Tensor, timestep: Tensor, noise: Tensor) -> Tensor:
        return Point(*self.schedule_np[self._index_of(timestep)]).add_noise(sample, noise)

    # -- one stage ------------------------------------------------------------------------------------

    def _finish_or_advance(
        self,
        prog: pg.Program,
        in_register: bool,
        model_transform: DiffusionModel,
        S0: Point,
        S1: Point,
        SN: Point,
        generator: torch.Generator | list[torch.Generator] | None,
        out_dtype: torch.dtype | None,
    ) -> int:
        """Append the ops that follow a newly available derivative: the final update when the tableau is full,
        otherwise the next stage input.  Returns the output slot of that tensor.  reference: diffusers.py:746-796"""
        nodes, weights = self.tableau()
        sample = self._sample
        assert sample is not None
        resident = len(self._derivatives) - 1 if in_register else None
        prog.load(pg.X, sample)
        if len(self._derivatives) == len(weights):
            noise = None
            if abs(self.stochasticity) > 1e-8:
                noise = self.get_step_noise(
                    Step.from_int(self._index // self.order, self._steps), sample, self.noise_type, self.noise_props, generator, self.compute_scale
                )
            functional._emit_combination(prog, self._derivatives, weights, resident)
            gamma, dlt, zeta = model_transform.step_scalars(DeltaPoint(S0, S1), self.stochasticity, noise is not None)
            prog.fwd(gamma, dlt, pg.A, noise if zeta != 0 else None, zeta)
            return prog.store(pg.R, out_dtype)
        couplings = nodes[len(self._derivatives)][1]
        if not couplings:
            raise ValueError
        functional._emit_combination(prog, self._derivatives, couplings, resident)
        prog.diva(math.fsum(couplings))
        delta = DeltaPoint(S0, SN)
        prog.fwd(model_transform.gamma(delta, 0), model_transform.delta(delta, 0), pg.A)
        return prog.store(pg.R, out_dtype)

    def step(
        self,
        model_output: Tensor,
        timestep: float | Tensor,
        sample: Tensor,
        s_churn: float = 0.0,
        s_tmin: float = 0.0,
        s_tmax: float = float("inf"),
        s_noise: float = 1.0,
        generator: torch.Generator | list[torch.Generator] | None = None,
        return_dict: bool = True,
    ) -> tuple[Tensor, Tensor] | OrderedDict[str, Tensor]:
        all_points = self.all_points
        expected = all_points[self._index].timestep
        known = self._index_without_sync(timestep) if isinstance(timestep, Tensor) else None
        if known is not None:  # a view of our own timesteps tensor: compare on the host, no stream drain
            timestep = float(self.schedule_np[known, 0])
        assert timestep == expected, f"Expected timestep {expected} for step {self._index}, got {timestep=}!"
        points = [*all_points, Point(0, 0, 1)]
        if self.invert_prediction:
            model_output = -model_output
        out_dtype, out_device = model_output.dtype, model_output.device
        fused = model_output.is_cuda and sample.is_cuda
        clean = self.schedule.point_0

        def is_clean(index: int) -> bool:
            return index < len(all_points) and (
                abs(all_points[index].timestep - clean.timestep) < 1e-8 or abs(all_points[index].sigma - clean.sigma) < 1e-8
            )

        model_transform = self.derivative_transform if self.derivative_transform else self.model
        convert = models.ModelConvert(self.model, self.derivative_transform) if self.derivative_transform else None

        S0_idx = self._index - len(self._derivatives)
        S1_idx = self._index + self.order - len(self._derivatives)
        SN_idx = self._index + 1

        # The reference converts the network output in the pipeline's dtype, then casts to compute_scale
        # (diffusers.py:817-835).  On the device the conversion is fused into this stage's launch whenever that is
        # value-identical or strictly more precise (16-bit storage, fp32 compute); otherwise it runs first.
        in_kernel = fused and _kernel_upcasts(model_output.dtype, self.compute_scale) and _kernel_upcasts(sample.dtype, self.compute_scale)
        specs = () if convert is None else convert.specs_to(points[self._index])
        prog = pg.Program()
        prediction_slot: int | None = None
        in_register = False
        prediction: Any = None
        if in_kernel and specs is not None:
            sample_in = sample
            final_dtype: Any = "compute" if is_clean(self._index + 1) else out_dtype
            live = [s for s in specs if s is not None]
            if live:
                prog.load(pg.X, sample_in)
                prog.conv(live[0], model_output)
                for spec in live[1:]:
                    prog.conv(spec)
                prediction_slot = prog.store(pg.P, "compute")
                derivative: Any = ("pending", prediction_slot)
                in_register = True
            else:
                derivative = prediction = model_output
        else:
            converted = model_output if convert is None else convert.output_to(sample, model_output, points[self._index])
            prediction = converted
            if fused:
                sample_in, derivative = _cast_inputs(self.compute_scale, sample, converted)
                final_dtype = "compute"
            else:  # CPU tensors: cast-compute-cast exactly like the reference, evaluated by the generic executor
                sample_in, derivative = sample.to(dtype=self.compute_scale), converted.to(dtype=self.compute_scale)
                final_dtype = None
        self._derivatives.append(derivative)
        if self._sample is None:
            self._sample = sample_in

        slot = self._finish_or_advance(
            prog, in_register, model_transform, points[S0_idx], points[S1_idx], points[SN_idx], generator, final_dtype
        )
        finished = len(self._derivatives) == len(self.tableau().weights)
        outs = prog.run()
        sampled = outs[slot]
        if in_register:
            prediction = self._derivatives[-1] = outs[prediction_slot]
        if finished:
            self._derivatives.clear()
            self._sample = None
        self._index += 1

        # stages that sit on the clean end of the schedule are reconstructed instead of asked of the network
        while is_clean(self._index):
            anchor = sample_in if self._sample is None else self._sample
            rebuilt = model_transform.backward(anchor, sampled, DeltaPoint(points[S0_idx], points[S1_idx]))
            self._derivatives.append(rebuilt)
            if self._sample is None:
                self._sample = sample_in
            prog = pg.Program()
            if final_dtype is not None and final_dtype != "compute":
                final_dtype = "compute" if is_clean(self._index + 1) else out_dtype
            slot = self._finish_or_advance(
                prog, False, model_transform, points[S0_idx], points[S1_idx], points[SN_idx + 1], generator, final_dtype
            )
            finished = len(self._derivatives) == len(self.tableau().weights)
            sampled = prog.run()[slot]
            if finished:
                self._derivatives.clear()
                self._sample = None
            self._index += 1

        return self._result(
            return_dict,
            sampled.to(device=out_device, dtype=out_dtype),
            prediction.to(device=out_device, dtype=out_dtype),
        )


def _rk_from_config(cls: type, config: "dict[str, Any] | ConfigMixin", kwargs: dict[str, Any]) -> Any:
    "Shared body of the RK wrappers' from_diffusers_config."
    parsed = parse_diffusers_config(config=config, sampler=None, schedule=kwargs["schedule"])
    built = _build_schedule(
        parsed,
        kwargs.pop("schedule"),
        kwargs.pop("subschedule"),
        kwargs.pop("schedule_modifiers"),
        kwargs.pop("schedule_props"),
        kwargs.pop("subschedule_props"),
        kwargs.pop("modifier_merge_strategy"),
    )
    invert = kwargs.pop("invert_prediction")
    model = kwargs.pop("model")
    return cls(
        built,
        kwargs.pop("sampler_order"),
        kwargs.pop("stochasticity"),
        model or parsed.model,
        fake_config=config.copy() if isinstance(config, dict) else dict(config.config),
        invert_prediction=parsed.invert_prediction if invert is None else invert,
        **kwargs,
    )


@dataclasses.dataclass
class RKUltraWrapperScheduler[T: TensorNoiseProps | None](RKWrapperCore[T, functional.RKUltra]):
    "reference: diffusers.py:876-963"

    providers: Mapping[int, tableaux.TableauProvider] = functional.RKUltra.providers

    @classmethod
    def from_diffusers_config[N: TensorNoiseProps | None](
        cls,
        config: "dict[str, Any] | ConfigMixin",
        schedule: type[SkrampleSchedule] | None = None,
        sampler_order: int = functional.RKUltra.order,
        stochasticity: float = 0,
        subschedule: type[SubSchedule] | None = None,
        schedule_modifiers: list[tuple[type[ScheduleModifier], dict[str, Any]]] = [],
        providers: Mapping[int, tableaux.TableauProvider] = functional.RKUltra.providers,
        model: DiffusionModel | None = None,
        noise_type: type[TensorNoiseCommon[N]] = Random,  # type: ignore[assignment]
        derivative_transform: DiffusionModel | None = functional.RKUltra.derivative_transform,
        compute_scale: torch.dtype | None = torch.float32,
        schedule_props: dict[str, Any] = {},
        subschedule_props: dict[str, Any] = {},
        noise_props: N | None = None,
        modifier_merge_strategy: MergeStrategy = MergeStrategy.UniqueBefore,
        allow_dynamic: bool = True,
        invert_prediction: bool | None = None,
    ) -> "RKUltraWrapperScheduler[N]":
        return _rk_from_config(
            cls,
            config,
            dict(
                schedule=schedule,
                sampler_order=sampler_order,
                stochasticity=stochasticity,
                subschedule=subschedule,
                schedule_modifiers=schedule_modifiers,
                providers=providers,
                model=model,
                noise_type=noise_type,
                derivative_transform=derivative_transform,
                compute_scale=compute_scale,
                schedule_props=schedule_props,
                subschedule_props=subschedule_props,
                noise_props=noise_props,
                modifier_merge_strategy=modifier_merge_strategy,
                allow_dynamic=allow_dynamic,
                invert_prediction=invert_prediction,
            ),
        )

    def functional_sampler(self) -> functional.RKUltra:
        return functional.RKUltra(
            order=self.sampler_order,
            stochasticity=self.stochasticity,
            derivative_transform=self.derivative_transform,
            providers=MappingProxyType(self.providers),
        )

    def tableau(self) -> tableaux.Tableau:
        return self.functional_sampler().tableau()

    def _schedule_full(self, steps: int) -> Sequence[Point]:
        "Run the RK step on plain floats to enumerate every evaluation point (the scalar path of step_tableau)."
        tableau = self.tableau()
        seen: list[Point] = []

        def record(x: float, t: float, s: float, a: float) -> float:
            seen.append(Point(t, s, a))
            return x

        for n in range(steps):
            functional.step_tableau(tableau, 1, record, models.DataModel(), self.schedule, Step.from_int(n, steps), epsilon=-math.inf)
        return seen


@dataclasses.dataclass
class DynasauRKWrapperScheduler[T: TensorNoiseProps | None](RKWrapperCore[T, functional.DynasauRK]):
    "reference: diffusers.py:966-1041"

    @classmethod
    def from_diffusers_config[N: TensorNoiseProps | None](
        cls,
        config: "dict[str, Any] | ConfigMixin",
        schedule: type[SkrampleSchedule] | None = None,
        sampler_order: int = functional.RKUltra.order,
        stochasticity: float = 0,
        subschedule: type[SubSchedule] | None = None,
        schedule_modifiers: list[tuple[type[ScheduleModifier], dict[str, Any]]] = [],
        model: DiffusionModel | None = None,
        noise_type: type[TensorNoiseCommon[N]] = Random,  # type: ignore[assignment]
        derivative_transform: DiffusionModel | None = functional.RKUltra.derivative_transform,
        compute_scale: torch.dtype | None = torch.float32,
        schedule_props: dict[str, Any] = {},
        subschedule_props: dict[str, Any] = {},
        noise_props: N | None = None,
        modifier_merge_strategy: MergeStrategy = MergeStrategy.UniqueBefore,
        allow_dynamic: bool = True,
        invert_prediction: bool | None = None,
    ) -> "DynasauRKWrapperScheduler[N]":
        return _rk_from_config(
            cls,
            config,
            dict(
                schedule=schedule,
                sampler_order=sampler_order,
                stochasticity=stochasticity,
                subschedule=subschedule,
                schedule_modifiers=schedule_modifiers,
                model=model,
                noise_type=noise_type,
                derivative_transform=derivative_transform,
                compute_scale=compute_scale,
                schedule_props=schedule_props,
                subschedule_props=subschedule_props,
                noise_props=noise_props,
                modifier_merge_strategy=modifier_merge_strategy,
                allow_dynamic=allow_dynamic,
                invert_prediction=invert_prediction,
            ),
        )

    def functional_sampler(self) -> functional.DynasauRK:
        return functional.DynasauRK(order=self.sampler_order, stochasticity=self.stochasticity, derivative_transform=self.derivative_transform)

    def tableau(self) -> tableaux.Tableau:
        sampler = self.functional_sampler()
        stages = len(sampler.tableau(Step(0, 1)).stages)
        return sampler.tableau(Step.from_int(self._index // stages, self._steps))

    def _schedule_full(self, steps: int) -> Sequence[Point]:
        seen: list[Point] = []

        def record(x: float, t: float, s: float, a: float) -> float:
            seen.append(Point(t, s, a))
            return x

        self.functional_sample_model(1, record, steps)
        assert len(seen) == self.order * steps
        return seen
