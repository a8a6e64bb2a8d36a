"""Noise generators for stochastic samplers, with in-kernel Philox on CUDA generators.

Public surface follows reference: skrample/pytorch/noise.py:11-466 (``Random``, ``Offset``, ``Pyramid``,
``Colored``, ``Brownian``, ``BatchTensorNoise`` and their ``*Props``).  The result lives on the device of the
``torch.Generator`` handed in, exactly like the reference:

* **CUDA generator** - everything is generated on the device by hand-written kernels
  (``csrc/noise_kernels.cu``): counter-based Philox4x32-10 normals written directly in the storage dtype, the
  Offset broadcast folded into the same kernel, the Pyramid evaluated per element from Philox streams with
  on-the-fly bilinear upsampling and a two-pass (moments, regenerate+normalise) scheme, and the Colored
  spectrum shaped in place around ``torch.fft`` (cuFFT).  No host synchronisation anywhere: the ``rand().item()``
  per pyramid level and the ``if std > eps`` tests of the reference are evaluated on the host RNG / on device.
  Streams are keyed ``(generator.initial_seed(), philox offset)`` and the generator's offset is advanced, so
  draws are reproducible, independent between calls, and independent of how a batch is sharded over GPUs.
* **CPU generator** - the values come from the generator's own stream in the reference's draw order, so a
  seeded CPU generator reproduces the reference's noise bit for bit (this is what its parity tests use).
"""

from __future__ import annotations

import ctypes
import math
import threading
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Self

import torch

from skrample_b200.common import Step, divf, rescale_positive

_MOMENTS = 4 + 2 * 2048
"Doubles in an accumulator of grid-wide sums (SKR_MOMENTS_DOUBLES in the C header)."
_SUBSTREAMS = 64
"Philox stream ids consumed per generate() call on a CUDA generator."
_RESERVED_CALLS = 64
"generate() calls whose stream ids one advance of the generator's offset reserves (see TensorNoiseCommon._tick)."


@dataclass(frozen=True)
class TensorNoiseProps:
    "Immutable settings of a generator; share these, not the generator."


@dataclass
class SkrampleTensorNoise(ABC):
    @abstractmethod
    def generate(self, step: Step | None) -> torch.Tensor:
        "Next noise tensor of the sequence (stateful: one generator per job)."
        raise NotImplementedError


# ------------------------------------------------------------------------------------------------------------
# native plumbing


def _native() -> Any:
    from skrample_b200 import native

    return native


class _SkrOffset(ctypes.Structure):
    _fields_ = [
        ("ndim", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("shape", ctypes.c_int64 * 8),
        ("keep", ctypes.c_int32 * 8),
        ("stream", ctypes.c_uint64),
        ("scale", ctypes.c_double),
    ]


class _SkrPyramidLevel(ctypes.Structure):
    _fields_ = [
        ("stream", ctypes.c_uint64),
        ("buffer", ctypes.c_void_p),
        ("extent", ctypes.c_int64 * 2),
        ("weight", ctypes.c_double),
    ]


class _SkrPyramid(ctypes.Structure):
    _fields_ = [
        ("ndim", ctypes.c_int32),
        ("n_levels", ctypes.c_int32),
        ("shape", ctypes.c_int64 * 8),
        ("masked", ctypes.c_int32 * 8),
        ("seed", ctypes.c_uint64),
        ("base_stream", ctypes.c_uint64),
        ("base_buffer", ctypes.c_void_p),
        ("scratch", ctypes.c_void_p),
        ("levels_scratch", ctypes.c_void_p),
        ("levels_scratch_floats", ctypes.c_int64),
        ("levels", _SkrPyramidLevel * 16),
    ]


_bound = False


def _lib() -> Any:
    global _bound
    native = _native()
    lib = native.load()
    if not _bound:
        vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
        lib.skr_noise_fill.restype = ctypes.c_int
        lib.skr_noise_fill.argtypes = [vp, i32, i64, u64, u64, ctypes.POINTER(_SkrOffset), vp, vp]
        lib.skr_noise_fill_batch.restype = ctypes.c_int
        lib.skr_noise_fill_batch.argtypes = [vp, i32, ctypes.POINTER(native.SkrPhilox), vp]
        lib.skr_noise_brownian.restype = ctypes.c_int
        lib.skr_noise_brownian.argtypes = [vp, i32, i64, u64, dbl, dbl, i32, dbl, vp]
        lib.skr_noise_brownian_batch.restype = ctypes.c_int
        lib.skr_noise_brownian_batch.argtypes = [vp, i32, ctypes.POINTER(u64), i32, i64, dbl, dbl, i32, dbl, vp]
        lib.skr_noise_moments.restype = ctypes.c_int
        lib.skr_noise_moments.argtypes = [vp, i32, i64, vp, vp]
        lib.skr_noise_scale.restype = ctypes.c_int
        lib.skr_noise_scale.argtypes = [vp, i32, vp, i32, i64, dbl, vp, i64, vp, i64, dbl, vp]
        lib.skr_noise_pyramid.restype = ctypes.c_int
        lib.skr_noise_pyramid.argtypes = [vp, i32, ctypes.POINTER(_SkrPyramid), vp, vp]
        lib.skr_colored_shape.restype = ctypes.c_int
        lib.skr_colored_shape.argtypes = [vp, i32, ctypes.POINTER(i64), i32, dbl, vp]
        _bound = True
    return lib


def _same_device(generator_device: torch.device, tensor_device: torch.device) -> bool:
    "A generator made with ``device='cuda'`` carries no index: it lives on the current device."
    if generator_device.type != tensor_device.type:
        return False
    if generator_device.index is None or tensor_device.index is None:
        current = torch.cuda.current_device() if tensor_device.type == "cuda" else None
        return (generator_device.index if generator_device.index is not None else current) == (
            tensor_device.index if tensor_device.index is not None else current
        )
    return generator_device.index == tensor_device.index


def _stream() -> int:
    return _native().raw_stream()


def _code(dtype: torch.dtype) -> int:
    try:
        return _native().DTYPE_CODE[dtype]
    except KeyError:
        raise TypeError(f"skrample_b200 noise kernels support fp32/fp64/bf16/fp16, not {dtype}") from None


class _DeviceGuard:
    "Make ``device`` current for the launches inside (no-op when it already is)."

    def __init__(self, device: torch.device) -> None:
        self.ctx = torch.cuda.device(device) if torch._C._cuda_getDevice() != device.index else None

    def __enter__(self) -> None:
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc: object) -> None:
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _host_uniform(seed: int, stream: int, index: int) -> float:
    "Counter-based U[0,1) on the host (SplitMix64 of the key) - the pyramid's per-level ratio, without a device sync."
    z = (seed * 0x9E3779B97F4A7C15 + stream * 0xBF58476D1CE4E5B9 + index * 0x94D049BB133111EB + 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    z ^= z >> 31
    return (z >> 11) / float(1 << 53)


_scratch = threading.local()


def _philox_scratch() -> Any:
    "One reusable key table per thread (the C call copies it before returning)."
    keys = getattr(_scratch, "keys", None)
    if keys is None:
        keys = _scratch.keys = _native().SkrPhilox()
    return keys


class PhiloxDraw:
    """A normal noise tensor that exists only as Philox keys: the fused step kernel draws it while it runs.

    Passing this as ``noise`` to a sampler removes the noise tensor's write (generation) and every read of it
    (the step itself and, for UniPC, the corrector of the next step, which re-draws it).  ``materialize()``
    produces the bit-identical tensor through ``skr_noise_fill``.  One (seed, stream) pair per batch item.
    ``offset_inner`` / ``offset_scale`` make it ``Offset`` noise (kept axes leading): every run of ``offset_inner``
    elements of an item shares one more normal from the item's next stream, added times ``offset_scale``.
    """

    is_lazy_noise = True
    __slots__ = ("device", "dtype", "item_numel", "numel", "offset_inner", "offset_scale", "seeds", "shape", "streams", "_tensor")

    def __init__(
        self,
        shape: tuple[int, ...],
        seeds: tuple[int, ...],
        streams: tuple[int, ...],
        dtype: torch.dtype,
        device: torch.device,
        numel: int | None = None,
        offset_inner: int = 0,
        offset_scale: float = 0.0,
    ) -> None:
        self.offset_inner = offset_inner
        self.offset_scale = offset_scale
        self.shape = shape if shape.__class__ is tuple else tuple(shape)
        self.seeds = seeds
        self.streams = streams
        self.dtype = dtype
        self.device = device
        self.numel = math.prod(self.shape) if numel is None else numel
        self.item_numel = self.numel // len(seeds)
        self._tensor: torch.Tensor | None = None

    def materialize(self) -> torch.Tensor:
        if self._tensor is None:
            out = torch.empty(self.shape, dtype=self.dtype, device=self.device)
            self.materialize_into(out)
            self._tensor = out
        return self._tensor

    def materialize_into(self, out: torch.Tensor) -> None:
        "Write the draw into ``out`` (contiguous, on this draw's device; rounded once to ``out``'s dtype)."
        keys = _philox_scratch()
        total = len(self.seeds)
        item_numel = self.item_numel
        limit = _native().MAX_PHILOX_ITEMS  # items one launch's key table holds
        code, base, esize = _code(out.dtype), out.data_ptr(), out.element_size()
        with _DeviceGuard(self.device):
            lib, stream = _lib(), _stream()
            for first in range(0, total, limit):
                count = min(limit, total - first)
                keys.n_items = count
                keys.item_numel = item_numel
                keys.offset_inner = self.offset_inner
                keys.offset_scale = self.offset_scale
                keys.seed[:count] = self.seeds[first : first + count]
                keys.stream[:count] = self.streams[first : first + count]
                status = lib.skr_noise_fill_batch(base + first * item_numel * esize, code, ctypes.byref(keys), stream)
                if status:
                    _native().check(status, "skr_noise_fill_batch")

    def to(self, *args: Any, **kwargs: Any) -> torch.Tensor:
        return self.materialize().to(*args, **kwargs)

    def __repr__(self) -> str:
        return f"PhiloxDraw(shape={self.shape}, items={len(self.seeds)}, dtype={self.dtype}, device={self.device})"


@dataclass
class TensorNoiseCommon[T: TensorNoiseProps | None](SkrampleTensorNoise):
    "Shape / seed / dtype / props shared by the base generators. reference: noise.py:27-55"

    shape: tuple[int, ...]
    seed: torch.Generator
    dtype: torch.dtype
    props: T

    @property
    def on_device(self) -> bool:
        return self.seed.device.type == "cuda"

    def _randn(self, shape: tuple[int, ...] | None = None) -> torch.Tensor:
        "A plain normal draw from this generator (reference-compatible helper)."
        shape = tuple(self.shape if shape is None else shape)
        if self.on_device:
            out = torch.empty(shape, dtype=self.dtype, device=self.seed.device)
            self._fill(out, self._tick())
            return out
        return torch.randn(shape, generator=self.seed, dtype=self.dtype, device=self.seed.device)

    # -- device helpers
    def _tick(self) -> int:
        """This call's block of Philox streams (``_SUBSTREAMS`` stream ids starting at the returned one).

        The generator's own offset is advanced so interleaved ``torch.randn`` calls on it stay independent, but not
        once per call: ``_RESERVED_CALLS`` calls' worth of streams are reserved at a time, and a call inside the
        reservation costs one ``get_offset()`` - the check that nobody else drew from (or re-seeded) the generator in
        between; if somebody did, a fresh reservation starts at the generator's current state.  While this source
        is the generator's only user the stream ids are exactly those of advancing once per call."""
        seed = self.seed
        held = self.__dict__.get("_skr_reserved")
        if held is not None and held[1] < held[2]:
            try:
                untouched = seed.get_offset() == held[0]
            except RuntimeError:  # a CUDA graph is being captured: generator state cannot be read (see `reserve`)
                untouched = True
            if untouched:
                tick = held[1]
                held[1] = tick + _SUBSTREAMS
                return tick
        return self.reserve(_RESERVED_CALLS, _take=True)

    def reserve(self, calls: int, _take: bool = False) -> int:
        """Reserve the Philox streams of the next ``calls`` generate() calls with one advance of the generator's offset.
        Needed before capturing draws into a CUDA graph: torch refuses to read or move a generator's offset during
        capture, so a captured region can only consume streams that were reserved beforehand."""
        seed = self.seed
        at = int(seed.get_offset())
        tick = at // 4
        after = at + 4 * _SUBSTREAMS * max(1, calls)
        seed.set_offset(after)
        first = tick + _SUBSTREAMS if _take else tick
        # [offset after the reservation, next stream id, end of the reservation, Philox key, bound get_offset]
        self.__dict__["_skr_reserved"] = [after, first, tick + _SUBSTREAMS * max(1, calls), int(seed.initial_seed()) & 0xFFFFFFFFFFFFFFFF, seed.get_offset]
        return tick

    def _key(self) -> int:
        "Philox key = the generator's seed (read once per reservation: re-seeding resets the offset, which ends it)."
        held = self.__dict__.get("_skr_reserved")
        if held is not None:
            try:
                if self.seed.get_offset() == held[0]:
                    return held[3]
            except RuntimeError:  # capturing: the reservation made before the capture stands
                return held[3]
        return int(self.seed.initial_seed()) & 0xFFFFFFFFFFFFFFFF

    def _fill(self, out: torch.Tensor, stream: int, offset: _SkrOffset | None = None, moments: torch.Tensor | None = None) -> None:
        with _DeviceGuard(out.device):
            status = _lib().skr_noise_fill(
                out.data_ptr(),
                _code(out.dtype),
                out.numel(),
                self._key(),
                stream,
                ctypes.byref(offset) if offset is not None else None,
                moments.data_ptr() if moments is not None else None,
                _stream(),
            )
        _native().check(status, "skr_noise_fill")

    def generate_into(self, out: torch.Tensor, step: Step | None) -> None:
        "Write the next draw into ``out`` (used by the batch helper to avoid a stack copy)."
        out.copy_(self.generate(step))

    @classmethod
    @abstractmethod
    def from_inputs(
        cls,
        shape: tuple[int, ...],
        seed: torch.Generator,
        props: T = None,  # type: ignore[assignment]
        dtype: torch.dtype = torch.float32,
    ) -> Self:
        "Build the generator from what is at hand during inference."
        raise NotImplementedError


# ------------------------------------------------------------------------------------------------------------


@dataclass
class Random(TensorNoiseCommon[None]):
    "Independent standard normals. reference: noise.py:58-74"

    @classmethod
    def from_inputs(cls, shape: tuple[int, ...], seed: torch.Generator, props: None = None, dtype: torch.dtype = torch.float32) -> Self:
        return cls(shape, seed, dtype, props)

    def generate(self, step: Step | None) -> torch.Tensor:
        return self._randn()

    def generate_into(self, out: torch.Tensor, step: Step | None) -> None:
        if self.on_device and out.is_cuda and out.is_contiguous():
            self._fill(out, self._tick())
        else:
            out.copy_(self.generate(step))

    def lazy(self, step: Step | None) -> "PhiloxDraw | torch.Tensor":
        "The next draw as Philox keys for in-kernel generation (CUDA generators); a tensor otherwise."
        if not self.on_device:
            return self.generate(step)
        tick = self._tick()
        return PhiloxDraw(tuple(self.shape), (self.__dict__["_skr_reserved"][3],), (tick,), self.dtype, self.seed.device)


@dataclass(frozen=True)
class OffsetProps(TensorNoiseProps):
    dims: tuple[int, ...] = (0,)
    strength: float = 0.2
    static: bool = False


@dataclass
class Offset(TensorNoiseCommon[OffsetProps]):
    "Normal noise plus one random offset per slice along ``dims``. reference: noise.py:84-113"

    @classmethod
    def from_inputs(
        cls, shape: tuple[int, ...], seed: torch.Generator, props: OffsetProps = OffsetProps(), dtype: torch.dtype = torch.float32
    ) -> Self:
        return cls(shape, seed, dtype, props)

    def __post_init__(self) -> None:
        self.static_offset: torch.Tensor | None = None
        self._static_stream: int | None = None
        if self.props.static:
            if self.on_device:
                self._static_stream = self._tick() + 1
            else:
                self.static_offset = self.offset()

    def _offset_shape(self) -> tuple[int, ...]:
        return tuple(d if n in self.props.dims else 1 for n, d in enumerate(self.shape))

    def offset(self) -> torch.Tensor:
        return self._randn(self._offset_shape()) * self.props.strength**2

    def _descriptor(self, stream: int) -> _SkrOffset:
        desc = _SkrOffset()
        desc.ndim = len(self.shape)
        for n, d in enumerate(self.shape):
            desc.shape[n] = d
            desc.keep[n] = int(n in self.props.dims)
        desc.stream = stream
        desc.scale = self.props.strength**2
        return desc

    def generate_into(self, out: torch.Tensor, step: Step | None) -> None:
        if self.on_device and out.is_cuda and out.is_contiguous() and len(self.shape) <= 8:
            base = self._tick()
            offset_stream = self._static_stream if self._static_stream is not None else base + 1
            self._fill(out, base, self._descriptor(offset_stream))
        else:
            out.copy_(self.generate(step))

    def generate(self, step: Step | None) -> torch.Tensor:
        if self.on_device and len(self.shape) <= 8:
            out = torch.empty(tuple(self.shape), dtype=self.dtype, device=self.seed.device)
            self.generate_into(out, step)
            return out
        offset = self.static_offset if self.props.static and self.static_offset is not None else self.offset()
        return self._randn() + offset


@dataclass(frozen=True)
class PyramidProps(OffsetProps):
    dims: tuple[int] | tuple[int, int] | tuple[int, int, int] = (-1, -2)
    strength: float = 0.3
    depth: int = 99
    "Maximum number of pyramid levels kept, counted from the coarsest"


@dataclass
class Pyramid(TensorNoiseCommon[PyramidProps]):
    "Multi-resolution noise: progressively coarser normal fields upsampled and summed. reference: noise.py:125-207"

    @classmethod
    def from_inputs(
        cls, shape: tuple[int, ...], seed: torch.Generator, props: PyramidProps = PyramidProps(), dtype: torch.dtype = torch.float32
    ) -> Self:
        return cls(shape, seed, dtype, props)

    def __post_init__(self) -> None:
        self._static_pyramid: torch.Tensor | None = None
        self._static_plan: tuple[int, list[tuple[int, ...]]] | None = None
        if self.props.static:
            if self.on_device:
                tick = self._tick()
                self._static_plan = (tick, self._level_shapes(lambda level: _host_uniform(self._key(), tick, level)))
            else:
                self._static_pyramid = self.pyramid()

    def _mask(self) -> list[bool]:
        rank = len(self.shape)
        axes = [rank + d if d < 0 else d for d in self.props.dims]
        return [n in axes for n in range(rank)]

    def _level_shapes(self, uniform: Any) -> list[tuple[int, ...]]:
        "Shapes of the pyramid levels: every level shrinks the running shape by r**level, r ~ U[2, 4)."
        mask = self._mask()
        running = list(self.shape)
        shapes: list[tuple[int, ...]] = []
        for level in range(99):
            ratio = uniform(level) * 2 + 2
            running = [max(1, int(s / (ratio**level))) if m else s for m, s in zip(mask, running)]
            shapes.append(tuple(running))
            if any(s <= 1 for m, s in zip(mask, running) if m):
                break
        return shapes

    def _kept(self, count: int) -> int:
        "Index of the first level that is kept (``depth`` counts from the coarsest)."
        top = count - 1
        return min(top, max(0, top - self.props.depth))

    # -- reference-exact path (CPU generator): same draws in the same order
    def pyramid(self) -> torch.Tensor:
        "Just the summed pyramid component."
        if self.on_device:
            raise RuntimeError("Pyramid.pyramid() is only materialised for CPU generators; use generate() on CUDA")
        mask = self._mask()
        target = tuple(s for m, s in zip(mask, self.shape) if m)
        mode = ["linear", "bilinear", "bicubic"][len(target) - 1]
        order = [n for n, m in enumerate(mask) if not m] + [n for n, m in enumerate(mask) if m]
        inverse = [order.index(n) for n in range(len(order))]
        lead = len(order) - len(target)
        device = self.seed.device

        levels: list[torch.Tensor] = []
        running = list(self.shape)
        for level in range(99):
            ratio = torch.rand([1], dtype=self.dtype, device=device, generator=self.seed).item() * 2 + 2
            running = [max(1, int(s / (ratio**level))) if m else s for m, s in zip(mask, running)]
            field = torch.randn(running, dtype=self.dtype, device=device, generator=self.seed)
            moved = field.permute(order)
            lead_shape = moved.shape[:lead]
            stacked = moved.reshape(-1, 1, *moved.shape[lead:])  # every leading slice is an independent image
            grown = torch.nn.functional.interpolate(stacked, target, mode=mode)
            levels.append(grown.reshape(*lead_shape, *target).permute(inverse).reshape(self.shape) * self.props.strength**level)
            if any(s <= 1 for m, s in zip(mask, running) if m):
                break
        total = torch.zeros(self.shape, dtype=self.dtype, device=device)
        return total + sum(levels[self._kept(len(levels)) :])

    # -- device path
    _use_grids = True  # False: draw every interpolation corner inside the kernel (no scratch memory, ~20x slower)

    def _generate_device(self, out: torch.Tensor) -> None:
        rank = len(self.shape)
        if rank > 8:
            raise ValueError("Pyramid on CUDA supports unit shapes of rank <= 8")
        mask = self._mask()
        if not 1 <= sum(mask) <= 2:
            raise NotImplementedError("Pyramid resizes 1 or 2 axes (3 is broken in the reference too: bicubic on 5-D)")
        base = self._tick()
        if self._static_plan is not None:
            level_tick, shapes = self._static_plan
        else:
            level_tick = base
            shapes = self._level_shapes(lambda level: _host_uniform(self._key(), base, level))
        if len(shapes) > 16:
            shapes = shapes[:16]
        first = self._kept(len(shapes))

        desc = _SkrPyramid()
        desc.ndim = rank
        desc.n_levels = len(shapes)
        desc.seed = self._key()
        desc.base_stream = base
        for n, d in enumerate(self.shape):
            desc.shape[n] = d
            desc.masked[n] = int(mask[n])
        # Three launches: the coarse level grids (small: each is the unit shape shrunk by the level's ratio along the
        # resized axes) are drawn by one launch into scratch memory; the composition kernel draws the base and the
        # unit-sized level 0 in registers, interpolates the coarse grids, writes the unnormalised field and sums it; the
        # scale pass divides by the std.  (Drawing every interpolation corner inside the kernel instead - no scratch at
        # all, `_use_grids = False` - costs ~32 Philox blocks per element.)
        numel = math.prod(self.shape)
        align = lambda n: (n + 3) & ~3  # noqa: E731 - keep every grid 16-byte aligned for vector stores
        coarse = [math.prod(shape) if level >= first and tuple(shape) != tuple(self.shape) and self._use_grids else 0 for level, shape in enumerate(shapes)]
        # ... plus room to stretch each coarse level to the unit's width (rows of the level x unit width) when the
        # resized axes are the trailing ones: the composition then interpolates whole rows
        width = self.shape[-1]
        widened = [n // shape[-1] * width if n and mask[-1] else 0 for n, shape in zip(coarse, shapes)]
        levels_room = sum(align(n) for n in coarse) + sum(align(n) for n in widened)
        scratch = torch.empty(align(numel) + levels_room, dtype=torch.float32, device=out.device) if self._use_grids else None
        with _DeviceGuard(out.device):
            for level, shape in enumerate(shapes):
                slot = desc.levels[level]
                slot.stream = level_tick + 2 + level
                extents = [s for m, s in zip(mask, shape) if m]
                slot.extent[0] = extents[0]
                slot.extent[1] = extents[1] if len(extents) > 1 else 1
                slot.weight = self.props.strength**level if level >= first else 0.0
            if scratch is not None:
                desc.scratch = scratch.data_ptr()
                desc.levels_scratch = scratch.data_ptr() + 4 * align(numel)
                desc.levels_scratch_floats = levels_room
            moments = torch.zeros(_MOMENTS, dtype=torch.float64, device=out.device)
            status = _lib().skr_noise_pyramid(out.data_ptr(), _code(out.dtype), ctypes.byref(desc), moments.data_ptr(), _stream())
        _native().check(status, "skr_noise_pyramid")

    def generate_into(self, out: torch.Tensor, step: Step | None) -> None:
        if self.on_device and out.is_cuda and out.is_contiguous():
            self._generate_device(out)
        else:
            out.copy_(self.generate(step))

    def generate(self, step: Step | None) -> torch.Tensor:
        if self.on_device:
            out = torch.empty(tuple(self.shape), dtype=self.dtype, device=self.seed.device)
            self._generate_device(out)
            return out
        component = self._static_pyramid if self.props.static and self._static_pyramid is not None else None
        noise = self._randn() + (component if component is not None else self.pyramid())
        return noise / noise.std()


# -- host side of the Brownian interval (CPU generators without torchsde): the walk of skr_noise_brownian and the
#    kernel's arithmetic restated with torch CPU ops, so a seed names the same path on the host and on the device.

_BROWNIAN_TREE, _BROWNIAN_LEAF = 1 << 63, 1 << 62
_warned_no_torchsde = False
_M32 = 0xFFFFFFFF


def _host_philox_normals(seed: int, stream: int, groups: int) -> torch.Tensor:
    "[groups, 4] float32 standard normals of Philox stream (seed, stream): counter = (group, stream), as philox.cuh."
    index = torch.arange(groups, dtype=torch.int64)
    c0, c1 = index & _M32, (index >> 32) & _M32
    c2 = torch.full_like(index, stream & _M32)
    c3 = torch.full_like(index, (stream >> 32) & _M32)
    k0, k1 = seed & _M32, (seed >> 32) & _M32
    for _ in range(10):
        # 32x32-bit products wrap in int64, which keeps their low 64 bits: the halves are recovered with masks
        p0, p1 = c0 * 0xD2511F53, c2 * 0xCD9E8D57
        c0, c1, c2, c3 = ((p1 >> 32) & _M32) ^ c1 ^ k0, p1 & _M32, ((p0 >> 32) & _M32) ^ c3 ^ k1, p0 & _M32
        k0, k1 = (k0 + 0x9E3779B9) & _M32, (k1 + 0xBB67AE85) & _M32
    u = torch.stack([c0, c1, c2, c3], dim=1).to(torch.float32) * 2.0**-32 + 2.0**-33  # (0, 1]
    radius = torch.sqrt(-2.0 * torch.log(u[:, 0::2]))
    angle = (2.0 * math.pi) * u[:, 1::2]
    z = torch.stack([radius[:, 0] * torch.sin(angle[:, 0]), radius[:, 0] * torch.cos(angle[:, 0]),
                     radius[:, 1] * torch.sin(angle[:, 1]), radius[:, 1] * torch.cos(angle[:, 1])], dim=1)
    return z


def _host_brownian_increment(seed: int, t0: float, t1: float, depth: int, numel: int) -> torch.Tensor:
    "(W(t1) - W(t0)) of the seed's path as float32 [numel]: the bridge-tree walk of skr_noise_brownian on the host."
    groups = (numel + 3) // 4

    def draw(stream: int) -> torch.Tensor:
        return _host_philox_normals(seed, stream, groups)

    def descend(t: float, left: float, right: float, node: int, level: int, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
        for _ in range(level, depth):
            mid = 0.5 * (left + right)
            value = 0.5 * (lo + hi) + (0.5 * math.sqrt(right - left)) * draw(_BROWNIAN_TREE | node)
            if t >= mid:
                left, lo, node = mid, value, 2 * node + 1
            else:
                right, hi, node = mid, value, 2 * node
        width = right - left
        spread = math.sqrt((t - left) * (right - t) / width)
        return lo + ((t - left) / width) * (hi - lo) + spread * draw(_BROWNIAN_TREE | _BROWNIAN_LEAF | node)

    left, right, node, level = 0.0, 1.0, 1, 0
    increment = draw(_BROWNIAN_TREE)  # W(1) - W(0)
    while level < depth:
        mid = 0.5 * (left + right)
        if (t0 >= mid) != (t1 >= mid):
            break
        half = (0.5 * math.sqrt(right - left)) * draw(_BROWNIAN_TREE | node)
        if t0 >= mid:
            increment, left, node = 0.5 * increment - half, mid, 2 * node + 1
        else:
            increment, right, node = 0.5 * increment + half, mid, 2 * node
        level += 1
    if level == depth:  # both times inside one leaf
        width = right - left
        spread = math.sqrt((t1 - left) * (right - t1) / width) - math.sqrt((t0 - left) * (right - t0) / width)
        result = ((t1 - t0) / width) * increment + spread * draw(_BROWNIAN_TREE | _BROWNIAN_LEAF | node)
    else:  # the node that separates them: both sides relative to W(mid)
        mid = 0.5 * (left + right)
        half = (0.5 * math.sqrt(right - left)) * draw(_BROWNIAN_TREE | node)
        zero = torch.zeros_like(increment)
        before = descend(t0, left, mid, 2 * node, level + 1, -(0.5 * increment + half), zero)
        after = descend(t1, mid, right, 2 * node + 1, level + 1, zero, 0.5 * increment - half)
        result = after - before
    return result.reshape(-1)[:numel]


@dataclass(frozen=True)
class BrownianProps(TensorNoiseProps):
    max_steps: int = 10_000
    """Target resolution of the Brownian tree: steps shorter than 1/max_steps fall inside one leaf of the tree
    and lose exactness (the reference's tolerance caveat, noise.py:211-216)."""


@dataclass
class Brownian(TensorNoiseCommon[BrownianProps]):
    """Noise that is a deterministic function of the ``Step``: the increment of one Brownian path per seed.
    reference: noise.py:210-252 (which delegates to ``torchsde.BrownianInterval``).

    With a CUDA generator the increment comes from ``skr_noise_brownian``: a stateless Philox bridge tree over
    normalised time, ``ceil(log2(10 * max_steps))`` levels deep (the reference's ``tol``), evaluated per element in
    one launch - no tree object, no cache, no host state.  The same step gives the same tensor on every call,
    adjoining steps add up to the increment of their union and disjoint steps are independent, which is the
    contract the reference gets from torchsde.  The VALUES are this library's own: torchsde is a third-party
    module that is not part of the reference tree, so there is nothing to pin them against (parity statistical).
    With a CPU generator the reference's torchsde tree is used as is when the module is installed (the reference's
    values); without it the same bridge tree is evaluated with torch CPU ops, so a seed names one path on the host and on
    the device (equal up to float32 rounding of log / sin / cos) instead of raising ImportError.
    """

    def __post_init__(self) -> None:
        if self.on_device:
            self._tree = None
            self._depth = max(1, math.ceil(math.log2(self.props.max_steps * 10)))
            if self._depth > 40:
                raise ValueError(f"BrownianProps.max_steps={self.props.max_steps} is beyond the tree depth the kernel walks (40 levels)")
            return
        self._depth = max(1, math.ceil(math.log2(self.props.max_steps * 10)))
        try:
            import torchsde
        except ImportError:
            # The reference raises ImportError here (noise.py:223).  This library instead evaluates its own bridge tree
            # on the host - the SAME path the CUDA kernel walks for this seed - and says so once: the values are not
            # torchsde's (set SKRAMPLE_B200_REQUIRE_TORCHSDE=1 to get the reference's ImportError instead).
            import os
            import warnings

            if os.environ.get("SKRAMPLE_B200_REQUIRE_TORCHSDE"):
                raise
            global _warned_no_torchsde
            if not _warned_no_torchsde:
                _warned_no_torchsde = True
                warnings.warn(
                    "skrample_b200: torchsde is not installed; Brownian noise on a CPU generator comes from this library's own "
                    "Philox bridge tree (the path its CUDA kernel walks for the same seed), not from torchsde.BrownianInterval",
                    RuntimeWarning,
                    stacklevel=3,
                )
            self._tree = None  # host evaluation of the library's own construction
            return

        self._tree = torchsde.BrownianInterval(
            t0=0,
            t1=1,
            size=self.shape,
            entropy=self.seed.initial_seed(),
            dtype=self.dtype,
            device=self.seed.device,
            halfway_tree=True,
            tol=1 / (self.props.max_steps * 10),
            pool_size=2**6,
            cache_size=round(math.log2(self.props.max_steps * 10) * 1.3),
        )

    def _increment(self, out: torch.Tensor, step: Step) -> None:
        scale = 1 / math.sqrt(step.distance())
        with _DeviceGuard(out.device):
            status = _lib().skr_noise_brownian(
                out.data_ptr(), _code(out.dtype), out.numel(), self._key(), step.time_from, step.time_to, self._depth, scale, _stream()
            )
        _native().check(status, "skr_noise_brownian")

    def generate(self, step: Step | None) -> torch.Tensor:
        if not step:
            return self._randn()
        step = step.normal().clamp()
        if self._tree is not None:
            return self._tree(*step) / math.sqrt(step.distance())
        if not self.on_device:
            scale = 1 / math.sqrt(step.distance())
            if not (0.0 <= step.time_from < step.time_to <= 1.0):
                raise RuntimeError("Brownian: the step must lie inside 0..1 after clamping")
            walked = _host_brownian_increment(self._key(), step.time_from, step.time_to, self._depth, math.prod(self.shape))
            return (walked * scale).to(self.dtype).reshape(tuple(self.shape))
        out = torch.empty(tuple(self.shape), dtype=self.dtype, device=self.seed.device)
        self._increment(out, step)
        return out

    def generate_into(self, out: torch.Tensor, step: Step | None) -> None:
        if step and self._tree is None and self.on_device and out.is_cuda and out.is_contiguous() and out.numel() == math.prod(self.shape):
            self._increment(out, step.normal().clamp())
        else:
            out.copy_(self.generate(step))

    @classmethod
    def from_inputs(
        cls, shape: tuple[int, ...], seed: torch.Generator, props: BrownianProps = BrownianProps(), dtype: torch.dtype = torch.float32
    ) -> Self:
        return cls(shape=shape, seed=seed, dtype=dtype, props=props)


@dataclass(frozen=True)
class ColoredProps(TensorNoiseProps):
    energy: float | None = None
    "Target standard deviation; None keeps the white noise's own"
    color_start: float = 1 / 4
    "Power-law exponent at the start of sampling (step None); > 0 is redder"
    color_end: float = -2
    "Exponent at the end of sampling"
    color_curve: float = 2
    "Curvature of the exponent ramp (like FlowShift)"


@dataclass
class Colored(TensorNoiseCommon[ColoredProps]):
    "Power-law coloured noise whose exponent follows the step. reference: noise.py:274-436"

    @staticmethod
    def _radial_freq_grid(shape: torch.Size, device: torch.device) -> torch.Tensor:
        "Normalised radial frequency of every rfftn bin (host/torch form, used for CPU tensors). reference: noise.py:285-335"
        rank = len(shape)
        axes = []
        for i, dim in enumerate(shape):
            if i == rank - 1:
                axes.append(torch.arange(dim // 2 + 1, device=device) / dim)
            else:
                axes.append(torch.fft.fftfreq(dim, d=1.0, device=device).abs())
        radius = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).norm(p=2, dim=-1)
        top = radius.max()
        return radius / top if top > 0 else radius

    @staticmethod
    def colorize_noise(white: torch.Tensor, exponent: float = 0.0, energy: float | None = None) -> torch.Tensor:
        """Shape ``white`` to a ``f^-exponent`` power spectrum and restore its deviation (or ``energy``).

        reference: noise.py:338-405.  CUDA tensors: moments, spectral shaping and the final rescale are custom
        kernels around cuFFT, with no host synchronisation; CPU tensors use torch ops.
        """
        if white.is_cuda:
            return Colored._colorize_device(white, exponent, energy, None)
        wstd = white.std()
        if exponent == 0.0:
            return white if energy is None or wstd < 1e-8 else white * (energy / wstd)
        w = white.squeeze()
        if w.dtype not in (torch.float32, torch.float64):
            w = w.to(torch.float32)
        spectrum = torch.fft.rfftn(w)
        radius = Colored._radial_freq_grid(w.shape, w.device)
        mean_extent = sum(w.shape) / len(w.shape) if w.shape else 1.0
        floor = 0.5 / max(mean_extent, 4.0)
        shaped = spectrum * (torch.clamp(radius, min=floor) ** (-exponent / 2.0))
        colored = torch.fft.irfftn(shaped, s=w.shape)
        cstd = colored.std()
        if cstd > 1e-8:
            colored *= wstd / cstd if energy is None else energy / cstd
        return colored.view(white.shape).to(dtype=white.dtype)

    @staticmethod
    def _colorize_device(
        white: torch.Tensor, exponent: float, energy: float | None, white_moments: torch.Tensor | None, out_dtype: torch.dtype | None = None
    ) -> torch.Tensor:
        lib = _lib()
        native = _native()
        white = white.contiguous()  # raw pointers below: a transposed / expanded view would be read in storage order
        n = white.numel()
        out_dtype = out_dtype or white.dtype
        with _DeviceGuard(white.device):
            stream = _stream()
            if white_moments is None:
                white_moments = torch.zeros(_MOMENTS, dtype=torch.float64, device=white.device)
                native.check(lib.skr_noise_moments(white.data_ptr(), _code(white.dtype), n, white_moments.data_ptr(), stream), "skr_noise_moments")
            if exponent == 0.0:
                if energy is None:
                    return white if white.dtype == out_dtype else white.to(out_dtype)
                out = torch.empty(white.shape, dtype=out_dtype, device=white.device)
                native.check(
                    lib.skr_noise_scale(white.data_ptr(), _code(white.dtype), out.data_ptr(), _code(out.dtype), n, float(energy), None, 0, white_moments.data_ptr(), n, 1e-8, stream),
                    "skr_noise_scale",
                )
                return out
            w = white.squeeze()
            if w.dtype not in (torch.float32, torch.float64):
                w = w.to(torch.float32)
            spectrum = torch.fft.rfftn(w).contiguous()
            dims = (ctypes.c_int64 * max(1, w.dim()))(*(w.shape if w.dim() else (1,)))
            native.check(
                lib.skr_colored_shape(spectrum.data_ptr(), _code(w.dtype), dims, max(1, w.dim()), float(exponent), stream),
                "skr_colored_shape",
            )
            # norm="forward": no 1/N pass on the inverse; the field is renormalised by its own std just below, so a
            # constant factor only has to be carried into the degenerate-std threshold
            colored = torch.fft.irfftn(spectrum, s=w.shape, norm="forward").contiguous()
            colored_moments = torch.zeros(_MOMENTS, dtype=torch.float64, device=white.device)
            native.check(lib.skr_noise_moments(colored.data_ptr(), _code(colored.dtype), n, colored_moments.data_ptr(), stream), "skr_noise_moments")
            out = torch.empty(white.shape, dtype=out_dtype, device=white.device)
            native.check(
                lib.skr_noise_scale(
                    colored.data_ptr(),
                    _code(colored.dtype),
                    out.data_ptr(),
                    _code(out.dtype),
                    n,
                    1.0 if energy is None else float(energy),
                    white_moments.data_ptr() if energy is None else None,
                    n,
                    colored_moments.data_ptr(),
                    n,
                    1e-8 * n,
                    stream,
                ),
                "skr_noise_scale",
            )
        return out

    def exponent(self, step: Step | None) -> float:
        "Power-law exponent for this step. reference: noise.py:410-420"
        if step is None:
            return self.props.color_start
        if self.props.color_curve == math.inf:
            return self.props.color_end
        t = step.normal().clamp().time_to
        shift = rescale_positive(-self.props.color_curve)  # negative: steps ascend like alpha, not sigma
        t = shift / (shift + (divf(1, t) - 1))
        return (1 - t) * self.props.color_start + t * self.props.color_end

    def generate(self, step: Step | None) -> torch.Tensor:
        exponent = self.exponent(step)
        if self.on_device:
            work_dtype = self.dtype if self.dtype in (torch.float32, torch.float64) else torch.float32
            white = torch.empty(tuple(self.shape), dtype=work_dtype, device=self.seed.device)
            moments = torch.zeros(_MOMENTS, dtype=torch.float64, device=white.device)
            self._fill(white, self._tick(), moments=moments)
            return self._colorize_device(white, exponent, self.props.energy, moments, self.dtype)
        return self.colorize_noise(self._randn(), exponent=exponent, energy=self.props.energy)

    @classmethod
    def from_inputs(
        cls, shape: tuple[int, ...], seed: torch.Generator, props: ColoredProps = ColoredProps(), dtype: torch.dtype = torch.float32
    ) -> Self:
        return cls(shape=shape, seed=seed, dtype=dtype, props=props)


@dataclass
class BatchTensorNoise[T: TensorNoiseProps | None](SkrampleTensorNoise):
    """One generator per batch item so every item keeps its own seed. reference: noise.py:438-466

    Items write straight into their slice of one ``[batch, *unit]`` tensor (no ``torch.stack`` copy) when all
    generators live on the same CUDA device."""

    generators: list[TensorNoiseCommon[T]]

    def generate(self, step: Step | None) -> torch.Tensor:
        drawn = self.lazy(step, _fallback=False)
        if drawn is not None:
            return drawn.materialize()  # plain Random on one CUDA device: one launch for the whole batch
        first = self.generators[0]
        same_place = all(g.on_device and g.seed.device == first.seed.device and g.dtype == first.dtype and g.shape == first.shape for g in self.generators)
        if same_place:
            out = torch.empty((len(self.generators), *first.shape), dtype=first.dtype, device=first.seed.device)
            if step and self._brownian_batch(out, step):
                return out
            for row, generator in zip(out, self.generators, strict=True):
                generator.generate_into(row, step)
            return out
        return torch.stack([g.generate(step) for g in self.generators])

    def generate_into(self, out: torch.Tensor, step: Step | None) -> None:
        """Write the next batch of noise into ``out`` (``[batch, *unit]``, any floating dtype): the values of
        ``generate(step).to(out.dtype)`` without the intermediate tensor when the batch is plain ``Random`` on
        ``out``'s device."""
        if step and out.is_cuda and out.is_contiguous() and tuple(out.shape) == (len(self.generators), *self.generators[0].shape):
            if self._brownian_batch(out, step):
                return
        if out.is_cuda and out.is_contiguous() and self._uniform_random() and _same_device(self.generators[0].seed.device, out.device):
            drawn = self.lazy(step, _fallback=False)
            if drawn is not None and tuple(out.shape) == drawn.shape:
                drawn.materialize_into(out)
                return
            out.copy_(drawn.materialize() if drawn is not None else self.generate(step))
            return
        out.copy_(self.generate(step))

    def _brownian_batch(self, out: torch.Tensor, step: Step) -> bool:
        """Every item a device ``Brownian`` of one shape and tree depth on ``out``'s device: one launch writes the
        whole batch, item i from the path of its own seed (the values of the per-item calls)."""
        first = self.generators[0]
        count = len(self.generators)
        if count > 32 or not all(
            type(g) is Brownian and g._tree is None and g.on_device and _same_device(g.seed.device, out.device) and g.shape == first.shape and g._depth == first._depth
            for g in self.generators
        ):
            return False
        if any(g.dtype != out.dtype and g.dtype not in (torch.float32, torch.float64) for g in self.generators):
            return False  # a 16-bit generator read into a wider tensor keeps its own rounding: item by item
        step = step.normal().clamp()
        scale = 1 / math.sqrt(step.distance())
        seeds = (ctypes.c_uint64 * count)(*(g._key() for g in self.generators))
        with _DeviceGuard(out.device):
            status = _lib().skr_noise_brownian_batch(
                out.data_ptr(), _code(out.dtype), seeds, count, out.numel() // count, step.time_from, step.time_to, first._depth, scale, _stream()
            )
        _native().check(status, "skr_noise_brownian_batch")
        return True

    def reserve(self, calls: int) -> None:
        "Reserve the Philox streams of the next ``calls`` draws of every item (required before CUDA-graph capture)."
        for g in self.generators:
            if g.on_device:
                g.reserve(calls)

    def auto(self, step: Step | None) -> "PhiloxDraw | torch.Tensor":
        """The next batch of noise in the form a sampler makes the most of: Philox keys (a ``PhiloxDraw``) when every
        item is a plain ``Random`` (or a non-static ``Offset`` along leading axes) on one CUDA device - the step then draws the normals inside its own kernel, or has
        them written by the fill kernel first where that is faster for the step at hand (``Program.settle_noise``) - and
        the generated tensor otherwise.  Same values either way."""
        info = self._uniform_info()
        if info is not None and info[0] <= _native().MAX_PHILOX_ITEMS:
            return self.lazy(step)
        return self.generate(step)

    def _uniform_info(self) -> tuple | None:
        """(count, shape, numel, dtype, device, offset_inner, offset_scale) when every item is a plain ``Random`` - or
        every item an ``Offset`` with the same props, a fresh offset per draw and its kept axes leading - with one shape /
        dtype on one CUDA device, else None.  Checked once per generator list (identity of the list, its length and
        its ends)."""
        generators = self.generators
        cached = self.__dict__.get("_skr_uniform")
        if cached is not None and cached[0] is generators and cached[1] == len(generators) and cached[2] is generators[0] and cached[3] is generators[-1]:
            return cached[4]
        first = generators[0]
        kind = type(first)
        ok = kind in (Random, Offset) and all(
            type(g) is kind and g.on_device and g.seed.device == first.seed.device and g.dtype == first.dtype and g.shape == first.shape for g in generators
        )
        inner, scale = 0, 0.0
        if ok and kind is Offset:
            # the in-kernel form covers what the fill kernel's fast path covers: kept axes first, broadcast axes after
            props = first.props
            kept = [n in props.dims for n in range(len(first.shape))]
            lead = kept.index(False) if False in kept else len(kept)
            ok = not props.static and not any(kept[lead:]) and all(g.props == props for g in generators)
            inner, scale = math.prod(first.shape[lead:]), float(props.strength**2)
            if scale == 0.0:
                inner = 0
        shape = (len(generators), *first.shape)
        info = (len(generators), shape, math.prod(shape), first.dtype, first.seed.device, inner, scale) if ok else None
        self.__dict__["_skr_uniform"] = (generators, len(generators), first, generators[-1], info)
        return info

    def _uniform_random(self) -> bool:
        return self._uniform_info() is not None

    def lazy(self, step: Step | None, _fallback: bool = True) -> "PhiloxDraw | torch.Tensor | None":
        """The next batch of noise as Philox keys when every item is a plain ``Random`` (or a non-static ``Offset`` along
        leading axes) on one CUDA device; otherwise the materialised tensor."""
        info = self._uniform_info()
        if info is None:
            return self.generate(step) if _fallback else None
        streams, seeds = [], []
        try:
            # TensorNoiseCommon._tick, inlined for its common case: one get_offset() per generator (the check that nobody
            # re-seeded or drew from it since the reservation was made), everything else is list arithmetic
            for g in self.generators:
                held = g.__dict__.get("_skr_reserved")
                if held is not None and held[1] < held[2] and held[4]() == held[0]:
                    tick = held[1]
                    held[1] = tick + _SUBSTREAMS
                else:
                    tick = g._tick()
                    held = g.__dict__["_skr_reserved"]
                streams.append(tick)
                seeds.append(held[3])
        except RuntimeError:  # a CUDA graph is being captured: the reservations made beforehand stand (see `reserve`)
            streams, seeds = [], []
            for g in self.generators:
                streams.append(g._tick())
                seeds.append(g.__dict__["_skr_reserved"][3])
        return PhiloxDraw(info[1], tuple(seeds), tuple(streams), info[3], info[4], info[2], info[5], info[6])

    @classmethod
    def from_batch_inputs[U: TensorNoiseProps | None](
        cls,
        subclass: type[TensorNoiseCommon[U]],
        unit_shape: tuple[int, ...],
        seeds: list[torch.Generator],
        props: U | None = None,
        dtype: torch.dtype = torch.float32,
    ) -> "BatchTensorNoise[U]":
        "Batched ``from_inputs``: the result of ``generate`` is ``[len(seeds), *unit_shape]``."
        build = (lambda seed: subclass.from_inputs(unit_shape, seed, props, dtype)) if props is not None else (
            lambda seed: subclass.from_inputs(unit_shape, seed, dtype=dtype)
        )
        return cls([build(seed) for seed in seeds])  # type: ignore[arg-type]
