"""ctypes binding of ``libskrample_b200.so`` (the C ABI in ``include/skrample_b200.h``).

This is the only place device pointers leave PyTorch: tensors are passed as
``data_ptr()`` + dtype code, outputs are allocated here with ``torch.empty``
(caching allocator, capture-safe) and the launch goes on torch's current stream.
The library never allocates or synchronises.

There is deliberately no fallback: if the shared library is missing or fails to
load, using a CUDA tensor raises ``RuntimeError`` (build with
``python -m skrample_b200.build``).
"""

from __future__ import annotations

import ctypes
import struct
import os
import threading
from pathlib import Path
from typing import TYPE_CHECKING, Any

import torch

if TYPE_CHECKING:
    from skrample_b200.sampling.program import Program

LIB_DIR = Path(__file__).resolve().parent / "csrc"
LIB_PATH = Path(os.environ.get("SKRAMPLE_B200_LIB", LIB_DIR / "libskrample_b200.so"))

MAX_OPS = 64
MAX_INPUTS = 32
MAX_OUTPUTS = 8

F32, F64, BF16, F16 = 0, 1, 2, 3
DTYPE_CODE = {torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16, torch.float16: F16}


class SkrOp(ctypes.Structure):
    _fields_ = [
        ("code", ctypes.c_uint8),
        ("a", ctypes.c_uint8),
        ("b", ctypes.c_uint8),
        ("reserved", ctypes.c_uint8),
        ("src", ctypes.c_int16),
        ("dst", ctypes.c_int16),
        ("c", ctypes.c_double * 4),
    ]


class SkrTensor(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("reserved", ctypes.c_int32)]


MAX_PHILOX = 2
MOMENT_BLOCKS = 2048
MOMENTS_DOUBLES = 4 + 2 * MOMENT_BLOCKS
"Doubles in an accumulator of grid-wide sums (SKR_MOMENTS_DOUBLES): [0], [1] are the sums, the rest is reduction scratch."
MAX_PHILOX_ITEMS = 256


class SkrPhilox(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64 * MAX_PHILOX_ITEMS),
        ("stream", ctypes.c_uint64 * MAX_PHILOX_ITEMS),
        ("item_numel", ctypes.c_int64),
        ("n_items", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("offset_inner", ctypes.c_int64),
        ("offset_scale", ctypes.c_float),
        ("reserved", ctypes.c_int32),
    ]


class SkrProgram(ctypes.Structure):
    _fields_ = [
        ("n_ops", ctypes.c_int32),
        ("n_inputs", ctypes.c_int32),
        ("n_outputs", ctypes.c_int32),
        ("n_philox", ctypes.c_int32),
        ("ops", SkrOp * MAX_OPS),
        ("inputs", SkrTensor * MAX_INPUTS),
        ("outputs", SkrTensor * MAX_OUTPUTS),
        ("philox", SkrPhilox * MAX_PHILOX),
    ]


def _pack_draws(packed: SkrProgram, draws: list[Any]) -> None:
    "Copy lazy Philox noise descriptors (skrample_b200.pytorch.noise.PhiloxDraw) into the program."
    if len(draws) > MAX_PHILOX:
        raise RuntimeError("skrample_b200: at most two in-kernel noise draws per step program")
    packed.n_philox = len(draws)
    for slot, draw in zip(packed.philox, draws, strict=False):
        count = len(draw.seeds)
        slot.n_items = count
        slot.item_numel = draw.item_numel
        slot.seed[:count] = draw.seeds
        slot.stream[:count] = draw.streams
        slot.dtype = DTYPE_CODE.get(draw.dtype, F32)
        slot.offset_inner = draw.offset_inner
        slot.offset_scale = draw.offset_scale


EXPORTS = (
    "skr_version",
    "skr_last_error",
    "skr_launch_count",
    "skr_launch_count_kind",
    "skr_program_launch",
    "skr_program_classify",
    "skr_program_describe",
    "skr_axpby",
    "skr_error_norms",
    "skr_plan_create",
    "skr_plan_launch",
    "skr_plan_destroy",
    "skr_plan_kind",
    "skr_plan_shape",
    "skr_reload_env",
    "skr_set_arithmetic",
    "skr_get_arithmetic",
)

_lib: ctypes.CDLL | None = None
_switches_touched = False

ACCOUNT = {"on": False, "bytes": 0, "launches": 0}
"Opt-in byte accounting of launched programs (bench.py derives algorithmic bytes per step from it)."


def load() -> ctypes.CDLL:
    "Load the shared library once; raise loudly if it is not there."
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"skrample_b200: native library {LIB_PATH} not found - run `python -m skrample_b200.build` "
            "(CUDA tensors are never routed through a CPU/eager fallback)"
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    lib.skr_version.restype = ctypes.c_int
    lib.skr_last_error.restype = ctypes.c_char_p
    lib.skr_launch_count.restype = ctypes.c_int64
    lib.skr_launch_count_kind.restype = ctypes.c_int64
    lib.skr_launch_count_kind.argtypes = [ctypes.c_int32]
    lib.skr_program_launch.restype = ctypes.c_int
    lib.skr_program_launch.argtypes = [ctypes.POINTER(SkrProgram), ctypes.c_int64, ctypes.c_void_p]
    lib.skr_program_classify.restype = ctypes.c_int
    lib.skr_program_classify.argtypes = [ctypes.POINTER(SkrProgram)]
    lib.skr_program_describe.restype = ctypes.c_int
    lib.skr_program_describe.argtypes = [ctypes.POINTER(SkrProgram), ctypes.c_char_p, ctypes.c_int32]
    lib.skr_axpby.restype = ctypes.c_int
    lib.skr_axpby.argtypes = [
        ctypes.c_void_p,
        ctypes.c_void_p,
        ctypes.c_void_p,
        ctypes.c_int32,
        ctypes.c_int64,
        ctypes.c_double,
        ctypes.c_double,
        ctypes.c_int32,
        ctypes.c_void_p,
    ]
    lib.skr_plan_create.restype = ctypes.c_int
    lib.skr_plan_create.argtypes = [ctypes.POINTER(SkrProgram), ctypes.POINTER(ctypes.c_void_p)]
    lib.skr_plan_launch.restype = ctypes.c_int
    lib.skr_plan_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    lib.skr_plan_destroy.restype = None
    lib.skr_plan_destroy.argtypes = [ctypes.c_void_p]
    lib.skr_plan_kind.restype = ctypes.c_int
    lib.skr_plan_kind.argtypes = [ctypes.c_void_p]
    lib.skr_plan_shape.restype = ctypes.c_char_p
    lib.skr_plan_shape.argtypes = [ctypes.c_void_p]
    lib.skr_reload_env.restype = None
    lib.skr_set_arithmetic.restype = ctypes.c_int
    lib.skr_set_arithmetic.argtypes = [ctypes.c_int32]
    lib.skr_get_arithmetic.restype = ctypes.c_int
    lib.skr_error_norms.restype = ctypes.c_int
    lib.skr_error_norms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    _lib = lib
    return lib


def raw_stream(device_index: int | None = None) -> int:
    "cudaStream_t of torch's current stream as an integer (no Stream object construction on the hot path)."
    if device_index is None:
        device_index = torch._C._cuda_getDevice()
    return torch._C._cuda_getCurrentRawStream(device_index)


def check(status: int, what: str) -> None:
    if status != 0:
        message = load().skr_last_error().decode(errors="replace")
        raise RuntimeError(f"skrample_b200: {what} failed ({status}): {message}")


def launch_count() -> int:
    return int(load().skr_launch_count())


def launch_count_kind(kind: int) -> int:
    "kind 0 = structured block kernel, 1 = interpreter"
    return int(load().skr_launch_count_kind(kind))


def promoted_dtype(tensors: list[torch.Tensor]) -> torch.dtype:
    "torch's promotion rule restricted to the four storage types."
    kinds = {t.dtype for t in tensors}
    if torch.float64 in kinds:
        return torch.float64
    if torch.float32 in kinds or len(kinds) > 1:
        return torch.float32
    return next(iter(kinds))


def compute_dtype(tensors: list[torch.Tensor]) -> torch.dtype:
    return torch.float64 if any(t.dtype == torch.float64 for t in tensors) else torch.float32


def pack_program(program: "Program", inputs: list[torch.Tensor], outputs: list[torch.Tensor]) -> SkrProgram:
    if len(program.ops) > MAX_OPS or len(inputs) > MAX_INPUTS or len(outputs) > MAX_OUTPUTS:
        raise RuntimeError(
            f"skrample_b200: step program too large ({len(program.ops)} ops, {len(inputs)} inputs, {len(outputs)} outputs)"
        )
    packed = SkrProgram()
    packed.n_ops = len(program.ops)
    packed.n_inputs = len(inputs)
    packed.n_outputs = len(outputs)
    for slot, op in zip(packed.ops, program.ops, strict=False):
        slot.code = op.code
        slot.a = op.a
        slot.b = op.b
        slot.src = op.src
        slot.dst = op.dst
        for j, value in enumerate(op.c):
            slot.c[j] = value
    for slot, tensor in zip(packed.inputs, inputs, strict=False):
        slot.ptr = tensor.data_ptr()
        slot.dtype = DTYPE_CODE[tensor.dtype]
    for slot, tensor in zip(packed.outputs, outputs, strict=False):
        slot.ptr = tensor.data_ptr()
        slot.dtype = DTYPE_CODE[tensor.dtype]
    _pack_draws(packed, getattr(program, "philox", []))
    return packed


class _Compiled(threading.local):
    def __init__(self) -> None:
        self.by_ops: dict[tuple, CompiledProgram] = {}


_compiled = _Compiled()
_MAX_COMPILED = 1024


def launch_program(program: "Program") -> list[Any]:
    """Run a step program on CUDA tensors: allocate outputs, one kernel launch, return outputs.

    The packed op table of a program is remembered by its ops (frozen, hashable): samplers that are driven step by
    step without a step plan (RK wrappers, model conversions) re-emit equal programs, and filling the ctypes struct
    field by field costs more than the launch."""
    inputs = [t if t.is_contiguous() else t.contiguous() for t in program.inputs]
    if all(t.dtype in DTYPE_CODE for t in inputs):
        key = (tuple(program.ops), tuple(program.outputs), len(inputs))
        compiled = _compiled.by_ops.get(key)
        if compiled is None:
            if len(_compiled.by_ops) >= _MAX_COMPILED:
                _compiled.by_ops.clear()
            compiled = _compiled.by_ops[key] = CompiledProgram(program)
        outs = launch_compiled(compiled, inputs, list(getattr(program, "philox", ())) or None)
        if outs is not None:
            return outs
    lib = load()
    first = inputs[0]
    default_dtype = promoted_dtype(inputs)
    outputs: list[torch.Tensor] = []
    for want in program.outputs:
        if want is None:
            dtype = default_dtype
        elif isinstance(want, str):  # "compute"
            dtype = compute_dtype(inputs)
        else:
            dtype = want
        outputs.append(torch.empty(first.shape, dtype=dtype, device=first.device))

    if ACCOUNT["on"]:
        ACCOUNT["launches"] += 1
        ACCOUNT["bytes"] += sum(t.numel() * t.element_size() for t in inputs) + sum(t.numel() * t.element_size() for t in outputs)
    packed = pack_program(program, inputs, outputs)
    device = first.device
    if torch._C._cuda_getDevice() != device.index:
        with torch.cuda.device(device):
            status = lib.skr_program_launch(ctypes.byref(packed), first.numel(), raw_stream(device.index))
    else:
        status = lib.skr_program_launch(ctypes.byref(packed), first.numel(), raw_stream(device.index))
    check(status, "skr_program_launch")
    return outputs


class _NativePlan:
    """One ``skr_plan`` (the step parsed and its kernel chosen, for one combination of input dtypes) plus the
    reusable pointer table its launches are bound through."""

    __slots__ = ("handle", "n_inputs", "out_dtypes", "packer", "table")

    def __init__(self, handle: int, n_inputs: int, out_dtypes: tuple) -> None:
        self.handle = handle
        self.n_inputs = n_inputs
        self.out_dtypes = out_dtypes
        count = n_inputs + len(out_dtypes)
        self.table = (ctypes.c_uint64 * max(1, count))()
        self.packer = struct.Struct(f"<{count}Q")

    def __del__(self) -> None:
        lib, handle = _lib, self.handle
        if lib is not None and handle:
            self.handle = 0
            lib.skr_plan_destroy(handle)

    @property
    def kind(self) -> int:
        return int(load().skr_plan_kind(self.handle))

    @property
    def shape_name(self) -> str:
        return load().skr_plan_shape(self.handle).decode()


class CompiledProgram:
    """A step program whose op table is final.  Per combination of input dtypes it owns one native plan
    (``skr_plan_create``: validation, structure recognition and kernel selection happen once); a launch only binds
    tensor addresses (``skr_plan_launch``)."""

    __slots__ = ("fast", "n_inputs", "n_philox", "out_specs", "packed", "plans")

    def __init__(self, program: "Program") -> None:
        if len(program.ops) > MAX_OPS or len(program.inputs) > MAX_INPUTS or len(program.outputs) > MAX_OUTPUTS:
            raise RuntimeError("skrample_b200: step program too large")
        packed = SkrProgram()
        packed.n_ops = len(program.ops)
        packed.n_inputs = len(program.inputs)
        packed.n_outputs = len(program.outputs)
        packed.n_philox = len(getattr(program, "philox", ()))
        for slot, op in zip(packed.ops, program.ops, strict=False):
            slot.code, slot.a, slot.b, slot.src, slot.dst = op.code, op.a, op.b, op.src, op.dst
            for j, value in enumerate(op.c):
                slot.c[j] = value
        self.packed = packed
        self.n_inputs = len(program.inputs)
        self.n_philox = packed.n_philox
        self.out_specs = tuple(program.outputs)
        self.plans: dict[tuple, _NativePlan | None] = {}
        self.fast: dict[int, tuple] = {}  # the same plans keyed the way _fast.launch looks them up

    def specialise(self, signature: tuple) -> "_NativePlan | None":
        "The native plan for inputs of these dtypes (None: a dtype the kernels do not take)."
        codes = [DTYPE_CODE.get(d) for d in signature]
        if any(code is None for code in codes):
            self.plans[signature] = None
            return None
        any64 = torch.float64 in signature
        default = torch.float64 if any64 else (torch.float32 if torch.float32 in signature or len(set(signature)) > 1 else signature[0])
        compute = torch.float64 if any64 else torch.float32
        out_dtypes = tuple(default if want is None else (compute if want.__class__ is str else want) for want in self.out_specs)
        packed = self.packed
        for slot, code in zip(packed.inputs, codes, strict=False):
            slot.dtype = code
        for slot, dtype in zip(packed.outputs, out_dtypes, strict=False):
            slot.dtype = DTYPE_CODE[dtype]
        handle = ctypes.c_void_p()
        check(load().skr_plan_create(ctypes.byref(packed), ctypes.byref(handle)), "skr_plan_create")
        plan = self.plans[signature] = _NativePlan(handle.value, self.n_inputs, out_dtypes)
        packed_signature = 0
        for position, code in enumerate(codes):
            packed_signature |= code << (2 * position)
        self.fast[packed_signature] = (plan.handle, bytes(DTYPE_CODE[d] for d in out_dtypes))
        return plan


_ALLOWED = frozenset(DTYPE_CODE)
_MISSING = object()
_fast: Any = _MISSING
_draw_keys = threading.local()


def _pack_draw_table(draws: list[Any]) -> Any:
    "Philox key tables of the lazy noise draws of one launch (reused per thread; the C call copies them)."
    table = getattr(_draw_keys, "table", None)
    if table is None:
        table = _draw_keys.table = (SkrPhilox * MAX_PHILOX)()
    for slot, draw in zip(table, draws, strict=False):
        count = len(draw.seeds)
        slot.n_items = count
        slot.item_numel = draw.item_numel
        slot.seed[:count] = draw.seeds
        slot.stream[:count] = draw.streams
        slot.dtype = DTYPE_CODE.get(draw.dtype, F32)
        slot.offset_inner = draw.offset_inner
        slot.offset_scale = draw.offset_scale
    return table


_DTYPE_OF_CODE = {code: dtype for dtype, code in DTYPE_CODE.items()}


def _fast_module() -> Any:
    "The C++ hit path (``_fast.so``, built by skrample_b200.build), bound to this library's skr_plan_launch; or None."
    global _fast
    if _fast is _MISSING:
        try:
            if os.environ.get("SKRAMPLE_B200_NO_FAST"):
                raise ImportError("disabled")
            from skrample_b200 import _fast as module

            module.bind(ctypes.cast(load().skr_plan_launch, ctypes.c_void_p).value)
            _fast = module
        except ImportError:
            _fast = None
    return _fast


def launch_compiled(compiled: CompiledProgram, inputs: list[Any], draws: list[Any] | None = None) -> list[Any] | None:
    "Bind tensors (and lazy noise draws) to a compiled program and launch it.  None when they do not qualify."
    first = inputs[0]
    if not first.is_cuda:
        return None
    fast = _fast if _fast is not _MISSING else _fast_module()
    if fast is not None and inputs.__class__ is list:
        keys = 0
        if compiled.n_philox:
            if not draws or len(draws) != compiled.n_philox:
                return None
            numel, device = first.numel(), first.device
            for d in draws:
                if d.numel != numel or d.device != device:
                    return None
            keys = ctypes.addressof(_pack_draw_table(draws))
        elif draws:
            return None
        account = ACCOUNT["on"]
        got = fast.launch(compiled.fast, inputs, keys, account)
        if got.__class__ is int:  # no plan yet for this combination of input dtypes
            if compiled.specialise(tuple(t.dtype for t in inputs)) is None:
                return None
            got = fast.launch(compiled.fast, inputs, keys, account)
        if got.__class__ is list or got is None:
            return got
        if len(got) == 1:
            check(got[0], "skr_plan_launch")
        ACCOUNT["launches"] += 1
        ACCOUNT["bytes"] += got[1]
        return got[0]
    shape = first.shape
    where = first.get_device()  # every other operand must report the same ordinal (a CPU tensor reports -1)
    signature = []
    pointers = []
    for t in inputs:
        if t.get_device() != where or t.shape != shape or not t.is_contiguous():
            return None
        signature.append(t.dtype)
        pointers.append(t.data_ptr())
    signature = tuple(signature)
    plan = compiled.plans.get(signature, _MISSING)
    if plan is _MISSING:
        plan = compiled.specialise(signature)
    if plan is None:
        return None
    if len(draws or ()) != compiled.n_philox:
        return None
    device = first.device
    outputs = []
    for dtype in plan.out_dtypes:
        out = torch.empty(shape, dtype=dtype, device=device)
        pointers.append(out.data_ptr())
        outputs.append(out)
    plan.packer.pack_into(plan.table, 0, *pointers)
    numel = first.numel()
    keys = None
    if draws:
        for d in draws:
            if d.numel != numel or d.device != device:
                return None
        keys = _pack_draw_table(draws)
    if ACCOUNT["on"]:
        ACCOUNT["launches"] += 1
        ACCOUNT["bytes"] += sum(t.numel() * t.element_size() for t in inputs) + sum(t.numel() * t.element_size() for t in outputs)
    lib = _lib if _lib is not None else load()
    if torch._C._cuda_getDevice() != where:
        with torch.cuda.device(device):
            status = lib.skr_plan_launch(plan.handle, plan.table, numel, keys, torch._C._cuda_getCurrentRawStream(where))
    else:
        status = lib.skr_plan_launch(plan.handle, plan.table, numel, keys, torch._C._cuda_getCurrentRawStream(where))
    if status:
        check(status, "skr_plan_launch")
    return outputs


def _forget_plans() -> None:
    _compiled.by_ops.clear()
    from skrample_b200.sampling import functional, plan

    plan.clear()
    functional._scripts.known.clear()


def set_arithmetic(mode: str) -> None:
    """"exact" (default): fp32 / fp64 results bit-identical to the reference's torch-CPU path.  "contracted": the
    issue-bound steps (UniP / UniPC / SPC / Adams) use fused multiply-adds and reciprocal multiplication - a few fp32 ulp
    per step, inside the 1e-5 relative per step the specification allows, for ~30% fewer instructions.  Process-wide;
    this thread's cached plans are dropped so the next step is planned under the new mode."""
    codes = {"exact": 0, "contracted": 1}
    if mode not in codes:
        raise ValueError(f"arithmetic mode {mode!r}: expected 'exact' or 'contracted'")
    check(load().skr_set_arithmetic(codes[mode]), "skr_set_arithmetic")
    _forget_plans()


def get_arithmetic() -> str:
    return "contracted" if load().skr_get_arithmetic() == 1 else "exact"


def reset_switches() -> None:
    """Re-read the library's development switches (SKR_FORCE_INTERP, SKR_NO_PINNED, ...) from the environment and
    forget every plan of this thread that was created under the old ones (tests and A/B tooling)."""
    global _switches_touched
    if _lib is None and not LIB_PATH.exists():
        return
    _switches_touched = True
    load().skr_reload_env()
    _forget_plans()


def error_norms(low: Any, high: Any, power: int) -> tuple[float, float]:
    """(mean |low - high|^power, mean |high|^power) of two same-shaped CUDA tensors: one kernel, one host read
    (reference: FunctionalAdaptive.mae/.mse applied twice, functional.py:197-214)."""
    if low.shape != high.shape or low.dtype != high.dtype or low.device != high.device:
        raise ValueError("error_norms needs two tensors of one shape, dtype and device")
    low, high = low.contiguous(), high.contiguous()
    sums = torch.zeros(MOMENTS_DOUBLES, dtype=torch.float64, device=high.device)
    with torch.cuda.device(high.device):
        status = load().skr_error_norms(low.data_ptr(), high.data_ptr(), DTYPE_CODE[high.dtype], high.numel(), power, sums.data_ptr(), raw_stream())
    check(status, "skr_error_norms")
    count = max(high.numel(), 1)
    total = sums[:2].tolist()  # the adaptive controller needs the value: the one synchronisation of the step
    return total[0] / count, total[1] / count
