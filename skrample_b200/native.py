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
MAX_PHILOX_ITEMS = 32


class SkrPhilox(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64 * MAX_PHILOX_ITEMS),
        ("stream", ctypes.c_uint64 * MAX_PHILOX_ITEMS),
        ("item_numel", ctypes.c_int64),
        ("n_items", ctypes.c_int32),
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
)

_lib: ctypes.CDLL | None = None

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


class CompiledProgram:
    "A packed program whose op table is final; only tensor pointers / dtypes change between launches."

    __slots__ = ("n_inputs", "out_specs", "packed")

    def __init__(self, program: "Program") -> None:
        if len(program.ops) > MAX_OPS or len(program.inputs) > MAX_INPUTS or len(program.outputs) > MAX_OUTPUTS:
            raise RuntimeError("skrample_b200: step program too large")
        packed = SkrProgram()
        packed.n_ops = len(program.ops)
        packed.n_inputs = len(program.inputs)
        packed.n_outputs = len(program.outputs)
        for slot, op in zip(packed.ops, program.ops, strict=False):
            slot.code, slot.a, slot.b, slot.src, slot.dst = op.code, op.a, op.b, op.src, op.dst
            for j, value in enumerate(op.c):
                slot.c[j] = value
        self.packed = packed
        self.n_inputs = len(program.inputs)
        self.out_specs = tuple(program.outputs)


_ALLOWED = frozenset(DTYPE_CODE)


def _on_device(t: Any) -> bool:
    return t.is_cuda


_TENSOR_RECORD = "Qi4x"  # skr_tensor: pointer, dtype code, reserved
_PACKERS: dict[int, struct.Struct] = {}
_INPUTS_AT = SkrProgram.inputs.offset
_OUTPUTS_AT = SkrProgram.outputs.offset
assert ctypes.sizeof(SkrTensor) == struct.calcsize("<" + _TENSOR_RECORD)


def _packer(count: int) -> struct.Struct:
    "One ``struct`` call writes a whole tensor table into the packed program (field-by-field ctypes stores cost more than the launch)."
    made = _PACKERS.get(count)
    if made is None:
        made = _PACKERS[count] = struct.Struct("<" + _TENSOR_RECORD * count)
    return made


def launch_compiled(compiled: CompiledProgram, inputs: list[Any], draws: list[Any] | None = None) -> list[Any] | None:
    "Bind tensors (and lazy noise draws) to a compiled program and launch it.  None when they do not qualify."
    first = inputs[0]
    if not _on_device(first):
        return None
    shape, device = first.shape, first.device
    where = first.get_device()  # every other operand must report the same ordinal (a CPU tensor reports -1)
    packed = compiled.packed
    any64 = any32 = False
    kind = first.dtype
    mixed = False
    code_of = DTYPE_CODE.get
    table: list[int] = []
    for t in inputs:
        dtype = t.dtype
        code = code_of(dtype)
        if code is None or t.get_device() != where or t.shape != shape or not t.is_contiguous():
            return None
        if code == F64:
            any64 = True
        elif code == F32:
            any32 = True
        if dtype != kind:
            mixed = True
        table.append(t.data_ptr())
        table.append(code)
    _packer(len(inputs)).pack_into(packed, _INPUTS_AT, *table)
    default = torch.float64 if any64 else (torch.float32 if any32 or mixed else kind)
    compute = torch.float64 if any64 else torch.float32
    outputs = []
    table = []
    for want in compiled.out_specs:
        dtype = default if want is None else (compute if want.__class__ is str else want)
        out = torch.empty(shape, dtype=dtype, device=device)
        table.append(out.data_ptr())
        table.append(DTYPE_CODE[dtype])
        outputs.append(out)
    if outputs:
        _packer(len(outputs)).pack_into(packed, _OUTPUTS_AT, *table)
    if draws:
        numel = first.numel()
        for d in draws:
            if d.numel != numel or d.device != device:
                return None
        _pack_draws(packed, draws)
    if ACCOUNT["on"]:
        ACCOUNT["launches"] += 1
        ACCOUNT["bytes"] += sum(t.numel() * t.element_size() for t in inputs) + sum(t.numel() * t.element_size() for t in outputs)
    lib = _lib if _lib is not None else load()
    index = device.index
    if torch._C._cuda_getDevice() != index:
        with torch.cuda.device(device):
            status = lib.skr_program_launch(ctypes.byref(packed), first.numel(), torch._C._cuda_getCurrentRawStream(index))
    else:
        status = lib.skr_program_launch(ctypes.byref(packed), first.numel(), torch._C._cuda_getCurrentRawStream(index))
    if status:
        check(status, "skr_program_launch")
    return outputs


def error_norms(low: Any, high: Any, power: int) -> tuple[float, float]:
    """(mean |low - high|^power, mean |high|^power) of two same-shaped CUDA tensors: one kernel, one host read
    (reference: FunctionalAdaptive.mae/.mse applied twice, functional.py:197-214)."""
    if low.shape != high.shape or low.dtype != high.dtype or low.device != high.device:
        raise ValueError("error_norms needs two tensors of one shape, dtype and device")
    low, high = low.contiguous(), high.contiguous()
    sums = torch.zeros(2, dtype=torch.float64, device=high.device)
    with torch.cuda.device(high.device):
        status = load().skr_error_norms(low.data_ptr(), high.data_ptr(), DTYPE_CODE[high.dtype], high.numel(), power, sums.data_ptr(), raw_stream())
    check(status, "skr_error_norms")
    count = max(high.numel(), 1)
    total = sums.tolist()  # the adaptive controller needs the value: the one synchronisation of the step
    return total[0] / count, total[1] / count
