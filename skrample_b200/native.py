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
import os
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


class SkrProgram(ctypes.Structure):
    _fields_ = [
        ("n_ops", ctypes.c_int32),
        ("n_inputs", ctypes.c_int32),
        ("n_outputs", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("ops", SkrOp * MAX_OPS),
        ("inputs", SkrTensor * MAX_INPUTS),
        ("outputs", SkrTensor * MAX_OUTPUTS),
    ]


EXPORTS = (
    "skr_version",
    "skr_last_error",
    "skr_launch_count",
    "skr_launch_count_kind",
    "skr_program_launch",
    "skr_program_classify",
    "skr_axpby",
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
    _lib = lib
    return lib


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
    return packed


def launch_program(program: "Program") -> list[Any]:
    "Run a step program on CUDA tensors: allocate outputs, one kernel launch, return outputs."
    lib = load()
    inputs = [t if t.is_contiguous() else t.contiguous() for t in program.inputs]
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
    if torch.cuda.current_device() != device.index:
        with torch.cuda.device(device):
            status = lib.skr_program_launch(ctypes.byref(packed), first.numel(), torch.cuda.current_stream().cuda_stream)
    else:
        status = lib.skr_program_launch(ctypes.byref(packed), first.numel(), torch.cuda.current_stream().cuda_stream)
    check(status, "skr_program_launch")
    return outputs
