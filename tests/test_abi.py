"""The C-ABI library loads here (no GPU) and exports every symbol the header declares."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "skrample_b200.h").read_text()


def _declared() -> list[str]:
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(skr_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_header_symbols() -> None:
    from skrample_b200 import build, native

    lib = ctypes.CDLL(str(build.build()))
    names = _declared()
    assert names, "no declarations parsed from the header"
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/skrample_b200.h but not exported"
    assert set(native.EXPORTS) <= set(names)
    lib.skr_version.restype = ctypes.c_int
    assert lib.skr_version() >= 1


def test_struct_layout_matches_header() -> None:
    from skrample_b200 import native

    assert ctypes.sizeof(native.SkrOp) == 40
    assert ctypes.sizeof(native.SkrTensor) == 16
    assert ctypes.sizeof(native.SkrPhilox) == 8 * 32 + 8 * 32 + 8 + 4 + 4
    assert ctypes.sizeof(native.SkrProgram) == 16 + 40 * 64 + 16 * 32 + 16 * 8 + 2 * ctypes.sizeof(native.SkrPhilox)


def test_bad_arguments_are_reported_without_a_gpu() -> None:
    from skrample_b200 import native

    lib = native.load()
    assert lib.skr_program_launch(None, 16, None) == -1
    prog = native.SkrProgram()
    prog.n_ops = 65
    assert lib.skr_program_launch(ctypes.byref(prog), 16, None) == -2
    assert b"n_ops" in lib.skr_last_error()
    prog.n_ops = 1
    prog.ops[0].code = 200
    assert lib.skr_program_launch(ctypes.byref(prog), 16, None) == -4


def test_op_codes_in_lockstep() -> None:
    from skrample_b200.sampling import program as pg

    for name, value in re.findall(r"SKR_OP_([A-Z0-9]+) = (\d+)", HEADER):
        assert getattr(pg, f"OP_{name}") == int(value), name
    for name, value in re.findall(r"SKR_CONV_([A-Z_]+) = (\d+)", HEADER):
        assert getattr(pg, f"CONV_{name}") == int(value), name


def test_tensor_tables_are_written_in_the_struct_layout() -> None:
    "native.launch_compiled writes a whole tensor table with one struct.pack_into; the ctypes fields must read it back."
    from skrample_b200 import native

    packed = native.SkrProgram()
    native._packer(3).pack_into(packed, native._INPUTS_AT, 0x1000, native.F32, 0xFFFF_FFFF_FFF0, native.F64, 0x2000, native.DTYPE_CODE[__import__("torch").bfloat16])
    native._packer(2).pack_into(packed, native._OUTPUTS_AT, 0x3000, native.F32, 0x4000, native.DTYPE_CODE[__import__("torch").float16])
    assert [(t.ptr, t.dtype, t.reserved) for t in packed.inputs[:4]] == [
        (0x1000, native.F32, 0),
        (0xFFFF_FFFF_FFF0, native.F64, 0),
        (0x2000, native.DTYPE_CODE[__import__("torch").bfloat16], 0),
        (None, 0, 0),
    ]
    assert [(t.ptr, t.dtype) for t in packed.outputs[:3]] == [(0x3000, native.F32), (0x4000, native.DTYPE_CODE[__import__("torch").float16]), (None, 0)]
    assert packed.n_ops == 0 and packed.n_inputs == 0
