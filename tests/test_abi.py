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
    assert ctypes.sizeof(native.SkrPhilox) == 8 * 256 + 8 * 256 + 8 + 4 + 4 + 8 + 4 + 4
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


def _euler_program(dtype: int = 0):
    "X = in0; P = in1; R = X*g + P*d + in2*z; out0 = R - an Euler SDE step, as the samplers emit it."
    from skrample_b200 import native
    from skrample_b200.sampling import program as pg

    prog = native.SkrProgram()
    ops = [(pg.OP_LOAD, pg.X, 0, 0, -1, ()), (pg.OP_LOAD, pg.P, 0, 1, -1, ()), (pg.OP_FWD, pg.P, 1, 2, -1, (0.9, 0.1, 0.2)), (pg.OP_STORE, pg.R, 0, -1, 0, ())]
    prog.n_ops, prog.n_inputs, prog.n_outputs = len(ops), 3, 1
    for slot, (code, a, b, src, dst, c) in zip(prog.ops, ops):
        slot.code, slot.a, slot.b, slot.src, slot.dst = code, a, b, src, dst
        for j, v in enumerate(c):
            slot.c[j] = v
    for t in (*prog.inputs[:3], prog.outputs[0]):
        t.dtype = dtype
    return prog


def test_plan_entry_points_without_a_gpu() -> None:
    "skr_plan_create parses and picks a kernel on the host; bad arguments come back as codes, never as crashes."
    from skrample_b200 import native

    lib = native.load()
    handle = ctypes.c_void_p()
    assert lib.skr_plan_create(None, ctypes.byref(handle)) == -1 and handle.value is None
    prog = _euler_program(native.F32)
    assert lib.skr_plan_create(ctypes.byref(prog), None) == -1
    assert lib.skr_plan_create(ctypes.byref(prog), ctypes.byref(handle)) == 0 and handle.value
    assert lib.skr_plan_kind(handle) == 0
    assert lib.skr_plan_shape(handle) == b"euler/f32"
    assert lib.skr_plan_launch(None, None, 16, None, None) == -1
    assert lib.skr_plan_launch(handle, None, 16, None, None) == -1  # tensors missing
    assert lib.skr_plan_launch(handle, None, -1, None, None) == -2
    table = (ctypes.c_uint64 * 4)(0x1000, 0, 0x3000, 0x4000)
    assert lib.skr_plan_launch(handle, table, 16, None, None) == -1 and b"tensor 1" in lib.skr_last_error()
    assert lib.skr_plan_launch(handle, table, 0, None, None) == 0  # an empty latent is a no-op
    lib.skr_plan_destroy(handle)
    lib.skr_plan_destroy(None)
    bf16 = _euler_program(native.BF16)
    assert lib.skr_plan_create(ctypes.byref(bf16), ctypes.byref(handle)) == 0
    assert lib.skr_plan_shape(handle) == b"euler/bf16"
    lib.skr_plan_destroy(handle)
    bad = _euler_program(native.F32)
    bad.ops[2].src = 7
    assert lib.skr_plan_create(ctypes.byref(bad), ctypes.byref(handle)) == -2 and handle.value is None
    odd = _euler_program(native.F32)  # not the head / block skeleton: LOAD S; BLEND -> interpreter
    odd.ops[1].a = 4
    odd.ops[2].code, odd.ops[2].a, odd.ops[2].b, odd.ops[2].src = 17, 0, 0, -1
    odd.ops[3].a = 0
    assert lib.skr_plan_create(ctypes.byref(odd), ctypes.byref(handle)) == 0
    assert lib.skr_plan_kind(handle) == 1 and lib.skr_plan_shape(handle) == b"interpreter"
    lib.skr_plan_destroy(handle)


def test_host_entry_points_are_reentrant() -> None:
    """the C ABI is re-entrant and thread-safe.  ctypes drops the GIL around every call, so eight threads
    really are inside skr_program_describe / skr_program_classify / skr_plan_create at the same time; each must see
    its own result (the describe text differs per thread) and its own error text."""
    import threading

    from skrample_b200 import native

    lib = native.load()
    failures: list[str] = []
    barrier = threading.Barrier(8)

    def worker(index: int) -> None:
        dtype = (native.F32, native.BF16, native.F16, native.F64)[index % 4]
        want = (b"shape=euler/f32", b"shape=euler/bf16", b"shape=euler/f16", b"compute=f64")[index % 4]
        prog = _euler_program(dtype)
        bad = _euler_program(dtype)
        bad.n_ops = 65 + index
        text = ctypes.create_string_buffer(512)
        handle = ctypes.c_void_p()
        barrier.wait()
        for _ in range(2000):
            if lib.skr_program_describe(ctypes.byref(prog), text, len(text)) != 0 or want not in text.value:
                failures.append(f"describe[{index}]: {text.value!r}")
                return
            if lib.skr_program_classify(ctypes.byref(prog)) != 0:
                failures.append(f"classify[{index}]")
                return
            if lib.skr_plan_create(ctypes.byref(prog), ctypes.byref(handle)) != 0 or lib.skr_plan_kind(handle) != 0:
                failures.append(f"plan[{index}]")
                return
            lib.skr_plan_destroy(handle)
            if lib.skr_program_launch(ctypes.byref(bad), 16, None) != -2 or f"n_ops {65 + index} ".encode() not in lib.skr_last_error():
                failures.append(f"error text[{index}]: {lib.skr_last_error()!r}")
                return

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures, failures[:3]
