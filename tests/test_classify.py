"""Host-side check that the standard samplers' step programs take the structured (block) kernel."""

from __future__ import annotations

import ctypes
import json
from pathlib import Path

import pytest
import torch

import cases
from skrample_b200 import native
from skrample_b200.sampling import program as pg
from test_host_layer import run_product

INDEX = json.loads((Path(__file__).resolve().parent / "golden" / "structured.json").read_text())


def classify(program: pg.Program) -> int:
    tensors = [v for v in program.inputs]
    outs = [torch.empty_like(tensors[0]) for _ in program.outputs]
    packed = native.pack_program(program, tensors, outs)
    return native.load().skr_program_classify(ctypes.byref(packed))


@pytest.mark.parametrize("case", [c for c in INDEX if c["dtype"] == "f32"], ids=lambda c: c["id"])
def test_structured_programs_are_block_shaped(case: dict, monkeypatch: pytest.MonkeyPatch) -> None:
    seen: list[int] = []
    real = pg.execute

    def spy(program: pg.Program):
        seen.append(classify(program))
        return real(program)

    monkeypatch.setattr(pg, "execute", spy)
    run_product(case)
    if cases.composite(case):  # separately fused launches; only the standalone blend is interpreter-shaped
        assert seen and 0 in seen and seen.count(1) <= case["steps"], seen
        return
    assert seen and all(kind == 0 for kind in seen), seen


def test_bf16_unipc_program_with_two_xhat_stores_is_block_shaped(monkeypatch: pytest.MonkeyPatch) -> None:
    "UniPC through the wrapper stores x-hat twice (fp32 state + 16-bit copy for the pipeline) and must stay on the fast path."
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    seen: list[tuple[int, int]] = []

    def spy(program: pg.Program):
        stores_of_p = sum(1 for op in program.ops if op.code == pg.OP_STORE and op.a == pg.P)
        seen.append((classify(program), stores_of_p))
        return pg.execute_generic(program)

    monkeypatch.setattr(pg, "execute", spy)
    monkeypatch.setattr(pg, "is_cuda_tensor", lambda v: isinstance(v, torch.Tensor))  # emit what a device run would emit
    sampler = structured.UniPC(order=3, stochasticity=1)
    x = torch.randn(64).bfloat16()
    previous: list = []
    with structured.step_options(final_dtype=torch.bfloat16):
        for n in range(4):
            res = sampler.sample(x, torch.randn(64).bfloat16(), Step.from_int(n, 8), models.NoiseModel(), scheduling.Scaled(), torch.randn(64), previous)
            previous = (previous + [res])[-sampler.require_previous :]
            x = res.final.bfloat16()
    assert len(seen) == 4 and all(kind == 0 and stores == 2 for kind, stores in seen), seen


def describe(program: pg.Program, state_dtype: torch.dtype = torch.float32) -> str:
    "skr_program_describe with outputs typed the way the device path allocates them (solver state in fp32)."
    tensors = list(program.inputs)
    first = tensors[0].dtype
    dtypes = [d if isinstance(d, torch.dtype) else (state_dtype if d == "compute" else first) for d in program.outputs]
    packed = native.pack_program(program, tensors, [torch.empty(0, dtype=d) for d in dtypes])
    text = ctypes.create_string_buffer(1024)
    assert native.load().skr_program_describe(ctypes.byref(packed), text, len(text)) in (0, 1)
    return text.value.decode()


PINNED = [
    ("Euler", {}, "euler"),
    ("Euler", {"stochasticity": 1}, "euler"),
    ("DPM", {"order": 2}, "dpm2"),
    ("DPM", {"order": 3, "stochasticity": 1}, "dpm3"),
    ("Adams", {"order": 4}, "acc"),
    ("UniP", {"order": 3}, "unip"),
    ("UniPC", {"order": 3, "stochasticity": 1}, "unipc3"),
    ("UniPC", {"order": 2}, "unipc2"),
    ("UniPC", {"order": 4}, "unipc"),
    ("SPC", {}, "spc"),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize(("sampler", "kw", "shape"), PINNED, ids=[f"{s}-{n}" for n, (s, _, _) in enumerate(PINNED)])
def test_steady_state_steps_take_a_pinned_kernel_shape(sampler: str, kw: dict, shape: str, dtype: torch.dtype, monkeypatch: pytest.MonkeyPatch) -> None:
    "Past the warm-up steps every standard sampler runs a block-kernel instantiation compiled for exactly its step."
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    seen: list[str] = []

    def spy(program: pg.Program):
        seen.append(describe(program))
        first = program.inputs[0].dtype
        dtypes = [d if isinstance(d, torch.dtype) else (torch.float32 if d == "compute" else first) for d in program.outputs]
        return [out.to(d) for out, d in zip(pg.execute_generic(program), dtypes)]

    monkeypatch.setattr(pg, "execute", spy)
    monkeypatch.setattr(pg, "is_cuda_tensor", lambda v: isinstance(v, torch.Tensor))  # emit what a device run would emit
    instance = getattr(structured, sampler)(**kw)
    x = torch.randn(64).to(dtype)
    previous: list = []
    for n in range(6):
        res = instance.sample(x, torch.randn(64).to(dtype), Step.from_int(n, 8), models.NoiseModel(), scheduling.Scaled(), torch.randn(64).to(dtype), previous)
        previous = (previous + [res])[-instance.require_previous :] if instance.require_previous else []
        x = res.final.to(dtype)
    name = {torch.float32: "f32", torch.bfloat16: "bf16", torch.float16: "f16"}[dtype]
    assert len(seen) == 6
    assert f"shape={shape}/{name} " in seen[-1], seen[-1]


def test_describe_reports_interpreter_for_unstructured_programs() -> None:
    program = pg.Program()
    x = torch.randn(8)
    program.load(pg.X, x)
    program.axpby(torch.randn(8), 0.5, 0.25)
    program.store(pg.R)
    assert describe(program) == "interpreter"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("derivative", ["default", None], ids=["converted", "raw"])
def test_rk_stages_take_a_pinned_kernel_shape(dtype: torch.dtype, derivative, monkeypatch: pytest.MonkeyPatch) -> None:  # noqa: ANN001
    "Every launch of an explicit RK step (stage inputs and the final update) runs a shape compiled for it."
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import functional, models

    seen: list[str] = []

    def spy(program: pg.Program):
        seen.append(describe(program))
        first = program.inputs[0].dtype
        dtypes = [d if isinstance(d, torch.dtype) else (torch.float32 if d == "compute" else first) for d in program.outputs]
        return [out.to(d) for out, d in zip(pg.execute_generic(program), dtypes)]

    monkeypatch.setattr(pg, "execute", spy)
    monkeypatch.setattr(pg, "is_cuda_tensor", lambda v: isinstance(v, torch.Tensor))
    sampler = functional.RKUltra(order=4) if derivative == "default" else functional.RKUltra(order=4, derivative_transform=None)
    x = torch.randn(64).to(dtype)
    sampler.step(x, lambda s, t, sigma, alpha: (s * 0.3).to(dtype), models.FlowModel(), scheduling.FlowShift(scheduling.Linear(), shift=3.0), Step.from_int(3, 25))
    name = {torch.float32: "f32", torch.bfloat16: "bf16", torch.float16: "f16"}[dtype]
    assert len(seen) == 4
    if derivative == "default":
        assert all(f"shape=rk/{name} " in line for line in seen), seen
    else:  # the first stage of a raw-derivative step is "stage input = sample" with nothing to launch
        assert all(f"shape=rk-stage/{name} " in line or f"shape=rk-final/{name} " in line for line in seen), seen
