"""Host-side check that the standard samplers' step programs take the structured (block) kernel."""

from __future__ import annotations

import ctypes
import json
from pathlib import Path

import pytest
import torch

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
