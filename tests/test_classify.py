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
