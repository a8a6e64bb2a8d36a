"""Host scalar layer + generic executor of skrample_b200 against the reference's golden tensors (CPU)."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest
import torch

import cases
from skrample_b200 import scheduling
from skrample_b200.common import Step
from skrample_b200.sampling import models, structured

GOLDEN = Path(__file__).resolve().parent / "golden"
STRUCTURED = np.load(GOLDEN / "structured.npz")
STRUCTURED_INDEX = json.loads((GOLDEN / "structured.json").read_text())


def run_product(case: dict, device: str = "cpu", dtype: torch.dtype | None = None, objects: tuple | None = None):
    "``objects`` = (sampler, schedule, model) to reuse (the step-plan cache is keyed on their identity)."
    if objects is None:
        objects = (cases.make_sampler(structured, models, case), cases.make_schedule(scheduling, case["schedule"]), cases.make_model(models, case["model"]))
    sampler, schedule, model = objects
    dtype = dtype or {"f32": torch.float32, "f64": torch.float64}[case["dtype"]]
    x0, outs, noises = cases.trajectory_inputs(case)
    x = torch.from_numpy(x0).to(device=device, dtype=dtype)
    previous: list = []
    result = None
    for n in range(case["steps"]):
        result = sampler.sample(
            x,
            torch.from_numpy(outs[n]).to(device=device, dtype=dtype),
            Step.from_int(n, case["steps"]),
            model,
            schedule,
            torch.from_numpy(noises[n]).to(device=device, dtype=dtype) if sampler.require_noise else None,
            previous,
        )
        previous.append(result)
        previous = previous[max(len(previous) - sampler.require_previous, 0) :]
        x = result.final
    return result


@pytest.mark.parametrize("case", STRUCTURED_INDEX, ids=lambda c: c["id"])
def test_cpu_tensors_match_reference(case: dict) -> None:
    result = run_product(case)
    for field in ("final", "sample", "prediction"):
        cases.assert_matches(getattr(result, field).numpy(), STRUCTURED[f"{case['id']}/{field}"], case, field)


def test_bench_reference_arm_prints_exactly_one_json_line() -> None:
    """bench.py's contract: stdout carries ONE JSON line (library banners are diverted to stderr); the reference arm
    is the part of the bench that runs without a GPU."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    done = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "25", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert done.returncode == 0, done.stderr[-2000:]
    lines = done.stdout.splitlines()
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "latent-steps/s" and line["value"] > 0
    # the unmodified reference when tools/install_reference.py has put it under baseline/_ref, else the oracle port
    expect = "reference" if (root / "baseline" / "_ref" / "skrample" / "__init__.py").exists() else "port"
    assert line["cpu_baseline"]["kind"] == expect and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["steps"] == 25 and set(line["config"]) == {"workload", "name", "per_gpu_batch", "global_batch", "parallelism"}
