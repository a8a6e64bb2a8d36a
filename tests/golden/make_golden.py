#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE (Beinsezii/skrample) on CPU.

Run in the build container only (needs /root/reference or $SKRAMPLE_REF):

    python tests/golden/make_golden.py

Outputs ``tests/golden/structured.npz`` + ``structured.json`` (and friends).  The
fixtures travel to the GPU box; the reference does not.  Inputs are regenerated
from NumPy ``default_rng`` seeds on both sides (see ``tests/cases.py``), so only
expected outputs are stored.
"""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))  # tests/ for cases.py
sys.path.insert(0, os.environ.get("SKRAMPLE_REF", "/root/reference"))

import torch  # noqa: E402

import cases  # noqa: E402
import skrample.sampling.models as ref_models  # noqa: E402
import skrample.sampling.structured as ref_structured  # noqa: E402
import skrample.scheduling as ref_scheduling  # noqa: E402
from skrample.common import Step as RefStep  # noqa: E402

torch.set_num_threads(1)


def build_structured() -> None:
    arrays: dict[str, np.ndarray] = {}
    index: list[dict] = []
    for case in cases.STRUCTURED_CASES:
        sampler = cases.make_sampler(ref_structured, ref_models, case)
        schedule = cases.make_schedule(ref_scheduling, case["schedule"])
        model = cases.make_model(ref_models, case["model"])
        dtype = {"f32": torch.float32, "f64": torch.float64}[case["dtype"]]
        x0, outs, noises = cases.trajectory_inputs(case)
        x = torch.from_numpy(x0).to(dtype)
        previous: list = []
        result = None
        for n in range(case["steps"]):
            result = sampler.sample(
                x,
                torch.from_numpy(outs[n]).to(dtype),
                RefStep.from_int(n, case["steps"]),
                model,
                schedule,
                torch.from_numpy(noises[n]).to(dtype) if sampler.require_noise else None,
                previous,
            )
            previous.append(result)
            previous = previous[max(len(previous) - sampler.require_previous, 0) :]
            x = result.final
        assert result is not None
        key = case["id"]
        arrays[f"{key}/final"] = result.final.numpy()
        arrays[f"{key}/sample"] = result.sample.numpy()
        arrays[f"{key}/prediction"] = result.prediction.numpy()
        index.append(case)
    np.savez_compressed(HERE / "structured.npz", **arrays)
    (HERE / "structured.json").write_text(json.dumps(index, indent=1))
    print("structured:", len(index), "cases,", (HERE / "structured.npz").stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    build_structured()
