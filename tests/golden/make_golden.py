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


def build_functional() -> None:
    import skrample.sampling.functional as ref_functional
    import skrample.sampling.interface as ref_interface
    import skrample.sampling.tableaux as ref_tableaux

    arrays: dict[str, np.ndarray] = {}
    index: list[dict] = []
    for case in cases.FUNCTIONAL_CASES:
        sampler = cases.make_functional(ref_functional, ref_interface, ref_structured, ref_models, ref_tableaux, case)
        schedule = cases.make_schedule(ref_scheduling, case["schedule"])
        model = cases.make_model(ref_models, case["model"])
        dtype = {"f32": torch.float32, "f64": torch.float64}[case["dtype"]]
        draw = cases.functional_rng(case)
        result = sampler.generate_model(
            cases.network, model, schedule, lambda step: torch.from_numpy(draw(step)).to(dtype), case["steps"]
        )
        arrays[f"{case['id']}/final"] = result.numpy()
        index.append(case)
    np.savez_compressed(HERE / "functional.npz", **arrays)
    (HERE / "functional.json").write_text(json.dumps(index, indent=1))
    print("functional:", len(index), "cases,", (HERE / "functional.npz").stat().st_size // 1024, "KiB")


NOISE_PYRAMID_CASES = [((4, 64, 64), (-1, -2), 99), ((3, 40, 56), (-1, -2), 1), ((5, 96), (-1,), 99)]
NOISE_COLORED_CASES = [((4, 32, 32), 1.5, None), ((2, 1, 40, 31), -2.0, 3.0), ((4096,), 0.7, None)]


def build_noise() -> None:
    "Reference Pyramid / Colored outputs together with every random draw they consumed (recorded by patching torch)."
    import skrample.pytorch.noise as ref_noise

    arrays: dict[str, np.ndarray] = {}
    for n, (shape, dims, depth) in enumerate(NOISE_PYRAMID_CASES):
        drawn: list[tuple[str, np.ndarray]] = []
        real_randn, real_rand = torch.randn, torch.rand

        def randn(*a, **k):
            out = real_randn(*a, **k)
            drawn.append(("randn", out.numpy().copy()))
            return out

        def rand(*a, **k):
            out = real_rand(*a, **k)
            drawn.append(("rand", out.numpy().copy()))
            return out

        torch.randn, torch.rand = randn, rand
        try:
            gen = ref_noise.Pyramid.from_inputs(shape, torch.Generator().manual_seed(40 + n), ref_noise.PyramidProps(dims=dims, depth=depth))
            out = gen.generate(None)
        finally:
            torch.randn, torch.rand = real_randn, real_rand
        normals = [v for kind, v in drawn if kind == "randn"]
        arrays[f"pyramid{n}/out"] = out.numpy()
        arrays[f"pyramid{n}/base"] = normals[0]
        arrays[f"pyramid{n}/ratios"] = np.asarray([float(v[0]) * 2 + 2 for kind, v in drawn if kind == "rand"])
        for l, level in enumerate(normals[1:]):
            arrays[f"pyramid{n}/level{l}"] = level
    for n, (shape, exponent, energy) in enumerate(NOISE_COLORED_CASES):
        white = torch.randn(shape, generator=torch.Generator().manual_seed(60 + n))
        arrays[f"colored{n}/white"] = white.numpy()
        arrays[f"colored{n}/out"] = ref_noise.Colored.colorize_noise(white, exponent, energy).numpy()
    np.savez_compressed(HERE / "noise.npz", **arrays)
    print("noise:", len(arrays), "arrays,", (HERE / "noise.npz").stat().st_size // 1024, "KiB")


def build_wrappers() -> None:
    import skrample.diffusers as ref_diffusers
    import skrample.pytorch.noise as ref_noise

    arrays: dict[str, np.ndarray] = {}
    for case in cases.WRAPPER_CASES:
        final, pred = cases.run_wrapper(ref_diffusers, ref_noise, case)
        arrays[f"{case['id']}/final"] = final.float().numpy()
        arrays[f"{case['id']}/pred"] = pred.float().numpy()
    np.savez_compressed(HERE / "wrappers.npz", **arrays)
    (HERE / "wrappers.json").write_text(json.dumps(cases.WRAPPER_CASES, indent=1))
    print("wrappers:", len(cases.WRAPPER_CASES), "cases,", (HERE / "wrappers.npz").stat().st_size // 1024, "KiB")


def build_schedules() -> None:
    "Exact float64 points of schedule stacks (hex floats): sigma / alpha / timestep must be bit-identical."
    table = {}
    with np.errstate(all="ignore"):
        for n, case in enumerate(cases.SCHEDULE_CASES):
            sch = cases.make_schedule_stack(ref_scheduling, case)
            table[str(n)] = {
                "points": [[float(v).hex() for v in row] for row in sch.points_np(cases.SCHEDULE_TIMES).tolist()],
                "ipoints": [[float(v).hex() for v in row] for row in sch.ipoints_np(cases.SCHEDULE_TIMES).tolist()],
                "schedule7": [[float(v).hex() for v in row] for row in sch.schedule_np(7).tolist()],
                "schedule25": [[float(v).hex() for v in row] for row in sch.schedule_np(25).tolist()],
            }
    (HERE / "schedules.json").write_text(json.dumps(table))
    print("schedules:", len(table), "stacks")


if __name__ == "__main__":
    build_schedules()
    build_structured()
    build_functional()
    build_noise()
    build_wrappers()
