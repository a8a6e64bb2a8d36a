"""Shared parity-case definitions.

The same case table drives (a) the golden-vector generator that runs the reference
(``tests/golden/make_golden.py``), (b) the CPU tests of the oracle and of the host
layer, and (c) the GPU parity tests.  Inputs are regenerated from seeds so fixtures
only hold expected outputs.
"""

from __future__ import annotations

import itertools
from typing import Any

import numpy as np

N_GOLDEN = 1024 + 5  # one full TMA tile + a ragged tail


def _case(sampler: str, kw: dict, schedule: str, model: str, dtype: str, steps: int, seed: int) -> dict:
    kw_id = ",".join(f"{k}={v}" for k, v in sorted(kw.items(), key=lambda kv: kv[0]))
    return {
        "id": f"{sampler}({kw_id})|{schedule}|{model}|{dtype}|{steps}",
        "sampler": sampler,
        "kw": kw,
        "schedule": schedule,
        "model": model,
        "dtype": dtype,
        "steps": steps,
        "seed": seed,
        "numel": N_GOLDEN,
    }


def _structured_cases() -> list[dict]:
    table: list[tuple[str, dict]] = [
        ("Euler", {}),
        ("Euler", {"stochasticity": 1}),
        ("DPM", {"order": 1, "stochasticity": 1}),
        ("DPM", {"order": 2}),
        ("DPM", {"order": 3, "stochasticity": 0.5}),
        ("Adams", {"order": 4}),
        ("Adams", {"order": 9, "stochasticity": 1}),
        ("UniP", {"order": 3}),
        ("UniP", {"order": 2, "fast_solve": True, "stochasticity": 1}),
        ("UniPC", {"order": 3, "stochasticity": 1}),
        ("UniPC", {"order": 9}),
        ("UniPC", {"order": 2, "predictor": ["Adams", {"order": 3}]}),
        ("SPC", {}),
        ("SPC", {"predictor": ["DPM", {"order": 2, "stochasticity": 1}], "corrector": ["UniP", {"order": 3}]}),
        ("Adams", {"order": 3, "derivative_transform": "FlowModel"}),
        ("DPM", {"order": 3, "derivative_transform": None}),
        ("UniPC", {"order": 3, "derivative_transform": "VelocityModel", "stochasticity": 1}),
    ]
    out: list[dict] = []
    seed = 1000
    for (sampler, kw), (schedule, model) in itertools.product(
        table, [("scaled", "NoiseModel"), ("flow", "FlowModel"), ("scaled", "VelocityModel"), ("karras", "DataModel")]
    ):
        for dtype in ("f32", "f64"):
            seed += 1
            out.append(_case(sampler, kw, schedule, model, dtype, 12, seed))
    # Round-2 additions (appended so the seeds of the cases above do not move): the SPC power mean - the one nonlinear op
    # of the step machine (reference: common.py:187-190, structured.py:557-575) - and ScaleX, both as the network's
    # parameterisation and as the derivative space (reference: models.py:184-212).
    seed = 3000
    power_table: list[tuple[str, dict]] = [
        ("SPC", {"power": 2}),
        ("SPC", {"power": 0.5}),
        ("SPC", {"power": 1.7, "invert": True}),
        ("SPC", {"power": 3, "adaptive": False, "bias": 0.25}),
    ]
    for (sampler, kw), (schedule, model) in itertools.product(power_table, [("scaled", "NoiseModel"), ("flow", "FlowModel")]):
        for dtype in ("f32", "f64"):
            seed += 1
            out.append(_case(sampler, kw, schedule, model, dtype, 12, seed))
    scalex_table: list[tuple[str, dict, str]] = [
        ("Euler", {"stochasticity": 1}, "ScaleX"),
        ("Adams", {"order": 4}, "ScaleX"),
        ("DPM", {"order": 3, "stochasticity": 0.5, "derivative_transform": "ScaleX"}, "NoiseModel"),
        ("UniPC", {"order": 3, "stochasticity": 1, "derivative_transform": "ScaleX"}, "FlowModel"),
        ("UniPC", {"order": 3, "stochasticity": 1}, "ScaleX"),
        ("SPC", {"derivative_transform": "ScaleX"}, "VelocityModel"),
    ]
    for sampler, kw, model in scalex_table:
        for dtype in ("f32", "f64"):
            seed += 1
            out.append(_case(sampler, kw, "scaled", model, dtype, 12, seed))
    return out


STRUCTURED_CASES = _structured_cases()


def tolerance(case: dict) -> float | None:
    """None: the case is bit-exact against the reference.  SPC with ``power != 1`` is not, by construction of the
    REFERENCE: ``abs(x) ** f`` on torch-CPU goes through MKL's vector sqrt (exponent 0.5, i.e. power 2 and 0.5; measured
    0.55 ulp, not correctly rounded, and only for tensors longer than a vector) or Sleef's 1-ulp pow (general exponents),
    while NumPy (the oracle) and CUDA (the product: IEEE sqrt, double-precision pow rounded once) each round differently.
    The stated bound is on the whole 12-step trajectory, relative to the tensor's scale: 2e-6 in fp32 (the north star
    allows 1e-4 over a trajectory), 1e-14 in fp64."""
    if case["sampler"] == "SPC" and abs(case["kw"].get("power", 1) - 1) > 1e-8:
        return 2e-6 if case["dtype"] == "f32" else 1e-14
    return None


def composite(case: dict) -> bool:
    """Steps that cannot be ONE program and run as a composition of fused launches (the standalone SPC blend among them,
    which is interpreter-shaped): SPC whose derivative space needs a conversion the corrector repeats (its own
    derivative_transform differs from the SPC's), so the in-register prediction would be clobbered."""
    return case["sampler"] == "SPC" and case["kw"].get("derivative_transform") not in (None, "DataModel") and "derivative_transform" in case["kw"]


def assert_matches(got: np.ndarray, want: np.ndarray, case: dict, what: str = "") -> None:
    "Bit-exact unless ``tolerance(case)`` states a bound (then: max |diff| <= bound * max |want|)."
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    bound = tolerance(case)
    if bound is None:
        assert np.array_equal(got, want, equal_nan=True), f"{what}: max abs diff {np.nanmax(np.abs(got - want))}"
    else:
        assert np.isfinite(got).all(), what
        worst = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
        scale = float(np.abs(want).max())
        assert worst <= bound * scale, f"{what}: max abs diff {worst:.3e} > {bound:.1e} x scale {scale:.3e}"


def make_schedule(mod: Any, name: str) -> Any:
    if name == "scaled":
        return mod.Scaled()
    if name == "flow":
        return mod.FlowShift(mod.Linear())
    if name == "karras":
        return mod.Karras(mod.Scaled())
    if name == "linear":
        return mod.Linear()
    if name == "hyper_linear":
        return mod.Hyper(mod.Linear())
    if name == "hyper_scaled":
        return mod.Hyper(mod.Scaled())
    if name == "sinner_linear":
        return mod.Sinner(mod.Linear())
    raise KeyError(name)


def make_model(mod: Any, name: str | None) -> Any:
    return None if name is None else getattr(mod, name)()


def make_sampler(mod_structured: Any, mod_models: Any, case: dict) -> Any:
    kw = dict(case["kw"])
    for key in ("predictor", "corrector"):
        if key in kw:
            name, sub = kw[key]
            kw[key] = getattr(mod_structured, name)(**sub)
    if "derivative_transform" in kw:
        kw["derivative_transform"] = make_model(mod_models, kw["derivative_transform"])
    return getattr(mod_structured, case["sampler"])(**kw)


def trajectory_inputs(case: dict) -> tuple[np.ndarray, list[np.ndarray], list[np.ndarray]]:
    "(x0, network outputs per step, noises per step) as float32-exact float64 arrays."
    rng = np.random.default_rng(case["seed"])
    n = case["numel"]

    def draw(scale: float = 1.0) -> np.ndarray:
        return (rng.standard_normal(n).astype(np.float32) * np.float32(scale)).astype(np.float64)

    x0 = draw()
    outs = [draw(0.5) for _ in range(case["steps"])]
    noises = [draw() for _ in range(case["steps"])]
    return x0, outs, noises


# ---------------------------------------------------------------------------------------------------------
# functional (Runge-Kutta) cases: whole generate_model trajectories with a closed-form "network"


def _functional_cases() -> list[dict]:
    out: list[dict] = []
    seed = 5000
    table = [
        ("RKUltra", {"order": 1}),
        ("RKUltra", {"order": 2, "stochasticity": 1}),
        ("RKUltra", {"order": 4}),
        ("RKUltra", {"order": 4, "stochasticity": 0.5, "derivative_transform": "FlowModel"}),
        ("RKUltra", {"order": 6, "derivative_transform": None}),
        ("RKUltra", {"order": 15}),
        ("RKUltra", {"order": 2, "providers": "heun"}),
        ("DynasauRK", {"order": 2}),
        ("DynasauRK", {"order": 3, "stochasticity": 1}),
        ("DynasauRK", {"order": 4}),
        ("RKMoire", {"order": 4}),
        ("Adapter", {"sampler": ["DPM", {"order": 3, "stochasticity": 1}]}),
    ]
    for (sampler, kw), (schedule, model) in itertools.product(table, [("scaled", "NoiseModel"), ("flow", "FlowModel"), ("sinner_linear", "VelocityModel")]):
        for dtype in ("f32", "f64"):
            seed += 1
            kw_id = ",".join(f"{k}={v}" for k, v in sorted(kw.items()))
            out.append(
                {
                    "id": f"{sampler}({kw_id})|{schedule}|{model}|{dtype}",
                    "sampler": sampler,
                    "kw": kw,
                    "schedule": schedule,
                    "model": model,
                    "dtype": dtype,
                    "steps": 6,
                    "seed": seed,
                    "numel": N_GOLDEN,
                }
            )
    return out


FUNCTIONAL_CASES = _functional_cases()


def make_functional(mod_functional: Any, mod_interface: Any, mod_structured: Any, mod_models: Any, mod_tableaux: Any, case: dict) -> Any:
    kw = dict(case["kw"])
    if "derivative_transform" in kw:
        kw["derivative_transform"] = make_model(mod_models, kw["derivative_transform"])
    if kw.get("providers") == "heun":
        kw["providers"] = {2: mod_tableaux.RKE2.Heun}
    if case["sampler"] == "Adapter":
        name, sub = kw["sampler"]
        return mod_interface.StructuredFunctionalAdapter(getattr(mod_structured, name)(**sub))
    return getattr(mod_functional, case["sampler"])(**kw)


def network(x: Any, t: float, s: float, a: float) -> Any:
    "Stand-in denoiser: two individually rounded elementwise ops, identical on CPU and CUDA."
    import math

    return x * 0.3 + math.sin(t) * 0.1


def functional_rng(case: dict) -> Any:
    "Deterministic noise source: call k returns the k-th float32-exact normal draw."
    rng = np.random.default_rng(case["seed"])
    n = case["numel"]

    def draw(_step: Any = None) -> np.ndarray:
        return rng.standard_normal(n).astype(np.float32).astype(np.float64)

    return draw


# ---------------------------------------------------------------------------------------------------------
# diffusers-wrapper cases: a pipeline-style loop over wrapper.timesteps with a closed-form "network"

FLOW_CONFIG = {
    "base_image_seq_len": 256,
    "base_shift": 0.5,
    "flow_shift": 3.0,
    "max_image_seq_len": 4096,
    "max_shift": 1.15,
    "num_train_timesteps": 1000,
    "prediction_type": "flow_prediction",
    "shift": 3.0,
    "use_dynamic_shifting": True,
}
SCALED_CONFIG = {
    "beta_end": 0.012,
    "beta_schedule": "scaled_linear",
    "beta_start": 0.00085,
    "num_train_timesteps": 1000,
    "prediction_type": "epsilon",
    "steps_offset": 1,
    "timestep_spacing": "trailing",
    "use_karras_sigmas": False,
}


def _wrapper_cases() -> list[dict]:
    table = [
        ("struct", "scaled", {"_class_name": "UniPCMultistepScheduler", "solver_order": 3}, {}, "Random"),
        ("struct", "scaled", {"_class_name": "DPMSolverSDEScheduler"}, {}, "Random"),
        ("struct", "scaled", {"_class_name": "EulerAncestralDiscreteScheduler", "use_karras_sigmas": True}, {}, "Pyramid"),
        ("struct", "flow", {"_class_name": "FlowMatchEulerDiscreteScheduler"}, {}, "Random"),
        ("struct", "flow", {"_class_name": "IPNDMScheduler"}, {}, "Random"),
        ("struct", "flow", {"_class_name": "MiniMaxH3Scheduler"}, {}, "Random"),
        ("rku", "flow", {}, {"sampler_order": 4}, "Random"),
        ("rku", "scaled", {}, {"sampler_order": 2, "stochasticity": 1}, "Random"),
        ("dyna", "flow", {}, {"sampler_order": 3, "stochasticity": 0.5}, "Random"),
    ]
    out = []
    for n, (kind, base, extra, kw, noise) in enumerate(table):
        for dtype in ("f32", "bf16"):
            out.append({"id": f"{kind}|{base}|{extra.get('_class_name', '')}|{kw}|{noise}|{dtype}", "kind": kind, "base": base, "extra": extra, "kw": kw, "noise": noise, "dtype": dtype, "steps": 8, "mu": 0.8 if base == "flow" else None, "seed": 7000 + n})
    return out


WRAPPER_CASES = _wrapper_cases()


def run_wrapper(mod_diffusers: Any, mod_noise: Any, case: dict, device: str = "cpu") -> Any:
    "A pipeline-style denoising loop; returns (final latents, last pred_original_sample)."
    import torch

    config = (FLOW_CONFIG if case["base"] == "flow" else SCALED_CONFIG) | case["extra"]
    cls = {"struct": "SkrampleWrapperScheduler", "rku": "RKUltraWrapperScheduler", "dyna": "DynasauRKWrapperScheduler"}[case["kind"]]
    wrapper = getattr(mod_diffusers, cls).from_diffusers_config(config, noise_type=getattr(mod_noise, case["noise"]), **case["kw"])
    wrapper.set_timesteps(case["steps"], device=device, mu=case["mu"])
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[case["dtype"]]
    x = torch.randn((2, 4, 24, 24), generator=torch.Generator().manual_seed(case["seed"])).to(device=device, dtype=dtype)
    generators = [torch.Generator().manual_seed(case["seed"] + 1), torch.Generator().manual_seed(case["seed"] + 2)]
    pred = x
    for t in wrapper.timesteps:
        out = (x * 0.31 + x.roll(1, -1) * (0.05 * float(t) / 1000)).to(dtype)  # individually rounded ops, same on CPU and CUDA
        x, pred = wrapper.step(out, t, x, generator=generators, return_dict=False)
    return x, pred


# ---------------------------------------------------------------------------------------------------------
# schedule stacks: (base, base kwargs, sub, sub kwargs, [modifiers...])

SCHEDULE_CASES = [
    ("Linear", {}, None, {}, []),
    ("Linear", {"sigma_start": 14.6}, None, {}, []),
    ("Linear", {"base_timesteps": -1}, None, {}, [("FlowShift", {"shift": 12})]),
    ("Scaled", {}, None, {}, []),
    ("Scaled", {"beta_scale": 1}, None, {}, []),
    ("Scaled", {"base_timesteps": -1000}, None, {}, [("Hyper", {})]),
    ("ZSNR", {}, None, {}, []),
    ("Scaled", {}, "Karras", {}, []),
    ("Scaled", {}, "Exponential", {"rho": 2.0}, []),
    ("Scaled", {}, "Beta", {}, []),
    ("Scaled", {}, "Probit", {}, [("Sinner", {})]),
    ("Linear", {}, "Karras", {"steps": 30}, [("FlowShift", {})]),
    ("Linear", {}, None, {}, [("FlowShift", {"shift": 0.7}), ("Hyper", {"scale": -1.5, "tail": False})]),
    ("Linear", {}, None, {}, [("Sinner", {"count": 3, "scale": -2}), ("NoMod", {})]),
    ("Linear", {}, "NoSub", {}, [("Hyper", {}), ("Hyper", {})]),
]


def make_schedule_stack(mod: Any, case: tuple) -> Any:
    base, base_kw, sub, sub_kw, mods = case
    built = getattr(mod, base)(**base_kw)
    if sub:
        built = getattr(mod, sub)(built, **sub_kw)
    for name, kw in mods:
        built = getattr(mod, name)(built, **kw)
    return built


SCHEDULE_TIMES = [0.0, 1.0, 0.5, 0.04, 0.96, 1 / 3, 0.123456789, 0.75, 2 / 7]
