"""Functional (Runge-Kutta) samplers: host layer, oracle and CUDA kernels against the reference's golden tensors."""

from __future__ import annotations

import ctypes
import json
import math
from pathlib import Path

import numpy as np
import pytest
import torch

import cases
import oracle_run
from oracle import skrample_oracle as O
from skrample_b200 import scheduling
from skrample_b200.sampling import functional, interface, models, structured, tableaux
from skrample_b200.sampling import program as pg

GOLDEN = Path(__file__).resolve().parent / "golden"
FUNCTIONAL = np.load(GOLDEN / "functional.npz")
INDEX = json.loads((GOLDEN / "functional.json").read_text())


def run_product(case: dict, device: str = "cpu") -> torch.Tensor:
    sampler = cases.make_functional(functional, interface, structured, models, tableaux, case)
    schedule = cases.make_schedule(scheduling, case["schedule"])
    model = cases.make_model(models, case["model"])
    dtype = {"f32": torch.float32, "f64": torch.float64}[case["dtype"]]
    draw = cases.functional_rng(case)
    return sampler.generate_model(
        cases.network, model, schedule, lambda step: torch.from_numpy(draw(step)).to(device=device, dtype=dtype), case["steps"]
    )


def test_fixture_table_is_current() -> None:
    assert [c["id"] for c in INDEX] == [c["id"] for c in cases.FUNCTIONAL_CASES]


@pytest.mark.parametrize("case", INDEX, ids=lambda c: c["id"])
def test_cpu_tensors_match_reference(case: dict) -> None:
    got = run_product(case).numpy()
    want = FUNCTIONAL[f"{case['id']}/final"]
    assert got.dtype == want.dtype
    assert np.array_equal(got, want, equal_nan=True)


def _oracle_supported(case: dict) -> bool:
    return case["sampler"] in ("RKUltra", "DynasauRK") and case["kw"].get("order") in (1, 2, 3, 4) and not (
        case["sampler"] == "RKUltra" and case["kw"].get("order") == 3
    )


def _oracle_tableau(case: dict, step: O.St) -> O.Tableau:
    kw = case["kw"]
    if case["sampler"] == "DynasauRK":
        return O.dynasaurk_tableau(step, kw["order"])
    if kw.get("providers") == "heun":
        return O.HEUN
    order = kw["order"]
    if order == 1:
        return O.Tableau(((0.0, ()),), ((1.0,),))
    if order == 2:
        return O.rk2_tableau(1 / 2)  # RK2.Mid, reference: functional.py:20
    return O.ees27_tableau(1 / 14 * (5 - 3 * 2**0.5))  # RK2.EES7_MIN, reference: functional.py:22


@pytest.mark.parametrize("case", [c for c in INDEX if _oracle_supported(c)], ids=lambda c: c["id"])
def test_oracle_matches_reference_functional(case: dict) -> None:
    "Pins oracle.step_tableau / tableaux generators against the reference (bit-exact fp32/fp64)."
    np_dtype = {"f32": np.float32, "f64": np.float64}[case["dtype"]]
    sch = oracle_run.schedule(case["schedule"])
    model = oracle_run.MODELS[case["model"]]
    kw = case["kw"]
    deriv = O.DATA
    if "derivative_transform" in kw:
        deriv = None if kw["derivative_transform"] is None else O.Model(oracle_run.MODELS[kw["derivative_transform"]].kind)
    draw = cases.functional_rng(case)
    sample = draw(None).astype(np_dtype)
    steps = case["steps"]
    for n in range(steps):
        step = O.St.from_int(n, steps)
        noise = draw(step).astype(np_dtype)
        sample = O.step_tableau(_oracle_tableau(case, step), sample, cases.network, model, sch, step, deriv, noise, kw.get("stochasticity", 0))[0]
    want = FUNCTIONAL[f"{case['id']}/final"]
    assert np.array_equal(np.asarray(sample), want, equal_nan=True)


def test_rk_programs_are_block_shaped(monkeypatch: pytest.MonkeyPatch) -> None:
    "Every launch of a 4-stage flow-matching RK step takes the structured kernel (host-side classification)."
    from skrample_b200 import native

    kinds: list[int] = []
    real = pg.execute

    def spy(program: pg.Program):
        if all(isinstance(v, torch.Tensor) for v in program.inputs):
            outs = [torch.empty_like(program.inputs[0]) for _ in program.outputs]
            kinds.append(native.load().skr_program_classify(ctypes.byref(native.pack_program(program, list(program.inputs), outs))))
        return real(program)

    monkeypatch.setattr(pg, "execute", spy)
    case = next(c for c in INDEX if c["sampler"] == "RKUltra" and c["kw"] == {"order": 4} and c["schedule"] == "flow" and c["dtype"] == "f32")
    run_product(case)
    assert kinds and all(k == 0 for k in kinds), kinds


@pytest.mark.gpu
@pytest.mark.parametrize("case", INDEX, ids=lambda c: c["id"])
def test_cuda_matches_reference_golden(case: dict) -> None:
    from skrample_b200 import native

    before = native.launch_count()
    got = run_product(case, device="cuda").cpu().numpy()
    assert native.launch_count() > before
    want = FUNCTIONAL[f"{case['id']}/final"]
    assert got.dtype == want.dtype
    assert np.array_equal(got, want, equal_nan=True), f"max abs diff {np.nanmax(np.abs(got - want))}"


@pytest.mark.gpu
def test_rkultra4_launch_count() -> None:
    "A 4-stage RK step is 4 fused launches (one per model-call boundary)."
    from skrample_b200 import native
    from skrample_b200.common import Step

    x = torch.randn(4, 16, 32, 32, device="cuda")
    sampler = functional.RKUltra(order=4)
    assert len(sampler.tableau().stages) == 4
    before = native.launch_count()
    sampler.step(x, cases.network, models.FlowModel(), scheduling.FlowShift(scheduling.Linear()), Step.from_int(3, 10))
    assert native.launch_count() - before == 4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Feagin14", "Feagin12", "Stepanov10"])
def test_many_stage_tableaux_chunk_across_launches(name: str) -> None:
    "35/25/15-stage tableaux exceed one launch's 32 tensors: partial sums continue across launches, bit-exact vs CPU."
    from skrample_b200.common import Step

    tab = getattr(tableaux.RKZ, name).tableau()
    schedule, model = scheduling.FlowShift(scheduling.Linear()), models.FlowModel()
    x = torch.randn(3000, generator=torch.Generator().manual_seed(4))
    noise = torch.randn(3000, generator=torch.Generator().manual_seed(5))
    step = Step.from_int(2, 6)
    want = functional.step_tableau(tab, x, cases.network, model, schedule, step, models.DataModel(), noise, 1.0)[0]
    got = functional.step_tableau(tab, x.cuda(), cases.network, model, schedule, step, models.DataModel(), noise.cuda(), 1.0)[0]
    assert torch.equal(got.cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16])
@pytest.mark.parametrize("power", [1, 2])
def test_fused_error_norms_match_the_evaluators(dtype: torch.dtype, power: int) -> None:
    "skr_error_norms = FunctionalAdaptive.mae/.mse on (low, high) and (0, high) (reference: functional.py:197-214)."
    from skrample_b200 import native
    from skrample_b200.sampling.functional import FunctionalAdaptive

    g = torch.Generator(device="cuda").manual_seed(8)
    high = torch.randn(3, 4, 57, 61, device="cuda", generator=g).to(dtype)
    low = (high.double() + 1e-2 * torch.randn(high.shape, device="cuda", generator=g).double()).to(dtype)
    difference, magnitude = native.error_norms(low, high, power)
    evaluator = FunctionalAdaptive.mse if power == 2 else FunctionalAdaptive.mae
    want_difference, want_magnitude = evaluator(low.double(), high.double()), evaluator(0, high.double())
    # 16-bit / fp32 inputs: differences and per-vector partial sums in fp32, accumulated in fp64 in a fixed order - far
    # inside the precision of the reference's own evaluators, which reduce in the tensor's dtype (fp32: ~1e-7)
    tolerance = 1e-12 if dtype == torch.float64 else 2e-7
    assert math.isclose(difference, want_difference, rel_tol=tolerance)
    assert math.isclose(magnitude, want_magnitude, rel_tol=tolerance)
    assert native.error_norms(low, high, power) == (difference, magnitude), "grid-wide sums are added in a fixed order: bit-reproducible"
    assert native.error_norms(low[:0], high[:0], power) == (0.0, 0.0)
    with pytest.raises(ValueError):
        native.error_norms(low, high[:1], power)


@pytest.mark.gpu
def test_rkmoire_takes_one_host_read_per_adaptive_step(monkeypatch: pytest.MonkeyPatch) -> None:
    "The stock evaluators go through the fused reduction; a user evaluator is still called as in the reference."
    from skrample_b200 import native, scheduling
    from skrample_b200.sampling import functional, models

    calls = {"fused": 0, "user": 0}
    real = native.error_norms
    monkeypatch.setattr(native, "error_norms", lambda *a: (calls.__setitem__("fused", calls["fused"] + 1), real(*a))[1])
    x = torch.randn(2, 4, 16, 16, device="cuda")

    def model(sample: torch.Tensor, t: float, sigma: float, alpha: float) -> torch.Tensor:
        return sample * 0.25

    stock = functional.RKMoire(order=4)
    out = stock.sample_model(x, model, models.FlowModel(), scheduling.Linear(), 12)
    assert calls["fused"] > 0 and torch.isfinite(out).all()

    def evaluator(a, b) -> float:  # noqa: ANN001
        calls["user"] += 1
        return functional.FunctionalAdaptive.mse(a, b)

    fused_before = calls["fused"]
    custom = functional.RKMoire(order=4, evaluator=evaluator)
    same = custom.sample_model(x, model, models.FlowModel(), scheduling.Linear(), 12)
    assert calls["user"] > 0 and calls["fused"] == fused_before
    assert torch.equal(out, same), "same strides, same kernels: the fused norm must not change the trajectory here"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize(
    "make",
    [
        lambda f: f.RKUltra(order=2),
        lambda f: f.RKUltra(order=4, stochasticity=1),
        lambda f: f.RKUltra(order=7),
        lambda f: f.RKUltra(order=4, derivative_transform=None),
        lambda f: f.DynasauRK(order=4),
    ],
    ids=["rku2", "rku4-sde", "rku7", "rku4-raw", "dynasaurk4"],
)
def test_rk_step_replay_is_bit_identical(make, dtype: torch.dtype, monkeypatch: pytest.MonkeyPatch) -> None:  # noqa: ANN001
    "The second time a step is taken its recorded launches are replayed: no program is emitted, same bits."
    from skrample_b200 import scheduling
    from skrample_b200.sampling import functional, models
    from skrample_b200.sampling import program as pg

    sampler = make(functional)
    schedule, model_transform = scheduling.FlowShift(scheduling.Linear(), shift=3.0), models.FlowModel()
    g = torch.Generator(device="cuda").manual_seed(21)
    shape = (2, 4, 33, 31)
    emitted = {"n": 0}
    real = pg.execute
    monkeypatch.setattr(pg, "execute", lambda program: (emitted.__setitem__("n", emitted["n"] + 1), real(program))[1])
    functional._scripts.known.clear()

    def trajectory(x0: torch.Tensor, weights: list[torch.Tensor], noises: list[torch.Tensor]) -> torch.Tensor:
        calls = iter(range(1 << 20))

        def model(sample: torch.Tensor, t: float, sigma: float, alpha: float) -> torch.Tensor:
            return (sample.float() * weights[next(calls) % len(weights)].float()).to(sample.dtype)

        draws = iter(noises)
        return sampler.sample_model(x0, model, model_transform, schedule, 5, rng=lambda step=None: next(draws))

    def inputs() -> tuple:
        x0 = torch.randn(shape, device="cuda", generator=g).to(dtype)
        weights = [torch.randn(shape, device="cuda", generator=g).to(dtype) * 0.2 for _ in range(5)]
        noises = [torch.randn(shape, device="cuda", generator=g).to(dtype) for _ in range(8)]
        return x0, weights, noises

    first = inputs()
    want_first = trajectory(*first)
    recorded = emitted["n"]
    assert recorded > 0 and functional._scripts.known
    second = inputs()
    got_second = trajectory(*second)  # replayed
    assert emitted["n"] == recorded, "a cached step emitted programs again"
    functional._scripts.known.clear()
    want_second = trajectory(*second)  # emitted afresh on the same data
    assert torch.equal(got_second, want_second)
    assert torch.equal(trajectory(*first), want_first)


@pytest.mark.gpu
def test_rk_replay_is_not_used_when_roles_are_ambiguous() -> None:
    "A model that hands back its input (or one tensor twice) cannot be recorded by tensor identity: steps stay correct."
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import functional, models

    functional._scripts.known.clear()
    sampler = functional.RKUltra(order=4)
    schedule, model_transform = scheduling.Linear(), models.FlowModel()
    x = torch.randn(2, 4, 16, 16, device="cuda")
    constant = torch.randn_like(x)
    a = sampler.step(x, lambda s, t, sigma, alpha: constant, model_transform, schedule, Step.from_int(2, 10))
    assert not functional._scripts.known
    other = [torch.randn_like(x) for _ in range(4)]
    calls = iter(range(100))
    b = sampler.step(x, lambda s, t, sigma, alpha: other[next(calls)], model_transform, schedule, Step.from_int(2, 10))
    assert functional._scripts.known and not torch.equal(a, b)
    calls = iter(range(100))
    c = sampler.step(x, lambda s, t, sigma, alpha: constant, model_transform, schedule, Step.from_int(2, 10))  # replayed with one tensor four times
    assert torch.equal(a, c)
