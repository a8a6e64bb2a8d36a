"""diffusers-wrapper facade: CPU path and CUDA path against golden tensors from the reference wrappers."""

from __future__ import annotations

import json
import math
import random
from pathlib import Path

import numpy as np
import pytest
import torch

import cases
from skrample_b200 import diffusers, scheduling
from skrample_b200.common import Point
from skrample_b200.pytorch import noise
from skrample_b200.sampling import models, structured

GOLDEN = Path(__file__).resolve().parent / "golden"
WRAPPERS = np.load(GOLDEN / "wrappers.npz")
INDEX = json.loads((GOLDEN / "wrappers.json").read_text())


def test_fixture_table_is_current() -> None:
    assert [c["id"] for c in INDEX] == [c["id"] for c in cases.WRAPPER_CASES]


@pytest.mark.parametrize("case", INDEX, ids=lambda c: c["id"])
def test_cpu_wrapper_matches_reference(case: dict) -> None:
    final, pred = cases.run_wrapper(diffusers, noise, case)
    assert np.array_equal(final.float().numpy(), WRAPPERS[f"{case['id']}/final"], equal_nan=True)
    assert np.array_equal(pred.float().numpy(), WRAPPERS[f"{case['id']}/pred"], equal_nan=True)


@pytest.mark.parametrize("wrapper", [diffusers.SkrampleWrapperScheduler, diffusers.RKUltraWrapperScheduler, diffusers.DynasauRKWrapperScheduler])
@pytest.mark.parametrize("steps", [1, 7, 30])
def test_timesteps_cover_every_model_call(wrapper: type, steps: int) -> None:
    "reference: tests/self_scheduling.py:128-151 - len(timesteps) == steps * order"
    w = wrapper.from_diffusers_config(cases.FLOW_CONFIG)
    w.set_timesteps(steps)
    assert len(w.timesteps) == steps * w.order
    assert len(w.sigmas) == len(w.timesteps) + 1


def test_mu_overrides_flow_shift() -> None:
    "reference: tests/self_scheduling.py:48-54"
    mu = 1.2345
    a = diffusers.SkrampleWrapperScheduler(structured.DPM(), scheduling.Hyper(scheduling.FlowShift(scheduling.Hyper(scheduling.Linear()))))
    b = diffusers.SkrampleWrapperScheduler(structured.DPM(), scheduling.Hyper(scheduling.FlowShift(scheduling.Hyper(scheduling.Linear()), shift=math.exp(mu))))
    a.set_timesteps(123, mu=mu)
    assert a.schedule == b.schedule


@pytest.mark.parametrize("wrapper", [diffusers.SkrampleWrapperScheduler, diffusers.RKUltraWrapperScheduler, diffusers.DynasauRKWrapperScheduler])
def test_inverted_prediction_is_bit_exact(wrapper: type) -> None:
    "reference: tests/self_sampling.py:540-579 - a + (-b)*c == a - b*c exactly"
    weights = torch.randn([64, 64], dtype=torch.float64)

    def network(x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        return x @ weights + x * (t / 1000)

    forward = wrapper.from_diffusers_config({"shift": 12})
    backward = wrapper.from_diffusers_config({"shift": 12}, invert_prediction=True)
    start = torch.randn_like(weights)
    results = []
    for sched, sign in ((forward, 1), (backward, -1)):
        x = start.clone()
        sched.set_timesteps(num_inference_steps=12)
        for t in sched.timesteps:
            x = sched.step(model_output=sign * network(x, t), timestep=t, sample=x, return_dict=False)[0]
        results.append(x)
    assert torch.equal(results[0], results[1])


@pytest.mark.parametrize("wrapper", [diffusers.RKUltraWrapperScheduler, diffusers.DynasauRKWrapperScheduler])
@pytest.mark.parametrize("order", [0, 2, 3, 4, 99])
@pytest.mark.parametrize("stochasticity", [0, 1])
def test_rk_wrapper_equals_functional_sampler(wrapper: type, order: int, stochasticity: float) -> None:
    "reference: tests/self_sampling.py:417-500 - the inside-out wrapper visits the same points and samples."
    seen_f: list[tuple[float, Point]] = []
    seen_w: list[tuple[float, Point]] = []

    def fake(x: float, _t: float, s: float, _a: float) -> float:
        return x + math.sin(x) * s

    w = wrapper(scheduling.Scaled(), sampler_order=order, stochasticity=stochasticity, model=models.VelocityModel(), compute_scale=torch.float64)
    steps = random.randint(5, 21)
    generator = torch.Generator().manual_seed(42)
    twin = generator.clone_state()
    start = 1 / (random.random() + 1e-4)

    def model_f(x: float, t: float, s: float, a: float) -> float:
        seen_f.append((x, Point(t, s, a)))
        return fake(x, t, s, a)

    want = w.functional_sample_model(start, model_f, steps, rng=lambda _: torch.randn([1], generator=twin).item())
    w.set_timesteps(steps)
    x = start
    for n, (t, s) in enumerate(zip(w.timesteps, w.sigmas)):
        sigma, alpha = (v.item() for v in w.schedule.space.normalize(s.item()))
        out = fake(x, t.item(), sigma, alpha)
        np.testing.assert_allclose((t.item(), sigma, alpha), seen_f[n][1], rtol=0, atol=1e-15)
        assert abs(seen_f[n][0] - x) < 1e-8
        x = w.step(torch.tensor(out, dtype=torch.float64).unsqueeze(0), t, torch.tensor(x, dtype=torch.float64).unsqueeze(0), generator=generator, return_dict=False)[0].squeeze(0).item()
    assert abs(want - x) < 1e-8


# ----------------------------------------------------------------------------------------------- CUDA


@pytest.mark.gpu
@pytest.mark.parametrize("case", INDEX, ids=lambda c: c["id"])
def test_cuda_wrapper_matches_reference(case: dict) -> None:
    """One fused launch per step() call.  fp32: bit-exact.  bf16 storage: the structured wrapper is value-identical to
    the reference's cast-compute-cast; the RK wrappers convert the network output in fp32 inside the kernel where the
    reference converts it in bf16 (diffusers.py:817-824) - for an epsilon model at alpha ~ 0.07 the reference's own
    rounding error is ~ 2^-8 / alpha ~ 6 %, so those cases are compared with rtol 2^-4 + atol 2^-4 * max|x|."""
    from skrample_b200 import native

    before = native.launch_count_kind(0)
    final, pred = cases.run_wrapper(diffusers, noise, case, device="cuda")
    assert native.launch_count_kind(0) - before >= case["steps"]
    want_final, want_pred = WRAPPERS[f"{case['id']}/final"], WRAPPERS[f"{case['id']}/pred"]
    got_final, got_pred = final.float().cpu().numpy(), pred.float().cpu().numpy()
    if case["dtype"] == "f32" or case["kind"] == "struct":
        assert np.array_equal(got_final, want_final, equal_nan=True), np.nanmax(np.abs(got_final - want_final))
        assert np.array_equal(got_pred, want_pred, equal_nan=True)
    else:
        np.testing.assert_allclose(got_final, want_final, rtol=2**-4, atol=2**-4 * float(np.abs(want_final).max()))
        np.testing.assert_allclose(got_pred, want_pred, rtol=2**-4, atol=2**-4 * float(np.abs(want_pred).max()))


@pytest.mark.gpu
def test_cuda_wrapper_device_generators() -> None:
    """CUDA generators: noise is drawn by the Philox kernels, the trajectory is deterministic and finite, and drawing
    the noise inside the step kernel (fused_noise) gives the same latents as reading the filled tensor."""
    outs = []
    for fused in (False, False, True):
        w = diffusers.SkrampleWrapperScheduler.from_diffusers_config(cases.SCALED_CONFIG | {"_class_name": "UniPCMultistepScheduler"}, sampler_props={"stochasticity": 1}, compute_scale=torch.float32)
        w.fused_noise = fused
        w.set_timesteps(10, device="cuda")
        x = torch.randn((2, 4, 32, 32), generator=torch.Generator().manual_seed(1)).cuda().bfloat16()
        gens = [torch.Generator(device="cuda").manual_seed(5), torch.Generator(device="cuda").manual_seed(6)]
        for t in w.timesteps:
            x = w.step((x * 0.3).to(torch.bfloat16), t, x, generator=gens, return_dict=False)[0]
        assert x.dtype == torch.bfloat16 and torch.isfinite(x).all()
        outs.append(x)
    assert torch.equal(outs[0], outs[1])
    # the filled tensor is bf16 (the wrapper generates in the sample's dtype), the in-kernel draw is unrounded fp32
    assert (outs[0].float() - outs[2].float()).abs().max().item() < 0.25


@pytest.mark.parametrize("wrapper", [diffusers.SkrampleWrapperScheduler, diffusers.RKUltraWrapperScheduler])
def test_timestep_views_resolve_without_a_host_read(wrapper: type, monkeypatch: pytest.MonkeyPatch) -> None:
    """`for t in scheduler.timesteps: scheduler.step(out, t, x)`: `t` is a view of a tensor the wrapper handed out, so
    its index comes from its address, not from `.item()` (on a GPU that read drains the stream every step;
    reference: diffusers.py:262-270, 540-548)."""
    sched = wrapper(schedule=scheduling.Scaled()) if wrapper is diffusers.RKUltraWrapperScheduler else wrapper(sampler=structured.DPM(order=2), schedule=scheduling.Scaled())
    sched.set_timesteps(6)
    timesteps = sched.timesteps
    reads: list[int] = []
    real_item = torch.Tensor.item
    real_bool = torch.Tensor.__bool__
    monkeypatch.setattr(torch.Tensor, "item", lambda self: (reads.append(1), real_item(self))[1])
    monkeypatch.setattr(torch.Tensor, "__bool__", lambda self: (reads.append(1), real_bool(self))[1])  # `assert t == x`
    x = torch.randn(2, 4, 8, 8, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    twin = wrapper(schedule=scheduling.Scaled()) if wrapper is diffusers.RKUltraWrapperScheduler else wrapper(sampler=structured.DPM(order=2), schedule=scheduling.Scaled())
    twin.set_timesteps(6)
    y = x.clone()
    for t in timesteps:
        out = torch.randn(x.shape, generator=g, dtype=torch.float64)
        before = len(reads)
        x = sched.step(out, t, x, return_dict=False)[0]
        assert len(reads) == before, "step() read the timestep back from the tensor"
        y = twin.step(out, float(real_item(t)), y, return_dict=False)[0]  # the value path gives the same trajectory
        assert torch.equal(x, y)
    # a tensor that is not ours (a copy) still works, through the value
    sched.set_timesteps(6)
    copy = sched.timesteps.clone()
    before = len(reads)
    sched.step(torch.randn(x.shape, generator=g, dtype=torch.float64), copy[0], x, return_dict=False)
    assert len(reads) > before
    # in-place edits of the handed-out tensor invalidate the shortcut
    sched = wrapper(schedule=scheduling.Scaled()) if wrapper is diffusers.RKUltraWrapperScheduler else wrapper(sampler=structured.DPM(order=2), schedule=scheduling.Scaled())
    sched.set_timesteps(6)
    edited = sched.timesteps
    edited += 0
    before = len(reads)
    sched.step(torch.randn(x.shape, generator=g, dtype=torch.float64), edited[0], x, return_dict=False)
    assert len(reads) > before


@pytest.mark.gpu
@pytest.mark.parametrize(("steps", "begin"), [(10, 5), (11, 6)])
@pytest.mark.parametrize("schedule", [scheduling.Sinner(scheduling.Linear()), scheduling.Scaled()], ids=["sinner", "scaled"])
def test_cuda_wrapper_brownian(steps: int, begin: int, schedule: scheduling.SkrampleSchedule) -> None:
    """The reference's test_diffusers_brownian (tests/self_sampling.py:503-537) with CUDA generators: the wrapper
    builds one Brownian generator per batch item, its noise comes from skr_noise_brownian (no torchsde), and because
    the noise is a function of the step the whole trajectory repeats bit for bit with fresh generators of the same
    seeds."""
    from skrample_b200 import native

    finals = []
    for _ in range(2):
        wrapper = diffusers.SkrampleWrapperScheduler(
            sampler=structured.Euler(stochasticity=1), schedule=schedule, model=models.DataModel(), compute_scale=torch.float32, noise_type=noise.Brownian
        )
        wrapper.set_timesteps(steps, device="cuda")
        wrapper.set_begin_index(begin * wrapper.order)
        generators = [torch.Generator(device="cuda").manual_seed(42), torch.Generator(device="cuda").manual_seed(43)]
        source = torch.Generator().manual_seed(0)
        x = torch.randn([2, 16, 128], generator=source).cuda()
        before = native.launch_count_kind(2)
        for t in wrapper.timesteps[begin * wrapper.order :]:
            x = wrapper.step(torch.randn([2, 16, 128], generator=source).cuda(), t, x, return_dict=False, generator=generators)[0]
        assert steps - begin <= native.launch_count_kind(2) - before < 2 * (steps - begin), "one Brownian launch per step for the whole batch"
        assert wrapper._noise_generator is not None
        assert len(wrapper._noise_generator.generators) == 2
        assert all(isinstance(g, noise.Brownian) for g in wrapper._noise_generator.generators)
        assert torch.isfinite(x).all()
        finals.append(x)
    assert torch.equal(finals[0], finals[1])
