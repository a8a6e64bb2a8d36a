"""GPU parity: the fused sm_100a step kernel (through the C ABI) vs golden tensors and the oracle."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest
import torch

import cases
import oracle_run
from oracle import skrample_oracle as O
from test_host_layer import run_product

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
STRUCTURED = np.load(GOLDEN / "structured.npz")
STRUCTURED_INDEX = json.loads((GOLDEN / "structured.json").read_text())


def test_native_library_is_the_path() -> None:
    from skrample_b200 import native
    from skrample_b200.common import Point

    before = native.launch_count()
    x = torch.randn(4096, device="cuda")
    n = torch.randn(4096, device="cuda")
    got = Point(10.0, 0.6, 0.8).add_noise(x, n)
    assert native.launch_count() == before + 1
    want = (x.cpu().numpy() * np.float32(0.8)) + (n.cpu().numpy() * np.float32(0.6))
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.fixture(params=["block", "block-generic", "interpreter"])
def kernel_kind(request: pytest.FixtureRequest, monkeypatch: pytest.MonkeyPatch) -> int:
    """Run a test once per kernel: the structured block kernel with its pinned-shape instantiations (default), the
    generic block-kernel instantiations only, and the general interpreter."""
    monkeypatch.delenv("SKR_FORCE_INTERP", raising=False)
    monkeypatch.delenv("SKR_NO_PINNED", raising=False)
    from skrample_b200 import native

    if request.param == "interpreter":
        monkeypatch.setenv("SKR_FORCE_INTERP", "1")
    if request.param == "block-generic":
        monkeypatch.setenv("SKR_NO_PINNED", "1")
    native.reset_switches()  # the library reads its switches once; plans made under other switches are dropped
    return 1 if request.param == "interpreter" else 0


@pytest.mark.parametrize("case", STRUCTURED_INDEX, ids=lambda c: c["id"])
def test_cuda_matches_reference_golden(case: dict, kernel_kind: int) -> None:
    "fp32 and fp64 trajectories are bit-identical to the reference's torch-CPU results, on both kernels."
    from skrample_b200 import native

    before = native.launch_count_kind(kernel_kind)
    other = native.launch_count_kind(1 - kernel_kind)
    result = run_product(case, device="cuda")
    assert native.launch_count_kind(kernel_kind) > before
    if not cases.composite(case):
        assert native.launch_count_kind(1 - kernel_kind) == other, "a step fell off the expected kernel"
    for field in ("final", "sample", "prediction"):
        cases.assert_matches(getattr(result, field).cpu().numpy(), STRUCTURED[f"{case['id']}/{field}"], case, field)


SIZES = [1, 3, 1023, 1024, 1025, 4096 + 17, 148 * 2 * 1024 * 3 + 5]


def _size_case(sampler: str, kw: dict, numel: int, dtype: str = "f32", schedule: str = "scaled", model: str = "NoiseModel") -> dict:
    return {"id": "x", "sampler": sampler, "kw": kw, "schedule": schedule, "model": model, "dtype": dtype, "steps": 5, "seed": numel + 7, "numel": numel}


@pytest.mark.parametrize("numel", SIZES)
@pytest.mark.parametrize(("sampler", "kw"), [("Euler", {"stochasticity": 1}), ("Adams", {"order": 4}), ("UniPC", {"order": 3, "stochasticity": 1}), ("DPM", {"order": 3})])
def test_sizes_and_tails_vs_oracle(sampler: str, kw: dict, numel: int, kernel_kind: int) -> None:
    "Empty-ish, ragged and multi-wave sizes: TMA tiles, the guarded tail and the persistent loop."
    case = _size_case(sampler, kw, numel)
    got = run_product(case, device="cuda")
    want = oracle_run.run_structured(case)
    assert np.array_equal(got.final.cpu().numpy(), want.final)
    assert np.array_equal(got.sample.cpu().numpy(), np.asarray(want.sample))
    assert np.array_equal(got.prediction.cpu().numpy(), np.asarray(want.prediction))


def test_empty_tensor() -> None:
    from skrample_b200 import scheduling
    from skrample_b200.sampling import models, structured

    x = torch.empty(0, device="cuda")
    out = structured.Euler().sample(x, x, (0.0, 0.1), models.FlowModel(), scheduling.Linear())
    assert out.final.shape == (0,)


@pytest.mark.parametrize("offset", [1, 2, 3])
def test_unaligned_views_vs_oracle(offset: int) -> None:
    "Tensors whose base is not 16-byte aligned take the element-wise path and still match bit for bit."
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    n = 5000
    rng = np.random.default_rng(offset)
    host = [rng.standard_normal(n + 8).astype(np.float32) for _ in range(3)]
    dev = [torch.from_numpy(h).cuda()[offset : offset + n] for h in host]
    step = Step.from_int(3, 10)
    got = structured.Euler(stochasticity=1).sample(dev[0], dev[1], step, models.NoiseModel(), scheduling.Scaled(), dev[2])
    cur = O.Rec(host[0][offset : offset + n], host[1][offset : offset + n], O.St(*step), host[2][offset : offset + n])
    want = O.euler_step(cur, O.Model("noise"), O.scaled(), 1)
    assert np.array_equal(got.final.cpu().numpy(), want)


def test_non_contiguous_inputs() -> None:
    from skrample_b200 import scheduling
    from skrample_b200.sampling import models, structured

    x = torch.randn(64, 96, device="cuda").t()
    o = torch.randn(96, 64, device="cuda")
    got = structured.Euler().sample(x, o, (0.2, 0.3), models.FlowModel(), scheduling.Linear())
    want = structured.Euler().sample(x.cpu(), o.cpu(), (0.2, 0.3), models.FlowModel(), scheduling.Linear())
    assert got.final.shape == (96, 64)
    assert torch.equal(got.final.cpu(), want.final)


@pytest.mark.parametrize(("sampler", "kw"), [("Euler", {"stochasticity": 1}), ("DPM", {"order": 2}), ("Adams", {"order": 9, "stochasticity": 1}), ("UniPC", {"order": 3, "stochasticity": 1}), ("SPC", {})])
@pytest.mark.parametrize("half", ["bf16", "f16"])
@pytest.mark.parametrize("numel", [3 * 1024 + 11, 2048 + 3, 1024 + 8, 5], ids=lambda n: f"n{n}")  # tails: bulk + leftover, leftover only, bulk only, no full tile
def test_half_storage_fp32_compute(sampler: str, kw: dict, half: str, numel: int) -> None:
    """16-bit storage: inputs are read as stored, all arithmetic is fp32 in the reference's op order, results are
    rounded ONCE to the storage type.  Against the same model in the oracle this is exact (0 ulp); it is the
    numerics of the reference's diffusers wrapper with compute_scale=float32 (reference: diffusers.py:575-599).
    Stated bound vs a pure-fp32 run: <= 2^-8 relative per step for bf16 (one rounding of the result)."""
    tdtype = {"bf16": torch.bfloat16, "f16": torch.float16}[half]
    case = _size_case(sampler, kw, numel, "f32", "flow", "FlowModel")

    def to_half_exact(a: np.ndarray) -> np.ndarray:
        return torch.from_numpy(a).to(tdtype).to(torch.float32).numpy()

    need_noise, need_prev = oracle_run.require(case)
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    sampler_obj = cases.make_sampler(structured, models, case)
    x0, outs, noises = cases.trajectory_inputs(case)
    x_host = to_half_exact(x0.astype(np.float32))
    x_dev = torch.from_numpy(x_host).cuda().to(tdtype)
    prev_dev: list = []
    prev_host: list[O.Rec] = []
    sch_o, model_o = oracle_run.schedule("flow"), oracle_run.MODELS["FlowModel"]
    for n in range(case["steps"]):
        out_host = to_half_exact(outs[n].astype(np.float32))
        noise_host = to_half_exact(noises[n].astype(np.float32)) if need_noise else None
        res = sampler_obj.sample(
            x_dev,
            torch.from_numpy(out_host).cuda().to(tdtype),
            Step.from_int(n, case["steps"]),
            models.FlowModel(),
            scheduling.FlowShift(scheduling.Linear()),
            torch.from_numpy(noise_host).cuda().to(tdtype) if need_noise else None,
            prev_dev,
        )
        prev_dev.append(res)
        prev_dev = prev_dev[max(len(prev_dev) - sampler_obj.require_previous, 0) :]
        rec = oracle_run.one_step(case, O.Rec(x_host, out_host, O.St.from_int(n, case["steps"]), noise_host), prev_host, model_o, sch_o)
        prev_host.append(rec)
        prev_host = prev_host[max(len(prev_host) - need_prev, 0) :]
        want_final = to_half_exact(np.asarray(rec.final, dtype=np.float32))
        assert res.final.dtype == tdtype
        assert np.array_equal(res.final.float().cpu().numpy(), want_final), f"step {n}"
        x_dev, x_host = res.final, want_final
        rec.final = want_final


def test_full_size_flux_latent_subset_vs_oracle() -> None:
    """BASELINE config 3 shape (16x16x128x128): the step is elementwise, so the oracle on a random subset of
    elements must equal the kernel's output at those elements, and a sharded run must equal the whole run."""
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    shape = (16, 16, 128, 128)
    g = torch.Generator(device="cuda").manual_seed(5)
    steps = 6
    sampler = structured.Adams(order=4, stochasticity=1)
    schedule, model = scheduling.FlowShift(scheduling.Linear()), models.FlowModel()
    x = torch.randn(shape, device="cuda", generator=g)
    pick = torch.randint(0, x.numel(), (4096,), device="cuda", generator=g)
    prev: list = []
    prev_lo: list = []
    prev_o: list[O.Rec] = []
    x_o = x.flatten()[pick].cpu().numpy()
    case = {"sampler": "Adams", "kw": {"order": 4, "stochasticity": 1}}
    for n in range(steps):
        out = torch.randn(shape, device="cuda", generator=g) * 0.5
        noise = torch.randn(shape, device="cuda", generator=g)
        step = Step.from_int(n, steps)
        res = sampler.sample(x, out, step, model, schedule, noise, prev)
        lo = sampler.sample(x[:8], out[:8], step, model, schedule, noise[:8], prev_lo)  # a 2-way batch shard
        assert torch.equal(lo.final, res.final[:8])
        prev = (prev + [res])[-sampler.require_previous :]
        prev_lo = (prev_lo + [lo])[-sampler.require_previous :]
        rec = oracle_run.one_step(
            case,
            O.Rec(x_o, out.flatten()[pick].cpu().numpy(), O.St(*step), noise.flatten()[pick].cpu().numpy()),
            prev_o,
            O.Model("flow"),
            O.flow_shift(O.linear()),
        )
        prev_o = (prev_o + [rec])[-3:]
        assert np.array_equal(res.final.flatten()[pick].cpu().numpy(), rec.final), f"step {n}"
        x, x_o = res.final, rec.final


def test_more_than_2_31_elements_subset_vs_oracle(kernel_kind: int) -> None:
    """Maximum sizes: a latent batch of 2^31 + 3077 elements (64-bit element indices, > 2^21 tiles, ragged tail).  The
    step is elementwise, so the oracle on elements around 0, around 2^31 and at the very end must match exactly."""
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    numel = 2**31 + 3 * 1024 + 5
    if torch.cuda.mem_get_info()[0] < 70 * 2**30:
        pytest.skip("needs ~60 GB of free device memory")
    edges = [0, 2**31 - 2048, numel - 4096]
    pick = torch.cat([torch.arange(e, e + 4096, device="cuda") for e in edges])
    sampler = structured.Adams(order=2, stochasticity=1)
    schedule, model = scheduling.Scaled(), models.NoiseModel()
    case = {"sampler": "Adams", "kw": {"order": 2, "stochasticity": 1}}
    x = torch.empty(numel, device="cuda", dtype=torch.bfloat16).normal_()
    x_o = x[pick].float().cpu().numpy()
    prev: list = []
    prev_o: list[O.Rec] = []
    for n in range(3):  # step 0: all-bf16 inputs (8 elements per thread); steps 1-2: fp32 history next to bf16 latents
        out = torch.empty(numel, device="cuda", dtype=torch.bfloat16).normal_()
        noise = torch.empty(numel, device="cuda", dtype=torch.bfloat16).normal_()
        step = Step.from_int(n, 5)
        res = sampler.sample(x, out, step, model, schedule, noise, prev)
        assert res.final.dtype == torch.bfloat16 and res.final.numel() == numel
        rec = oracle_run.one_step(
            case, O.Rec(x_o, out[pick].float().cpu().numpy(), O.St(*step), noise[pick].float().cpu().numpy()), prev_o, O.Model("noise"), O.scaled()
        )
        want = O.round_bf16(rec.final)
        got = res.final[pick].float().cpu().numpy()
        assert np.array_equal(got, want), f"step {n}: {np.flatnonzero(got != want)[:8]}"
        prev = (prev + [res])[-sampler.require_previous :]
        prev_o = (prev_o + [rec])[-1:]
        x, x_o = res.final, want
        del out, noise


@pytest.mark.parametrize("generator", ["Pyramid", "Colored"])
def test_full_size_video_latent_adams9_subset_vs_oracle(generator: str) -> None:
    """BASELINE config 4, one GPU's shard (1x16x21x90x160, bf16 storage, Adams-9 SDE, Flow, Pyramid / Colored noise):
    the oracle on a random subset of elements equals the kernel's output there, through the order ramp 1..9."""
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.pytorch import noise as sk_noise
    from skrample_b200.sampling import models, structured

    shape = (1, 16, 21, 90, 160)
    g = torch.Generator(device="cuda").manual_seed(17)
    steps = 12
    sampler = structured.Adams(order=9, stochasticity=1)
    schedule, model = scheduling.FlowShift(scheduling.Linear()), models.FlowModel()
    source = getattr(sk_noise, generator).from_inputs(shape[1:], torch.Generator(device="cuda").manual_seed(5), dtype=torch.float32)
    x = torch.randn(shape, device="cuda", generator=g).bfloat16()
    pick = torch.randint(0, x.numel(), (8192,), device="cuda", generator=g)
    prev: list = []
    prev_o: list[O.Rec] = []
    x_o = x.flatten()[pick].float().cpu().numpy()
    case = {"sampler": "Adams", "kw": {"order": 9, "stochasticity": 1}}
    for n in range(steps):
        out = (torch.randn(shape, device="cuda", generator=g) * 0.5).bfloat16()
        noise = source.generate(Step.from_int(n, steps)).unsqueeze(0).bfloat16()
        assert abs(noise.float().std().item() - 1.0) < 2e-2
        step = Step.from_int(n, steps)
        res = sampler.sample(x, out, step, model, schedule, noise, prev)
        prev = (prev + [res])[-sampler.require_previous :]
        rec = oracle_run.one_step(
            case,
            O.Rec(x_o, out.flatten()[pick].float().cpu().numpy(), O.St(*step), noise.flatten()[pick].float().cpu().numpy()),
            prev_o,
            O.Model("flow"),
            O.flow_shift(O.linear()),
        )
        prev_o = (prev_o + [rec])[-8:]
        want = O.round_bf16(rec.final)
        got = res.final.flatten()[pick].float().cpu().numpy()
        assert np.array_equal(got, want), f"step {n}: {int((got != want).sum())} of {got.size} differ"
        x, x_o = res.final, want


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("numel", [5, 1023, 1024 + 8, 3 * 1024 + 11, 148 * 4 * 1024 + 1001], ids=lambda n: f"n{n}")
def test_ragged_sizes_never_write_past_the_end(numel: int, dtype: torch.dtype, kernel_kind: int) -> None:
    """Outputs handed to the C ABI as slices of sentinel-filled buffers: whatever path the ragged last tile takes
    (staged partial tile, vector threads, element-wise thread), nothing outside [0, numel) is written."""
    import ctypes

    from skrample_b200 import native
    from skrample_b200.sampling import program as pg

    pad = 64  # elements of sentinel on both sides; keeps the slices 16-byte aligned
    g = torch.Generator(device="cuda").manual_seed(numel)
    x, y, z = (torch.randn(numel, device="cuda", generator=g).to(dtype) for _ in range(3))
    program = pg.Program()
    program.load(pg.X, x)
    program.conv(pg.ConvSpec(pg.CONV_USE_X | pg.CONV_MUL_Y | pg.CONV_DIV, 0.0, 0.7, 1.3), y)
    program.store(pg.P, "compute")
    program.fwd(0.9, 0.2, pred=pg.P, noise=z, zeta=0.3)
    program.store(pg.R)
    sentinel = 12345.0
    buffers = [torch.full((numel + 2 * pad,), sentinel, device="cuda", dtype=torch.float32), torch.full((numel + 2 * pad,), sentinel, device="cuda", dtype=dtype)]
    outs = [b[pad : pad + numel] for b in buffers]
    packed = native.pack_program(program, [x, y, z], outs)
    status = native.load().skr_program_launch(ctypes.byref(packed), numel, native.raw_stream())
    native.check(status, "skr_program_launch")
    torch.cuda.synchronize()
    for buffer in buffers:
        assert torch.all(buffer[:pad] == sentinel) and torch.all(buffer[pad + numel :] == sentinel), "wrote outside the tensor"
    xc, yc, zc = x.cpu().float(), y.cpu().float(), z.cpu().float()  # torch's CPU ops are the reference's arithmetic
    xhat = (xc - yc * 0.7) / 1.3
    assert torch.equal(outs[0].cpu(), xhat)
    want = ((0 + xc * 0.9) + xhat * 0.2) + zc * 0.3
    assert torch.equal(outs[1].cpu(), want.to(dtype))


@pytest.mark.parametrize("case", [c for c in STRUCTURED_INDEX if c["dtype"] == "f32"], ids=lambda c: c["id"])
def test_plan_cache_hits_match_reference_golden(case: dict) -> None:
    """The path every real loop runs: the second trajectory on the SAME sampler / schedule / model objects is served
    by the step-plan cache (no program is emitted, pointers are bound into the remembered launch).  Its results are
    compared with the reference's golden tensors, not with the first run."""
    from skrample_b200 import scheduling
    from skrample_b200.sampling import models, plan, structured

    objects = (cases.make_sampler(structured, models, case), cases.make_schedule(scheduling, case["schedule"]), cases.make_model(models, case["model"]))
    plan.clear()
    run_product(case, device="cuda", objects=objects)
    hits, misses = plan.stats()
    result = run_product(case, device="cuda", objects=objects)
    hits_after, misses_after = plan.stats()
    if not cases.composite(case):
        assert hits_after - hits >= case["steps"], f"expected every step of the second run to hit: {hits_after - hits} hits, {misses_after - misses} misses"
        assert misses_after == misses
    for field in ("final", "sample", "prediction"):
        cases.assert_matches(getattr(result, field).cpu().numpy(), STRUCTURED[f"{case['id']}/{field}"], case, field)


def test_config1_unipc3_sde_sdxl_bf16_subset_vs_oracle() -> None:
    """BASELINE.json configs[1] exactly: UniPC(order=3, stochasticity=1), Scaled, NoiseModel, 8x4x128x128 latent in bf16
    storage (fp32 compute and solver state), 25 steps.  The oracle on a random subset of elements equals the kernel there
    (0 ulp: same fp32 op order, one rounding to bf16).  Walked twice on the same objects: the second walk is plan hits."""
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, plan, structured

    shape, steps = (8, 4, 128, 128), 25
    sampler = structured.UniPC(order=3, stochasticity=1)
    schedule, model = scheduling.Scaled(), models.NoiseModel()
    case = {"sampler": "UniPC", "kw": {"order": 3, "stochasticity": 1}}
    sch_o, model_o = O.scaled(), O.Model("noise")
    g = torch.Generator(device="cuda").manual_seed(31)
    pick = torch.randint(0, 8 * 4 * 128 * 128, (8192,), device="cuda", generator=g)
    plan.clear()
    for walk in range(2):
        hits0 = plan.stats()[0]
        x = (torch.randn(shape, device="cuda", generator=g) * schedule.schedule(steps)[0].sigma).bfloat16()
        x_o = x.flatten()[pick].float().cpu().numpy()
        prev: list = []
        prev_o: list[O.Rec] = []
        for n in range(steps):
            out = torch.randn(shape, device="cuda", generator=g).bfloat16()
            noise = torch.randn(shape, device="cuda", generator=g).bfloat16()
            step = Step.from_int(n, steps)
            res = sampler.sample(x, out, step, model, schedule, noise, prev)
            assert res.final.dtype == torch.bfloat16 and res.prediction.dtype == torch.float32
            prev = (prev + [res])[-sampler.require_previous :]
            rec = oracle_run.one_step(
                case, O.Rec(x_o, out.flatten()[pick].float().cpu().numpy(), O.St(*step), noise.flatten()[pick].float().cpu().numpy()), prev_o, model_o, sch_o
            )
            prev_o = (prev_o + [rec])[-4:]
            want = O.round_bf16(rec.final)
            got = res.final.flatten()[pick].float().cpu().numpy()
            assert np.array_equal(got, want), f"walk {walk} step {n}: {int((got != want).sum())} of {got.size} differ"
            assert np.array_equal(res.prediction.flatten()[pick].cpu().numpy(), np.asarray(rec.prediction)), f"x-hat, walk {walk} step {n}"
            if n:
                assert np.array_equal(res.sample.flatten()[pick].cpu().numpy(), np.asarray(rec.sample)), f"corrected sample, walk {walk} step {n}"
            x, x_o = res.final, want
            rec.final = want
        if walk:
            assert plan.stats()[0] - hits0 == steps, "the second walk must be served by the plan cache"


def test_config2_rkultra4_flux_bf16_subset_vs_oracle() -> None:
    """BASELINE.json configs[2] exactly: RKUltra(order=4) (= RK2.EES7_MIN, 4 stages), FlowShift(Linear(), 3), FlowModel,
    16x16x128x128 latent in bf16 storage.  Stage inputs and results are rounded once to bf16 (they are the tensors the
    network sees / the caller gets), derivatives stay fp32: the oracle with that storage rule on a random subset of
    elements equals the kernels there.  First, interior and last step (the last one has the sigma = 0 backward stage);
    taken twice, the second time from the recorded launch script."""
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import functional, models

    shape, steps = (16, 16, 128, 128), 25
    sampler = functional.RKUltra(order=4)
    schedule, model_transform = scheduling.FlowShift(scheduling.Linear(), shift=3.0), models.FlowModel()
    tab = O.ees27_tableau(1 / 14 * (5 - 3 * 2**0.5))  # RK2.EES7_MIN, reference: functional.py:22, providers.py:207
    sch_o, model_o = O.flow_shift(O.linear(), 3.0), O.Model("flow")
    g = torch.Generator(device="cuda").manual_seed(77)
    numel = 16 * 16 * 128 * 128
    pick = torch.randint(0, numel, (8192,), device="cuda", generator=g)
    functional._scripts.known.clear()
    for walk in range(2):
        x = torch.randn(shape, device="cuda", generator=g).bfloat16()
        x_o = x.flatten()[pick].float().cpu().numpy()
        for n in (0, 1, 12, 24):
            table = [(torch.randn(shape, device="cuda", generator=g) * 0.5).bfloat16() for _ in range(4)]
            calls, calls_o = iter(range(4)), iter(range(4))
            step = Step.from_int(n, steps)
            got = sampler.step(x, lambda s, t, sigma, alpha: table[next(calls)], model_transform, schedule, step)
            assert got.dtype == torch.bfloat16
            want = O.step_tableau(
                tab, x_o, lambda s, t, sigma, alpha: table[next(calls_o)].flatten()[pick].float().cpu().numpy(), model_o, sch_o, O.St(*step), O.DATA, store=O.round_bf16
            )[0]
            have = got.flatten()[pick].float().cpu().numpy()
            assert np.array_equal(have, want), f"walk {walk} step {n}: {int((have != want).sum())} of {have.size} differ"
            x, x_o = got, want
        if walk == 0:
            assert functional._scripts.known, "the first walk should have recorded its launch scripts"


CONTRACTED = [c for c in STRUCTURED_INDEX if c["dtype"] == "f32" and c["sampler"] in ("UniPC", "UniP", "Adams", "SPC") and cases.tolerance(c) is None and not cases.composite(c)]


def test_contracted_arithmetic_is_opt_in_and_inside_the_stated_tolerance() -> None:
    """``native.set_arithmetic("contracted")``: the divided-difference / weighted-sum steps run kernels that fuse
    a*b + c and multiply by reciprocals.  Against the reference's golden trajectories (12 steps) every field stays within
    1e-5 of the tensor's scale - the per-STEP tolerance of the specification, here spent on a whole trajectory - the
    results are NOT the exact path's bits (so the contracted kernels really ran), and switching back restores them."""
    from skrample_b200 import native

    assert native.get_arithmetic() == "exact"
    differing = 0
    try:
        native.set_arithmetic("contracted")
        for case in CONTRACTED:
            result = run_product(case, device="cuda")
            for field in ("final", "sample", "prediction"):
                want = STRUCTURED[f"{case['id']}/{field}"]
                got = getattr(result, field).cpu().numpy()
                assert got.dtype == want.dtype and np.isfinite(got).all()
                worst = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
                assert worst <= 1e-5 * float(np.abs(want).max()), f"{case['id']} {field}: {worst:.3e}"
                differing += int(not np.array_equal(got, want))
    finally:
        native.set_arithmetic("exact")
    assert differing > len(CONTRACTED), "contracted mode produced the exact path's bits: its kernels did not run"
    case = CONTRACTED[0]
    result = run_product(case, device="cuda")
    assert np.array_equal(result.final.cpu().numpy(), STRUCTURED[f"{case['id']}/final"])


def test_contracted_shapes_are_selected_only_on_request() -> None:
    from skrample_b200 import native, scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    def shape_of_steady_step() -> str:
        sampler = structured.UniPC(order=3, stochasticity=1)
        schedule, model = scheduling.Scaled(), models.NoiseModel()
        x = torch.randn(4096, device="cuda").bfloat16()
        previous: list = []
        names = []
        real = native.CompiledProgram.specialise

        def spy(self, signature):  # noqa: ANN001, ANN202
            plan = real(self, signature)
            names.append(plan.shape_name)
            return plan

        native.CompiledProgram.specialise = spy
        try:
            for n in range(5):
                res = sampler.sample(x, torch.randn(4096, device="cuda").bfloat16(), Step.from_int(n, 10), model, schedule, torch.randn(4096, device="cuda").bfloat16(), previous)
                previous = (previous + [res])[-sampler.require_previous :]
                x = res.final
        finally:
            native.CompiledProgram.specialise = real
        return names[-1]

    assert shape_of_steady_step() == "unipc3/bf16"
    try:
        native.set_arithmetic("contracted")
        assert shape_of_steady_step() == "unipc3~contracted/bf16"
    finally:
        native.set_arithmetic("exact")


@pytest.mark.parametrize("name", ["UniPC(order=3,stochasticity=1)|scaled|NoiseModel|f32|12", "SPC()|flow|FlowModel|f32|12"])
def test_oversized_step_programs_fall_back_to_composed_launches(name: str, monkeypatch: pytest.MonkeyPatch) -> None:
    """A predictor-corrector step that does not fit one launch's op table (SKR_MAX_OPS; e.g. SPC over two high-order
    UniPCs on a long schedule) runs as separately fused launches in the reference's order instead of raising: forced
    here by shrinking the limit, the trajectory still equals the reference's golden tensors bit for bit."""
    from skrample_b200 import native
    from skrample_b200.sampling import plan
    from skrample_b200.sampling import program as pg

    case = next(c for c in STRUCTURED_INDEX if c["id"] == name)
    monkeypatch.setattr(pg, "MAX_OPS", 9)
    plan.clear()
    before = native.launch_count()
    result = run_product(case, device="cuda")
    assert native.launch_count() - before > case["steps"], "expected more than one launch per step"
    for field in ("final", "sample", "prediction"):
        cases.assert_matches(getattr(result, field).cpu().numpy(), STRUCTURED[f"{case['id']}/{field}"], case, field)
