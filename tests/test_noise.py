"""Noise generators: CPU-generator stream parity with the reference recipe, oracle pinning, CUDA kernels."""

from __future__ import annotations

import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import skrample_oracle as O
from skrample_b200.common import Step
from skrample_b200.pytorch import noise

# ------------------------------------------------------------------------------------------- CPU (no GPU needed)


def test_random_cpu_generator_is_torch_randn() -> None:
    "reference: noise.py:36-42,73-74 - a CPU generator keeps its own stream, so seeds reproduce the reference."
    g = torch.Generator().manual_seed(3)
    want = torch.randn((2, 8, 8), generator=torch.Generator().manual_seed(3))
    assert torch.equal(noise.Random.from_inputs((2, 8, 8), g).generate(None), want)


def test_offset_cpu_draw_order() -> None:
    "reference: noise.py:104-113 - offset first (reduced shape), then the full draw."
    g = torch.Generator().manual_seed(5)
    ref = torch.Generator().manual_seed(5)
    off = torch.randn((4, 1, 1), generator=ref) * 0.2**2
    want = torch.randn((4, 8, 8), generator=ref) + off
    assert torch.equal(noise.Offset.from_inputs((4, 8, 8), g).generate(None), want)


def _record_pyramid(shape: tuple[int, ...], props: noise.PyramidProps, seed: int) -> tuple[torch.Tensor, np.ndarray, list[np.ndarray], list[float]]:
    "Run the CPU-generator pyramid while recording every draw it makes (the reference's draw order)."
    g = torch.Generator().manual_seed(seed)
    replay = torch.Generator().manual_seed(seed)
    out = noise.Pyramid.from_inputs(shape, g, props).generate(None)
    base = torch.randn(shape, generator=replay)
    mask = [n in [len(shape) + d if d < 0 else d for d in props.dims] for n in range(len(shape))]
    running = list(shape)
    levels, ratios = [], []
    for level in range(99):
        r = torch.rand([1], generator=replay).item() * 2 + 2
        ratios.append(r)
        running = [max(1, int(s / (r**level))) if m else s for m, s in zip(mask, running)]
        levels.append(torch.randn(running, generator=replay).numpy())
        if any(s <= 1 for m, s in zip(mask, running) if m):
            break
    return out, base.numpy(), levels, ratios


GOLDEN_NOISE = np.load(__import__("pathlib").Path(__file__).resolve().parent / "golden" / "noise.npz")
GOLDEN_PYRAMID = [((4, 64, 64), (-1, -2), 99), ((3, 40, 56), (-1, -2), 1), ((5, 96), (-1,), 99)]  # make_golden.NOISE_PYRAMID_CASES
GOLDEN_COLORED = [((4, 32, 32), 1.5, None), ((2, 1, 40, 31), -2.0, 3.0), ((4096,), 0.7, None)]  # make_golden.NOISE_COLORED_CASES


def _golden_pyramid(n: int) -> tuple[np.ndarray, np.ndarray, list[np.ndarray], list[float]]:
    levels = []
    while f"pyramid{n}/level{len(levels)}" in GOLDEN_NOISE:
        levels.append(GOLDEN_NOISE[f"pyramid{n}/level{len(levels)}"])
    return GOLDEN_NOISE[f"pyramid{n}/out"], GOLDEN_NOISE[f"pyramid{n}/base"], levels, GOLDEN_NOISE[f"pyramid{n}/ratios"].tolist()


@pytest.mark.parametrize("n", range(len(GOLDEN_PYRAMID)))
def test_oracle_pyramid_matches_reference_golden(n: int) -> None:
    "Pins oracle.pyramid_compose / level shapes to the reference's own output for the draws it consumed."
    shape, dims, depth = GOLDEN_PYRAMID[n]
    out, base, levels, ratios = _golden_pyramid(n)
    mask = [i in [len(shape) + d if d < 0 else d for d in dims] for i in range(len(shape))]
    assert [lv.shape for lv in levels] == O.pyramid_level_shapes(shape, mask, ratios)
    np.testing.assert_allclose(O.pyramid_compose(base, levels, mask, 0.3, depth), out, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("n", range(len(GOLDEN_PYRAMID)))
def test_pyramid_cpu_generator_matches_reference_golden(n: int) -> None:
    "The product's CPU-generator path reproduces the reference bit for bit (same seed, same draw order)."
    shape, dims, depth = GOLDEN_PYRAMID[n]
    got = noise.Pyramid.from_inputs(shape, torch.Generator().manual_seed(40 + n), noise.PyramidProps(dims=dims, depth=depth)).generate(None)
    assert np.array_equal(got.numpy(), GOLDEN_NOISE[f"pyramid{n}/out"])


@pytest.mark.parametrize("n", range(len(GOLDEN_COLORED)))
def test_colorize_matches_reference_golden(n: int) -> None:
    shape, exponent, energy = GOLDEN_COLORED[n]
    white = GOLDEN_NOISE[f"colored{n}/white"]
    want = GOLDEN_NOISE[f"colored{n}/out"]
    assert np.array_equal(noise.Colored.colorize_noise(torch.from_numpy(white), exponent, energy).numpy(), want)
    np.testing.assert_allclose(O.colorize(white, exponent, energy), want, rtol=1e-3, atol=2e-4)


@pytest.mark.parametrize(("shape", "dims"), [((4, 64, 64), (-1, -2)), ((3, 40, 56), (-1, -2)), ((5, 96), (-1,)), ((48, 4, 48), (0, 2))])
def test_oracle_pyramid_matches_reference_recipe(shape: tuple[int, ...], dims: tuple[int, ...]) -> None:
    "Pins oracle.pyramid_compose (the checker of the CUDA kernel) to the torch recipe of noise.py:146-207."
    props = noise.PyramidProps(dims=dims)
    out, base, levels, ratios = _record_pyramid(shape, props, 11)
    mask = [n in [len(shape) + d if d < 0 else d for d in dims] for n in range(len(shape))]
    assert [lv.shape for lv in levels] == O.pyramid_level_shapes(shape, mask, ratios)
    want = O.pyramid_compose(base, levels, mask, props.strength, props.depth)
    np.testing.assert_allclose(out.numpy(), want, rtol=2e-5, atol=2e-6)
    assert abs(float(out.std()) - 1) < 1e-5


@pytest.mark.parametrize(("shape", "exponent", "energy"), [((4, 32, 32), 1.5, None), ((2, 1, 40, 31), -2.0, 3.0), ((4096,), 0.7, None)])
def test_oracle_colorize_matches_torch_recipe(shape: tuple[int, ...], exponent: float, energy: float | None) -> None:
    "Pins oracle.colorize to noise.py:338-405 (FFT libraries differ, hence a tolerance)."
    white = torch.randn(shape, generator=torch.Generator().manual_seed(2))
    got = noise.Colored.colorize_noise(white, exponent, energy).numpy()
    want = O.colorize(white.numpy(), exponent, energy)
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-4)


@pytest.mark.parametrize("n", [0, 3, 6, 9])
def test_colored_exponent_schedule(n: int) -> None:
    "reference: noise.py:410-420 (defaults give 0.17, -0.16, -0.73, -2.0 over 10 steps, SURVEY appendix A.12)"
    gen = noise.Colored.from_inputs((8,), torch.Generator())
    got = gen.exponent(Step.from_int(n, 10))
    assert got == pytest.approx(O.colored_exponent(O.St.from_int(n, 10)), abs=0)
    assert got == pytest.approx({0: 0.17, 3: -0.16, 6: -0.73, 9: -2.0}[n], abs=6e-3)


def _has_torchsde() -> bool:
    try:
        import torchsde  # noqa: F401
    except ImportError:
        return False
    return True


BROWNIAN_STEPS = [
    (0.0, 1.0),
    (0.0, 0.04),
    (0.48, 0.52),  # separated at the root
    (0.96, 1.0),
    (0.25, 0.5),  # both ends on dyadic points
    (0.3, 0.3001),
    (0.7000001, 0.7000003),  # inside one leaf of a 17-level tree
    (1 / 3, 2 / 3),
]


@pytest.mark.skipif(_has_torchsde(), reason="with torchsde installed a CPU generator uses the reference's tree")
@pytest.mark.parametrize("step", BROWNIAN_STEPS)
@pytest.mark.parametrize("max_steps", [10_000, 50])
def test_brownian_cpu_generator_vs_oracle(step: tuple[float, float], max_steps: int) -> None:
    """Without torchsde a CPU generator evaluates the library's own bridge tree on the host (reference: noise.py:219-245
    would raise ImportError): relative float32 increments here, absolute float64 path values in the oracle."""
    shape = (3, 17, 23)
    g = noise.Brownian.from_inputs(shape, torch.Generator().manual_seed(77), noise.BrownianProps(max_steps=max_steps))
    got = g.generate(Step(*step))
    assert got.shape == shape and got.dtype == torch.float32 and got.device.type == "cpu"
    want = O.brownian_increment(77, step[0], step[1], math.ceil(math.log2(max_steps * 10)), got.numel()).reshape(shape)
    np.testing.assert_allclose(got.numpy().astype(np.float64), want, rtol=0, atol=2e-5)


@pytest.mark.skipif(_has_torchsde(), reason="with torchsde installed a CPU generator uses the reference's tree")
def test_brownian_cpu_generator_contract() -> None:
    shape = (4, 32, 32)
    a = noise.Brownian.from_inputs(shape, torch.Generator().manual_seed(3), dtype=torch.float64)
    b = noise.Brownian.from_inputs(shape, torch.Generator().manual_seed(3), dtype=torch.float64)
    s0, s1 = Step.from_int(3, 25), Step.from_int(4, 25)
    x0, x1 = a.generate(s0), a.generate(s1)
    assert x0.dtype == torch.float64
    assert torch.equal(b.generate(s1), x1) and torch.equal(b.generate(s0), x0) and torch.equal(a.generate(Step(s0.time_to, s0.time_from)), x0)
    both = a.generate(Step(s0.time_from, s1.time_to))
    assert ((x0 + x1) * math.sqrt(s0.distance()) - both * math.sqrt(2 * s0.distance())).abs().max().item() < 2e-6
    assert abs((x0 * x1).mean().item()) < 5 / math.sqrt(x0.numel())
    assert torch.equal(a.generate(None), torch.randn(shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64))
    with pytest.raises(ZeroDivisionError):
        a.generate(Step(0.5, 0.5))
    batch = noise.BatchTensorNoise.from_batch_inputs(noise.Brownian, shape, [torch.Generator().manual_seed(3), torch.Generator().manual_seed(4)], noise.BrownianProps(), torch.float64)
    got = batch.generate(s0)
    assert got.shape == (2, *shape) and torch.equal(got[0], x0) and not torch.equal(got[1], x0)


def test_same_device_rule() -> None:
    "An index-less generator device means the current device; types must agree."
    same = noise._same_device
    assert same(torch.device("cpu"), torch.device("cpu"))
    assert not same(torch.device("cpu"), torch.device("cuda", 0))
    assert same(torch.device("cuda", 1), torch.device("cuda", 1)) and not same(torch.device("cuda", 0), torch.device("cuda", 1))


def test_oracle_philox_known_answers() -> None:
    "Philox4x32-10 known-answer vectors published with Random123 (kat_vectors: zeros, ones, pi digits)."
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for counter, key, want in cases:
        got = O.philox4x32_10(np.array([counter], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want


def test_oracle_brownian_contract() -> None:
    """The contract noise.py:244-245 relies on: unit-variance increments that add up over adjoining steps, are
    independent over disjoint ones, and depend on (seed, t0, t1) only."""
    n, depth = 1 << 16, 17
    a = O.brownian_increment(5, 0.20, 0.24, depth, n)
    b = O.brownian_increment(5, 0.24, 0.28, depth, n)
    ab = O.brownian_increment(5, 0.20, 0.28, depth, n)
    assert np.array_equal(a, O.brownian_increment(5, 0.20, 0.24, depth, n))
    assert np.abs((a + b) * math.sqrt(0.04) - ab * math.sqrt(0.08)).max() < 1e-12
    for x in (a, b, ab, O.brownian_increment(5, 0.0, 1.0, depth, n), O.brownian_increment(5, 0.5 - 1e-4, 0.5 + 2e-4, depth, n)):
        assert abs(x.mean()) < 5 / math.sqrt(n) and abs(x.var() - 1) < 2e-2
        assert abs((x**4).mean() - 3) < 0.15
    assert abs(np.corrcoef(a, b)[0, 1]) < 5 / math.sqrt(n)
    assert abs(np.corrcoef(a, O.brownian_increment(6, 0.20, 0.24, depth, n))[0, 1]) < 5 / math.sqrt(n)
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 5 / math.sqrt(n)
    # overlapping steps are correlated by the shared part: corr = overlap / sqrt(len_a * len_b)
    c = O.brownian_increment(5, 0.22, 0.26, depth, n)
    assert np.corrcoef(a, c)[0, 1] == pytest.approx(0.5, abs=0.02)


# ------------------------------------------------------------------------------------------- CUDA kernels

gpu = pytest.mark.gpu


def _gen(seed: int = 1234) -> torch.Generator:
    return torch.Generator(device="cuda").manual_seed(seed)


@gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16, torch.float64])
def test_random_determinism_and_moments(dtype: torch.dtype) -> None:
    from skrample_b200 import native

    shape = (4, 16, 128, 128)
    before = native.launch_count_kind(2)
    a = noise.Random.from_inputs(shape, _gen(), dtype=dtype)
    b = noise.Random.from_inputs(shape, _gen(), dtype=dtype)
    a0, a1 = a.generate(None), a.generate(Step.from_int(0, 10))
    b0, b1 = b.generate(None), b.generate(Step.from_int(0, 10))
    assert native.launch_count_kind(2) - before == 4
    assert a0.dtype == dtype and a0.is_cuda and a0.shape == shape
    assert torch.equal(a0, b0) and torch.equal(a1, b1), "same seed must give identical bits"
    assert not torch.equal(a0, a1), "successive draws must differ"
    x = a0.double().flatten()
    n = x.numel()
    assert abs(x.mean().item()) < 5 / math.sqrt(n)
    assert abs(x.var().item() - 1) < (2e-2 if dtype in (torch.bfloat16, torch.float16) else 6e-3)
    assert abs((x**3).mean().item()) < 2e-2  # skewness
    assert abs((x**4).mean().item() - 3) < 6e-2  # kurtosis
    assert abs((x[:-1] * x[1:]).mean().item()) < 5 / math.sqrt(n)  # lag-1 autocorrelation
    assert (x.abs() > 4).float().mean().item() < 2e-4  # tails present but sane
    assert x.abs().max().item() > 4


@gpu
def test_random_values_do_not_depend_on_sharding() -> None:
    "Per-item generators: generating items together or one by one gives the same bits (BatchTensorNoise)."
    unit = (16, 64, 64)
    seeds = [_gen(100 + i) for i in range(4)]
    batch = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, seeds).generate(None)
    assert batch.shape == (4, *unit)
    for i in range(4):
        alone = noise.Random.from_inputs(unit, _gen(100 + i)).generate(None)
        assert torch.equal(batch[i], alone)


@gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_batch_generate_into_equals_generate_then_cast(dtype: torch.dtype) -> None:
    "Writing the batch straight into a caller's buffer (graph-captured trajectories) gives generate().to(dtype)."
    unit = (4, 33, 31)
    a = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, [_gen(40 + i) for i in range(3)])
    b = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, [_gen(40 + i) for i in range(3)])
    for _ in range(3):  # consecutive draws stay in step
        out = torch.empty((3, *unit), device="cuda", dtype=dtype)
        b.generate_into(out, None)
        assert torch.equal(out, a.generate(None).to(dtype))
    mixed = noise.BatchTensorNoise([noise.Random.from_inputs(unit, _gen(1)), noise.Offset.from_inputs(unit, _gen(2))])
    twin = noise.BatchTensorNoise([noise.Random.from_inputs(unit, _gen(1)), noise.Offset.from_inputs(unit, _gen(2))])
    out = torch.empty((2, *unit), device="cuda", dtype=dtype)
    mixed.generate_into(out, None)  # not a plain Random batch: generic path
    assert torch.equal(out, twin.generate(None).to(dtype))


@gpu
@pytest.mark.parametrize("shape", [(4, 48, 64), (3, 5, 33, 31), (2, 40, 37), (6, 130), (2, 3, 4, 20, 24)])
@pytest.mark.parametrize("dims", [(-2, -1), (-1,), (-2,), (1, 2)])
def test_pyramid_grid_paths_agree_with_in_kernel_draws(shape: tuple[int, ...], dims: tuple[int, ...]) -> None:
    """The same pyramid three ways: every corner drawn inside the kernel (no scratch), grids + the generic composition
    kernel, grids + the four-wide kernel (last extent % 4 == 0).  Same values up to the last bit of the std."""
    if any(d >= len(shape) for d in dims if d > 0):
        pytest.skip("axis out of range for this shape")
    props = noise.PyramidProps(dims=dims)
    fast = noise.Pyramid.from_inputs(shape, _gen(77), props)
    slow = noise.Pyramid.from_inputs(shape, _gen(77), props)
    slow._use_grids = False
    for _ in range(2):
        a, b = fast.generate(None), slow.generate(None)
        assert torch.isfinite(a).all()
        torch.testing.assert_close(a, b, rtol=2e-6, atol=2e-6)
        assert abs(a.std().item() - 1.0) < 1e-3


@gpu
def test_ragged_and_unaligned_fill() -> None:
    big = torch.empty(1031 + 3, device="cuda")
    a = noise.Random.from_inputs((1031,), _gen(9))
    b = noise.Random.from_inputs((1031,), _gen(9))
    x = a.generate(None)
    b.generate_into(big[3:], None)  # misaligned view falls back to a copy
    assert torch.equal(x, big[3:])


@gpu
def test_offset_structure() -> None:
    "out - plain draw is constant along the non-kept axes and ~N(0, strength^4) along the kept one."
    shape = (64, 32, 32)
    props = noise.OffsetProps(dims=(0,), strength=0.5)
    off = noise.Offset.from_inputs(shape, _gen(7), props).generate(None)
    plain = noise.Random.from_inputs(shape, _gen(7)).generate(None)  # same seed, same main stream
    delta = off - plain
    per_slice = delta.reshape(64, -1)
    assert (per_slice.std(dim=1) < 1e-6).all(), "offset must be constant within a slice"
    assert abs(per_slice[:, 0].std().item() - 0.25) < 0.08
    static = noise.Offset.from_inputs(shape, _gen(7), noise.OffsetProps(static=True))
    d0 = static.generate(None)
    d1 = static.generate(None)
    assert not torch.equal(d0, d1)


def _pyramid_supplied(base: np.ndarray, levels: list[np.ndarray], mask: list[bool], strength: float, depth: int, fused: bool = False) -> torch.Tensor:
    """Drive skr_noise_pyramid with supplied draws (test hook of the C ABI).  ``fused``: give it the scratch areas of the
    fast path (composition into scratch + scale pass; trailing resized axes: levels stretched to the unit's width first)."""
    from skrample_b200 import native

    lib = noise._lib()
    dev = [torch.from_numpy(np.ascontiguousarray(lv, dtype=np.float32)).cuda() for lv in levels]
    base_dev = torch.from_numpy(np.ascontiguousarray(base, dtype=np.float32)).cuda()
    desc = noise._SkrPyramid()
    desc.ndim = base.ndim
    desc.n_levels = len(levels)
    desc.base_buffer = base_dev.data_ptr()
    top = len(levels) - 1
    skip = min(top, max(0, top - depth))
    for n, d in enumerate(base.shape):
        desc.shape[n] = d
        desc.masked[n] = int(mask[n])
    for l, lv in enumerate(levels):
        ext = [s for m, s in zip(mask, lv.shape) if m]
        desc.levels[l].buffer = dev[l].data_ptr()
        desc.levels[l].extent[0] = ext[0]
        desc.levels[l].extent[1] = ext[1] if len(ext) > 1 else 1
        desc.levels[l].weight = strength**l if l >= skip else 0.0
    out = torch.empty(base.shape, device="cuda")
    if fused:
        room = sum(int(np.prod(lv.shape)) // lv.shape[-1] * base.shape[-1] + 4 for lv in levels)
        scratch = torch.empty(base.size + 4 + room, device="cuda")
        desc.scratch = scratch.data_ptr()
        desc.levels_scratch = scratch.data_ptr() + 4 * ((base.size + 3) & ~3)
        desc.levels_scratch_floats = room
    moments = torch.zeros(native.MOMENTS_DOUBLES, dtype=torch.float64, device="cuda")
    native.check(lib.skr_noise_pyramid(out.data_ptr(), 0, ctypes.byref(desc), moments.data_ptr(), torch.cuda.current_stream().cuda_stream), "pyramid")
    return out


@gpu
@pytest.mark.parametrize(("shape", "dims", "depth"), [((4, 64, 64), (-1, -2), 99), ((3, 40, 56), (-1, -2), 1), ((5, 96), (-1,), 99), ((48, 4, 48), (0, 2), 99)])
@pytest.mark.parametrize("fused", [False, True], ids=["two-pass", "fused"])
def test_pyramid_kernel_vs_oracle_on_supplied_draws(shape: tuple[int, ...], dims: tuple[int, ...], depth: int, fused: bool) -> None:
    """Identical supplied draws -> the kernels' upsample / weights / std equal the oracle's (fp32 rounding only): the
    generic two-pass kernel, and the fused path (separable upsampling for trailing axes, four-wide otherwise)."""
    props = noise.PyramidProps(dims=dims, depth=depth)
    _out, base, levels, _ratios = _record_pyramid(shape, props, 21)
    mask = [n in [len(shape) + d if d < 0 else d for d in dims] for n in range(len(shape))]
    got = _pyramid_supplied(base, levels, mask, props.strength, depth, fused).cpu().numpy()
    want = O.pyramid_compose(base, levels, mask, props.strength, depth)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)


@gpu
@pytest.mark.parametrize("shape", [(4, 128, 128), (16, 21, 90, 160), (336, 90, 160)])
def test_pyramid_philox_statistics(shape: tuple[int, ...]) -> None:
    "In-kernel pyramid: unit std (1e-3), deterministic, low-frequency energy above white noise."
    a = noise.Pyramid.from_inputs(shape, _gen(5)).generate(None)
    b = noise.Pyramid.from_inputs(shape, _gen(5)).generate(None)
    assert torch.equal(a, b)
    assert abs(a.double().std().item() - 1) < 1e-3
    assert abs(a.double().mean().item()) < 0.05
    # coarse levels add spatially correlated energy: neighbouring pixels correlate positively
    corr = (a[..., :-1] * a[..., 1:]).double().mean().item()
    assert 0.01 < corr < 0.5


@gpu
@pytest.mark.parametrize(("shape", "exponent", "energy"), [((4, 32, 32), 1.5, None), ((2, 1, 40, 31), -2.0, 3.0), ((4096,), 0.7, None), ((16, 8, 30, 40), 0.25, None), ((8, 21, 45, 80), -1.0, None), ((3, 700, 37), 2.0, 0.5)])
def test_colorize_kernel_vs_oracle(shape: tuple[int, ...], exponent: float, energy: float | None) -> None:
    white = torch.randn(shape, generator=torch.Generator().manual_seed(2))
    got = noise.Colored.colorize_noise(white.cuda(), exponent, energy).cpu().numpy()
    want = O.colorize(white.numpy(), exponent, energy)
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=5e-4)


def _spectral_slope(data: np.ndarray) -> float:
    "Radially binned log-log PSD slope (same estimator idea as reference tests/self_noise.py:13-60)."
    spec = np.abs(np.fft.fftshift(np.fft.fftn(data))) ** 2
    freqs = [np.fft.fftshift(np.fft.fftfreq(s)) for s in data.shape]
    radius = np.sqrt(sum(m**2 for m in np.meshgrid(*freqs, indexing="ij")))
    keep = radius > 0
    r, p = radius[keep], spec[keep]
    bins = min(data.shape) // 2
    edges = np.linspace(r.min(), r.max(), bins + 1)
    which = np.digitize(r, edges) - 1
    centers = 0.5 * (edges[:-1] + edges[1:])
    power = np.array([p[which == i].mean() if (which == i).any() else 0 for i in range(bins)])
    ok = (power > 0) & (centers > 0)
    slope = np.polyfit(np.log(centers[ok]), np.log(power[ok]), 1)[0]
    return float(-slope)


@gpu
@pytest.mark.parametrize("exponent", [-3, -1.5, 0, 1.5, 3])
@pytest.mark.parametrize("shape", [(65536,), (512, 512), (64, 64, 64)])
def test_colored_generator_color_and_energy(exponent: float, shape: tuple[int, ...]) -> None:
    "Mirrors reference tests/self_noise.py:63-103 on the CUDA path: PSD slope within 0.1, std 1 (+-1e-2) or |energy|."
    gen = noise.Colored(shape, _gen(3), torch.float32, noise.ColoredProps(color_curve=0, color_start=exponent, color_end=-exponent))
    n0 = gen.generate(None)
    assert abs(exponent - _spectral_slope(n0.cpu().numpy())) < 0.1
    n1 = gen.generate(Step(0, 1))
    assert abs(-exponent - _spectral_slope(n1.cpu().numpy())) < 0.1
    assert abs(n0.std().item() - 1) < 1e-2
    fixed = noise.Colored(shape, _gen(3), torch.float32, noise.ColoredProps(energy=-1.5, color_start=exponent))
    assert abs(fixed.generate(None).std().item() - 1.5) < 1e-4


@gpu
@pytest.mark.parametrize("kernel", ["block", "interpreter"])
@pytest.mark.parametrize(("sampler_name", "kw"), [("Euler", {"stochasticity": 1}), ("UniPC", {"order": 3, "stochasticity": 1}), ("Adams", {"order": 4, "stochasticity": 1})])
def test_in_kernel_noise_equals_materialised_noise(sampler_name: str, kw: dict, kernel: str, monkeypatch: pytest.MonkeyPatch) -> None:
    """The noise term drawn inside the step kernel (PhiloxDraw) is bit-identical to supplying the tensor that
    skr_noise_fill writes for the same keys - for a batch of per-item generators, a ragged size, both kernels."""
    from skrample_b200 import native, scheduling
    from skrample_b200.sampling import models, structured

    if kernel == "interpreter":
        monkeypatch.setenv("SKR_FORCE_INTERP", "1")
    native.reset_switches()
    sampler = getattr(structured, sampler_name)(**kw)
    unit = (4, 33, 31)  # item_numel = 4092: multiple of 4 but not of the tile
    batch = 3
    schedule, model = scheduling.Scaled(), models.NoiseModel()
    results = []
    for lazy in (True, False):
        gens = [_gen(50 + i) for i in range(batch)]
        source = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, gens)
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn((batch, *unit), device="cuda", generator=g)
        previous: list = []
        fills_before = native.launch_count_kind(2)
        for n in range(5):
            out = torch.randn((batch, *unit), device="cuda", generator=g) * 0.5
            step = Step.from_int(n, 5)
            drawn = source.lazy(step)
            assert isinstance(drawn, noise.PhiloxDraw)
            res = sampler.sample(x, out, step, model, schedule, drawn if lazy else drawn.materialize(), previous)
            previous = (previous + [res])[-sampler.require_previous :] if sampler.require_previous else []
            x = res.final
        if lazy:
            assert native.launch_count_kind(2) == fills_before, "lazy noise must not launch a fill kernel"
        results.append(x)
    assert torch.equal(results[0], results[1])


@gpu
@pytest.mark.parametrize(("unit", "dims"), [((4, 33, 32), (0,)), ((4, 6, 10), (0, 1)), ((3, 7, 5), (0,)), ((8, 16), ()), ((2, 12), (0, 1))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kernel", ["block", "interpreter"])
def test_offset_batches_are_philox_keys_too(unit: tuple[int, ...], dims: tuple[int, ...], dtype: torch.dtype, kernel: str, monkeypatch: pytest.MonkeyPatch) -> None:
    """A batch of non-static Offset generators with leading kept axes (noise.py:84-113) is handed out as Philox keys:
    the materialised batch equals the per-item fill, and a step that draws it in its kernel equals the step that
    reads the tensor - bit for bit, with no fill launch."""
    from skrample_b200 import native, scheduling
    from skrample_b200.sampling import models, structured

    if kernel == "interpreter":
        monkeypatch.setenv("SKR_FORCE_INTERP", "1")
    native.reset_switches()
    props = noise.OffsetProps(dims=dims, strength=0.4)
    batch = 3
    source = noise.BatchTensorNoise.from_batch_inputs(noise.Offset, unit, [_gen(70 + i) for i in range(batch)], props, dtype)
    twins = [noise.Offset.from_inputs(unit, _gen(70 + i), props, dtype) for i in range(batch)]
    sampler, model, schedule = structured.Euler(stochasticity=1), models.FlowModel(), scheduling.Linear()
    x = torch.randn((batch, *unit), device="cuda").to(dtype)
    o = torch.randn((batch, *unit), device="cuda").to(dtype)
    for n in range(2):  # consecutive draws stay in step with the per-item generators
        drawn = source.auto(None)
        assert isinstance(drawn, noise.PhiloxDraw) and drawn.offset_inner > 0
        fills = native.launch_count_kind(2)
        a = sampler.sample(x, o, Step.from_int(n + 1, 9), model, schedule, drawn)
        assert native.launch_count_kind(2) == fills, "the offset draw must happen inside the step kernel"
        want = torch.stack([t.generate(None) for t in twins])
        assert torch.equal(drawn.materialize(), want)
        b = sampler.sample(x, o, Step.from_int(n + 1, 9), model, schedule, want)
        assert torch.equal(a.final, b.final)
    # the offset really is there: constant along the broadcast axes, different from the plain draw
    plain = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, [_gen(70 + i) for i in range(batch)], dtype=torch.float32)
    shifted = noise.BatchTensorNoise.from_batch_inputs(noise.Offset, unit, [_gen(70 + i) for i in range(batch)], props, torch.float32)
    delta = shifted.generate(None) - plain.generate(None)
    lead = len([d for d in range(len(unit)) if d in dims])
    rows = delta.reshape(batch * math.prod(unit[:lead]), -1)
    assert (rows.max(dim=1).values - rows.min(dim=1).values).abs().max().item() < 2e-6
    assert rows[:, 0].abs().max().item() > 0
    # static offsets and offsets along non-leading axes keep the per-item path
    static = noise.BatchTensorNoise.from_batch_inputs(noise.Offset, unit, [_gen(1), _gen(2)], noise.OffsetProps(dims=dims, static=True))
    assert isinstance(static.auto(None), torch.Tensor)
    if len(unit) > 1:
        trailing = noise.BatchTensorNoise.from_batch_inputs(noise.Offset, unit, [_gen(1), _gen(2)], noise.OffsetProps(dims=(len(unit) - 1,)))
        assert isinstance(trailing.auto(None), torch.Tensor)


@gpu
def test_in_kernel_noise_unaligned_items() -> None:
    "item_numel not a multiple of 4: per-element Philox indexing still matches the materialised tensor."
    from skrample_b200 import scheduling
    from skrample_b200.sampling import models, structured

    unit = (3, 7, 5)
    gens = [_gen(9), _gen(10)]
    source = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, gens)
    drawn = source.lazy(None)
    x = torch.randn((2, *unit), device="cuda")
    o = torch.randn((2, *unit), device="cuda")
    sampler = structured.Euler(stochasticity=1)
    a = sampler.sample(x, o, Step.from_int(2, 9), models.FlowModel(), scheduling.Linear(), drawn)
    b = sampler.sample(x, o, Step.from_int(2, 9), models.FlowModel(), scheduling.Linear(), drawn.materialize())
    assert torch.equal(a.final, b.final)


@gpu
def test_batch_auto_hands_out_philox_keys_and_the_step_decides() -> None:
    """BatchTensorNoise.auto: Philox keys whenever every item is a plain Random on one device (identical values to the
    filled tensor); the STEP decides where they become values - inside its kernel (one launch), or, for the
    issue-bound divided-difference steps on large latents, through the fill kernel first (Program.settle_noise)."""
    from skrample_b200 import native, scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    small = noise.BatchTensorNoise.from_batch_inputs(noise.Random, (4, 64, 64), [_gen(3), _gen(4)])
    twin = noise.BatchTensorNoise.from_batch_inputs(noise.Random, (4, 64, 64), [_gen(3), _gen(4)])
    drawn = small.auto(None)
    assert getattr(drawn, "is_lazy_noise", False)
    assert torch.equal(drawn.materialize(), twin.generate(None))
    mixed = noise.BatchTensorNoise([noise.Random.from_inputs((8,), _gen(1)), noise.Offset.from_inputs((8,), _gen(2))])
    assert isinstance(mixed.auto(None), torch.Tensor)

    unit = (16, 256, 256)  # 2 x 1 Mi elements: above Program.SETTLE_ELEMENTS
    schedule, model = scheduling.FlowShift(scheduling.Linear()), models.FlowModel()
    # UniPC(2): step 0 is a plain first-order step (drawn in the kernel); step 1 is the first divided-difference step and
    # fills this step's draw and the previous one (the corrector reads it); from then on one fill per step
    for sampler, expected in ((structured.Euler(stochasticity=1), (0, 0, 0)), (structured.UniPC(order=2, stochasticity=1), (0, 2, 1))):
        big = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, [_gen(5), _gen(6)])
        ref = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, [_gen(5), _gen(6)])
        x = torch.randn((2, *unit), device="cuda")
        previous: list = []
        previous_ref: list = []
        for n in range(3):
            out = torch.randn((2, *unit), device="cuda")
            lazy = big.auto(None)
            assert getattr(lazy, "is_lazy_noise", False)
            before = native.launch_count_kind(2)
            got = sampler.sample(x, out, Step.from_int(n, 8), model, schedule, lazy, previous)
            assert native.launch_count_kind(2) - before == expected[n], f"{type(sampler).__name__} step {n}"
            want = sampler.sample(x, out, Step.from_int(n, 8), model, schedule, ref.generate(None), previous_ref)
            assert torch.equal(got.final, want.final)
            previous = (previous + [got])[-sampler.require_previous :] if sampler.require_previous else []
            previous_ref = (previous_ref + [want])[-sampler.require_previous :] if sampler.require_previous else []
            x = got.final


@pytest.mark.skipif(_has_torchsde(), reason="torchsde present: the reference's tree is used")
def test_brownian_without_torchsde_says_so(monkeypatch: pytest.MonkeyPatch) -> None:
    """The reference raises ImportError without torchsde (noise.py:223); this library evaluates its own bridge tree
    instead, warns once that the values are not torchsde's, and raises like the reference on request."""
    import warnings

    monkeypatch.setattr(noise, "_warned_no_torchsde", False)
    with pytest.warns(RuntimeWarning, match="torchsde is not installed"):
        noise.Brownian.from_inputs((4, 8), torch.Generator().manual_seed(1))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        noise.Brownian.from_inputs((4, 8), torch.Generator().manual_seed(1))  # said once
    monkeypatch.setenv("SKRAMPLE_B200_REQUIRE_TORCHSDE", "1")
    with pytest.raises(ImportError):
        noise.Brownian.from_inputs((4, 8), torch.Generator().manual_seed(1))


@gpu
@pytest.mark.parametrize(("items", "unit"), [(64, (4, 32, 32)), (256, (3, 5, 8)), (40, (16, 64, 64))], ids=["64-items", "256-items", "40-items-2.6M"])
def test_in_kernel_noise_beyond_32_items_and_a_million_elements(items: int, unit: tuple[int, ...]) -> None:
    """Round 1 limited in-kernel draws to 32 batch items and 2^20 elements.  The key tables are now a kernel parameter of
    their own (256 items), and a batched fill takes any number of items (chunks of 256): an Euler SDE step that draws
    its noise from Philox keys equals the step on the materialised tensor bit for bit, item by item."""
    from skrample_b200 import native, scheduling
    from skrample_b200.sampling import models, structured

    gens = lambda: [_gen(900 + i) for i in range(items)]  # noqa: E731
    lazy = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, gens())
    full = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, gens())
    single = noise.Random.from_inputs(unit, _gen(900 + items - 1))
    sampler = structured.Euler(stochasticity=1)
    schedule, model = scheduling.FlowShift(scheduling.Linear()), models.FlowModel()
    x = torch.randn((items, *unit), device="cuda")
    out = torch.randn((items, *unit), device="cuda")
    drawn = lazy.auto(None)
    assert getattr(drawn, "is_lazy_noise", False) and len(drawn.seeds) == items
    fills = native.launch_count_kind(2)
    got = sampler.sample(x, out, (0.2, 0.3), model, schedule, drawn)
    assert native.launch_count_kind(2) == fills, "the step drew its own noise: no fill kernel"
    tensor = full.generate(None)
    assert torch.equal(tensor[-1], single.generate(None)), "item i of a batch is its own generator's stream"
    want = sampler.sample(x, out, (0.2, 0.3), model, schedule, tensor)
    assert torch.equal(got.final, want.final)


@gpu
def test_batched_fill_of_more_items_than_one_key_table() -> None:
    "300 per-item generators: the batched fill runs in chunks of 256 items; every item is its own generator's stream."
    unit = (2, 6, 10)
    batch = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, [_gen(3000 + i) for i in range(300)])
    assert isinstance(batch.auto(None), torch.Tensor)  # more items than the step kernel's key table: a tensor
    batch = noise.BatchTensorNoise.from_batch_inputs(noise.Random, unit, [_gen(3000 + i) for i in range(300)])
    got = batch.generate(None)
    for i in (0, 1, 255, 256, 257, 299):
        assert torch.equal(got[i], noise.Random.from_inputs(unit, _gen(3000 + i)).generate(None)), i


# ------------------------------------------------------------------------------------------- Brownian (CUDA)


def _fill_native(seed: int, stream: int, numel: int) -> torch.Tensor:
    from skrample_b200 import native

    out = torch.empty(numel, device="cuda")
    native.check(noise._lib().skr_noise_fill(out.data_ptr(), 0, numel, seed, stream, None, None, noise._stream()), "skr_noise_fill")
    return out


@gpu
@pytest.mark.parametrize(("seed", "stream", "numel"), [(1234, 0, 4096), (2**63 + 12345, 2**40 + 3, 1027), (7, O.BROWNIAN_TREE | 5, 64)])
def test_fill_values_match_the_oracle_philox(seed: int, stream: int, numel: int) -> None:
    "The normals skr_noise_fill writes are the oracle's Philox4x32-10 + Box-Muller stream, element by element."
    got = _fill_native(seed, stream, numel).cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, O.philox_normals(seed, stream, numel), rtol=0, atol=1e-5)  # the kernel's Box-Muller runs on the special-function unit (philox.cuh)


@gpu
@pytest.mark.parametrize(("shape", "dims"), [((6, 10, 12), (0,)), ((3, 4, 20), (0, 1)), ((5, 7, 3), (0,))])
def test_offset_values_match_the_oracle_philox(shape: tuple[int, ...], dims: tuple[int, ...]) -> None:
    """Offset on a CUDA generator (noise.py:104-113): the oracle's normals of the draw's stream plus, per kept slice, the
    oracle's normal of the next stream at the slice index times strength^2 - through the fill kernel (one generator)
    and through the key-table path (a batch of generators)."""
    props = noise.OffsetProps(dims=dims, strength=0.6)
    one = noise.Offset.from_inputs(shape, _gen(31), props)
    probe = noise.Offset.from_inputs(shape, _gen(31), props)
    key, tick = probe._key(), probe._tick()  # the stream block the first draw of an identically seeded generator takes
    numel = math.prod(shape)
    rows = math.prod(shape[: len(dims)])
    shift = O.philox_normals(key, tick + 1, rows) * props.strength**2
    want = (O.philox_normals(key, tick, numel).reshape(rows, -1) + shift[:, None]).reshape(shape)
    got = one.generate(None)
    np.testing.assert_allclose(got.cpu().numpy().astype(np.float64), want, rtol=0, atol=2e-5)
    batch = noise.BatchTensorNoise.from_batch_inputs(noise.Offset, shape, [_gen(31), _gen(31)], props)
    drawn = batch.lazy(None)
    assert isinstance(drawn, noise.PhiloxDraw)
    both = drawn.materialize()
    assert torch.equal(both[0], got) and torch.equal(both[1], got)


@gpu
@pytest.mark.parametrize("step", BROWNIAN_STEPS)
@pytest.mark.parametrize("max_steps", [10_000, 50])
def test_brownian_kernel_vs_oracle(step: tuple[float, float], max_steps: int) -> None:
    "Relative fp32 increments in the kernel against absolute float64 path values in the oracle."
    shape = (3, 17, 23)  # 1173 elements: ragged last group
    g = noise.Brownian.from_inputs(shape, _gen(77), noise.BrownianProps(max_steps=max_steps))
    depth = math.ceil(math.log2(max_steps * 10))
    got = g.generate(Step(*step))
    assert got.shape == shape and got.dtype == torch.float32 and got.is_cuda
    want = O.brownian_increment(77, step[0], step[1], depth, got.numel()).reshape(shape)
    np.testing.assert_allclose(got.cpu().numpy().astype(np.float64), want, rtol=0, atol=2e-4)


@gpu
@pytest.mark.skipif(_has_torchsde(), reason="with torchsde installed a CPU generator uses the reference's tree")
@pytest.mark.parametrize("step", [(0.0, 0.04), (0.48, 0.52), (0.3, 0.3001)])
def test_brownian_same_path_on_host_and_device(step: tuple[float, float]) -> None:
    "A seed names one path: the host evaluation (CPU generator) and the kernel (CUDA generator) agree to float32 rounding."
    shape = (5, 40, 33)
    host = noise.Brownian.from_inputs(shape, torch.Generator().manual_seed(21)).generate(Step(*step))
    device = noise.Brownian.from_inputs(shape, _gen(21)).generate(Step(*step))
    torch.testing.assert_close(device.cpu(), host, rtol=0, atol=3e-4)


@gpu
def test_brownian_is_a_function_of_the_step() -> None:
    "reference: noise.py:220 'deterministically over Step' - call order, call count and step direction do not matter."
    shape = (4, 64, 64)
    a = noise.Brownian.from_inputs(shape, _gen(3))
    b = noise.Brownian.from_inputs(shape, _gen(3))
    s0, s1 = Step.from_int(3, 25), Step.from_int(4, 25)
    x0, x1 = a.generate(s0), a.generate(s1)
    assert torch.equal(b.generate(s1), x1) and torch.equal(b.generate(s0), x0) and torch.equal(a.generate(s0), x0)
    assert torch.equal(a.generate(Step(s0.time_to, s0.time_from)), x0), "Step.normal() orders the ends"
    assert not torch.equal(x0, x1)
    assert not torch.equal(noise.Brownian.from_inputs(shape, _gen(4)).generate(s0), x0)
    # step=None is a plain draw from the generator (noise.py:241-242)
    assert torch.equal(a.generate(None), noise.Random.from_inputs(shape, _gen(3)).generate(None))
    # clamp() keeps out-of-range steps inside 0..1 (noise.py:243)
    torch.testing.assert_close(a.generate(Step(0.98, 1.02)), a.generate(Step(0.96, 1.0)), rtol=0, atol=1e-5)
    with pytest.raises(ZeroDivisionError):
        a.generate(Step(0.5, 0.5))


@gpu
def test_brownian_increments_add_up_and_are_independent() -> None:
    shape = (16, 128, 128)
    g = noise.Brownian.from_inputs(shape, _gen(11))
    steps = [Step.from_int(i, 25) for i in range(25)]
    parts = [g.generate(s).double() for s in steps]
    n = parts[0].numel()
    total = sum(p * math.sqrt(s.distance()) for p, s in zip(parts, steps))
    whole = g.generate(Step(0.0, 1.0)).double()
    assert (total - whole).abs().max().item() < 5e-6
    pair = g.generate(Step(steps[6].time_from, steps[7].time_to)).double()
    assert ((parts[6] + parts[7]) * math.sqrt(steps[6].distance()) - pair * math.sqrt(2 * steps[6].distance())).abs().max().item() < 2e-6
    for x in (*parts[:4], parts[12], parts[24], whole, pair):
        x = x.flatten()
        assert abs(x.mean().item()) < 5 / math.sqrt(n)
        assert abs(x.var().item() - 1) < 1e-2
        assert abs((x**3).mean().item()) < 3e-2 and abs((x**4).mean().item() - 3) < 8e-2
        assert abs((x[:-1] * x[1:]).mean().item()) < 5 / math.sqrt(n)
    for i, j in [(0, 1), (6, 7), (11, 12), (12, 13), (0, 24), (3, 17)]:
        assert abs((parts[i] * parts[j]).mean().item()) < 5 / math.sqrt(n), (i, j)


@gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float64])
def test_brownian_storage_dtypes_and_batches(dtype: torch.dtype) -> None:
    from skrample_b200 import native

    unit = (4, 33, 31)  # odd item size: items 1.. start unaligned
    step = Step.from_int(7, 20)
    want = torch.stack([noise.Brownian.from_inputs(unit, _gen(50 + i)).generate(step) for i in range(3)])
    batch = noise.BatchTensorNoise.from_batch_inputs(noise.Brownian, unit, [_gen(50 + i) for i in range(3)], noise.BrownianProps(), dtype)
    got = batch.generate(step)
    assert got.dtype == dtype and got.shape == (3, *unit)
    assert torch.equal(got, want.to(dtype)), "storage dtypes are the fp32 increment rounded once"
    out = torch.full((3, *unit), 9.0, device="cuda", dtype=dtype)
    before = native.launch_count_kind(2)
    batch.generate_into(out, step)
    assert native.launch_count_kind(2) - before == 1, "one launch for the whole batch"
    assert torch.equal(out, got)
    # the batch path writes any requested storage type directly; mixed tree depths go item by item
    wide = noise.BatchTensorNoise.from_batch_inputs(noise.Brownian, unit, [_gen(50 + i) for i in range(3)], noise.BrownianProps())
    other = torch.empty((3, *unit), device="cuda", dtype=dtype)
    wide.generate_into(other, step)
    assert torch.equal(other, got)
    narrow = torch.empty((3, *unit), device="cuda", dtype=torch.float64)
    batch.generate_into(narrow, step)
    assert torch.equal(narrow, got.double()), "a 16-bit generator keeps its rounding when read into a wider tensor"
    mixed = noise.BatchTensorNoise([noise.Brownian.from_inputs(unit, _gen(50)), noise.Brownian.from_inputs(unit, _gen(51), noise.BrownianProps(max_steps=100))])
    assert torch.equal(mixed.generate(step)[0], want[0]) and not torch.equal(mixed.generate(step)[1], want[1])


@gpu
def test_brownian_abi_rejects_bad_intervals() -> None:
    lib = noise._lib()
    out = torch.empty(8, device="cuda")
    for t0, t1, depth in [(0.5, 0.5, 17), (0.6, 0.5, 17), (-0.1, 0.5, 17), (0.5, 1.1, 17), (0.1, 0.2, 0), (0.1, 0.2, 41), (float("nan"), 0.5, 17)]:
        assert lib.skr_noise_brownian(out.data_ptr(), 0, 8, 1, t0, t1, depth, 1.0, noise._stream()) < 0
    assert lib.skr_noise_brownian(out.data_ptr(), 0, 0, 1, 0.1, 0.2, 17, 1.0, noise._stream()) == 0
    assert lib.skr_noise_brownian(None, 0, 8, 1, 0.1, 0.2, 17, 1.0, noise._stream()) < 0
