"""The block kernel divides by grid-uniform scalars with a host-side reciprocal and a residual correction
(csrc/machine.cuh, div_uniform); the interpreter uses IEEE division.  Both must agree bit for bit - with each other
and with torch's CPU division - on adversarial operands: every exponent, zeros, subnormals, infinities, NaNs.
tools/verify_divr.cu is the exhaustive proof for the guarded range; this is the regression test around the guard."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from skrample_b200.sampling import program as pg

pytestmark = pytest.mark.gpu

DIVISORS = [
    0.37,
    3.0,
    -0.77,
    1e-3,
    123456.0,
    2.0**-20,
    2.0**20,
    float(np.float32(2.0) - np.float32(2.0**-23)),  # significand all ones
    float(np.float32(1.0) + np.float32(2.0**-23)),
    1e-7,  # outside the guarded divisor range: IEEE path for the whole launch
    1e7,
    2.0**-126,
    3e38,
]


def adversarial(n: int, seed: int) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)  # every sign / exponent / significand
    values = bits.view(np.float32).copy()
    special = np.array(
        [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.1754942e-38, 1.17549435e-38, 3.4028235e38, 2.0**-60, 2.0**60,
         float(np.nextafter(np.float32(2.0**-60), np.float32(0))), float(np.nextafter(np.float32(2.0**60), np.float32(np.inf))), 1.0, -1.0],
        dtype=np.float32,
    )
    values[: special.size] = special
    values[special.size : n // 2] = rng.standard_normal(n // 2 - special.size).astype(np.float32)  # ordinary latents
    return torch.from_numpy(values)


def run(program: pg.Program, interpreter: bool, monkeypatch: pytest.MonkeyPatch) -> list[np.ndarray]:
    if interpreter:
        monkeypatch.setenv("SKR_FORCE_INTERP", "1")
    else:
        monkeypatch.delenv("SKR_FORCE_INTERP", raising=False)
    from skrample_b200 import native

    native.reset_switches()
    return [out.cpu().numpy() for out in program.run()]


def same_bits(a: np.ndarray, b: np.ndarray) -> bool:
    nan = np.isnan(a) & np.isnan(b)  # NaN payloads are not part of the contract
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | nan))


@pytest.mark.parametrize("divisor", DIVISORS)
@pytest.mark.parametrize("numel", [4096 + 3, 148 * 4 * 1024 * 2])
def test_conversion_division_matches_ieee(divisor: float, numel: int, monkeypatch: pytest.MonkeyPatch) -> None:
    "x-hat = y / c2 (the Data<-scaled conversions, reference: skrample/sampling/models.py:92-212)."
    y_cpu = adversarial(numel, seed=numel % 1000)
    y = y_cpu.cuda()

    def build() -> pg.Program:
        program = pg.Program()
        program.conv(pg.ConvSpec(pg.CONV_DIV, 0.0, 0.0, divisor), y)
        program.store(pg.P)
        return program

    from skrample_b200 import native

    before = native.launch_count_kind(0)
    (block,) = run(build(), False, monkeypatch)
    assert native.launch_count_kind(0) == before + 1, "the program must take the block kernel"
    (interp,) = run(build(), True, monkeypatch)
    want = (y_cpu / divisor).numpy()
    assert same_bits(interp, want)
    assert same_bits(block, want)


@pytest.mark.parametrize("divisor", DIVISORS[:8])
def test_uni_and_accumulator_divisions_match_interpreter(divisor: float, monkeypatch: pytest.MonkeyPatch) -> None:
    "UniP/UniPC difference terms (x_k - x_0) / r_k and the RK normaliser A / d, on adversarial history tensors."
    numel = 148 * 4 * 1024 + 1024 + 5
    x = torch.randn(numel).cuda()
    y = torch.randn(numel).cuda()
    history = [adversarial(numel, seed=s).cuda() for s in (1, 2, 3)]

    def uni() -> pg.Program:
        program = pg.Program()
        program.load(pg.X, x)
        program.load(pg.P, y)
        program.mov(pg.B, pg.P)
        program.uni(history[0], divisor, 0.625, first=True)
        program.uni(history[1], divisor * 1.5, -0.25, first=False)
        program.addb()
        program.fwd(0.5, 0.25, pred=pg.A)
        program.store(pg.R)
        return program

    def acc() -> pg.Program:
        program = pg.Program()
        program.load(pg.X, x)
        program.load(pg.P, y)
        program.acc(1.0, history[2], first=True)
        program.acc(0.5, None, reg=pg.P)
        program.diva(divisor)
        program.fwd(1.0, 0.125, pred=pg.A)
        program.store(pg.R)
        program.back(0.75, divisor)
        program.store(pg.P)
        return program

    for build in (uni, acc):
        block = run(build(), False, monkeypatch)
        interp = run(build(), True, monkeypatch)
        assert len(block) == len(interp)
        for got, want in zip(block, interp):
            assert same_bits(got, want)
