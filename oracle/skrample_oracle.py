"""CPU oracle for the skrample sampler-step path.  TEST INFRASTRUCTURE ONLY.

A plain NumPy restatement of the reference algorithm (Beinsezii/skrample,
``/root/reference`` at build time) used as the checker for the CUDA kernels.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
package ``skrample_b200`` never does.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this file against
  * the reference's own known-answer tables (reference: tests/self_sampling.py:57-82
    sampler trajectories, tests/self_scheduling.py:30-45 schedules,
    tests/miscellaneous.py:9-13 Bashforth weights), and
  * golden tensors produced by importing the reference itself in the build
    container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

Exception - PARITY UNPINNED: the Brownian section at the end (the reference delegates to the third-party
``torchsde``, which is neither under /root/reference nor installed; see that section's header).

Numerics model (what "the reference computes" means for tensors):
  * a Python float meeting a float32 array is rounded to float32 first, then every
    binary op is individually rounded - NumPy's weak-scalar promotion gives exactly
    torch-CPU's behaviour, so running the formulas on ``np.float32`` arrays IS the
    reference's fp32 arithmetic;
  * ``math.sumprod`` over tensors is the plain left fold ``((0 + a*x) + b*y) + ...``
    (reference: models.py:53-67, structured.py:319-322, functional.py:84,103);
  * 16-bit storage is modelled as "compute in fp32, round once" (``round_bf16`` /
    ``astype(float16)``), which is what the reference's diffusers wrapper does with
    ``compute_scale=float32`` (reference: skrample/diffusers.py:575-599).

Every function cites the reference lines it restates.  Schedules here are the
few curves the parity tests need, not the full library.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, NamedTuple, Sequence

import numpy as np

# --------------------------------------------------------------------------------------------
# scalars (reference: skrample/common.py)


class Pt(NamedTuple):
    "reference: common.py:24-30"
    timestep: float
    sigma: float
    alpha: float


class St(NamedTuple):
    "reference: common.py:55-97"
    time_from: float
    time_to: float

    @staticmethod
    def from_int(position: int, amount: int) -> "St":
        return St(position / amount, (position + 1) / amount)

    def distance(self) -> float:
        return self.time_to - self.time_from

    def position(self) -> float:
        return self.time_from / self.distance()

    def amount(self) -> float:
        return 1 / self.distance()

    def normal(self) -> "St":
        return St(min(self), max(self))

    def clamp(self) -> "St":
        d = self.distance()
        return St(max(0, min(1 - d, self.time_from)), max(d, min(1, self.time_to)))


def divf(a: float, b: float) -> float:
    "reference: common.py:133-140"
    if b != 0:
        return a / b
    if a == 0:
        raise ZeroDivisionError
    return math.copysign(math.inf, a)


def ln(x: float) -> float:
    "reference: common.py:143-150"
    if x > 0:
        return math.log(x)
    if x < 0:
        raise ValueError
    return -math.inf


def bashforth(order: int) -> list[float]:
    "reference: common.py:205-213"
    m = [[(-j) ** k for j in range(order)] for k in range(order)]
    v = [1 / (k + 1) for k in range(order)]
    return np.linalg.solve(m, v).tolist()


def softmax2(a: float, b: float) -> tuple[float, float]:
    "reference: common.py:181-184"
    ea, eb = math.e**a, math.e**b
    total = 0 + ea + eb
    return ea / total, eb / total


def fold_sumprod(values: Sequence[Any], coeffs: Sequence[float]) -> Any:
    "Left fold of math.sumprod's generic path: ((0 + v0*c0) + v1*c1) + ..."
    acc: Any = 0
    for v, c in zip(values, coeffs, strict=True):
        acc = acc + v * c
    return acc


# --------------------------------------------------------------------------------------------
# schedules (reference: skrample/scheduling.py) - just enough curves for the parity tests


@dataclass(frozen=True)
class Schedule:
    """``fn(t)`` maps noise-time (1 = all noise) to rows (timestep, sigma, alpha)."""

    fn: Callable[[np.ndarray], np.ndarray] = field(compare=False)
    name: str = ""

    def points_np(self, t: Sequence[float]) -> np.ndarray:
        "reference: scheduling.py:79-82"
        return self.fn(np.asarray(t, dtype=np.float64).clip(0, 1))

    def ipoints_np(self, t: Sequence[float]) -> np.ndarray:
        "reference: scheduling.py:89-92"
        return self.fn(1 - np.asarray(t, dtype=np.float64).clip(0, 1))

    def ipoints(self, t: Sequence[float]) -> list[Pt]:
        return [Pt(*row) for row in self.ipoints_np(t).tolist()]

    def point(self, t: float) -> Pt:
        "reference: scheduling.py:99-102"
        return Pt(*self.fn(np.expand_dims(np.float64(t).clip(0, 1), 0))[0].tolist())

    def ipoint(self, t: float) -> Pt:
        "reference: scheduling.py:104-107"
        return Pt(*self.fn(np.expand_dims(1 - np.float64(t).clip(0, 1), 0))[0].tolist())

    def schedule_np(self, steps: int) -> np.ndarray:
        "reference: scheduling.py:129-131"
        return self.fn(np.linspace(1, 0, steps, endpoint=False))

    def schedule(self, steps: int) -> list[Pt]:
        return [Pt(*row) for row in self.schedule_np(steps).tolist()]


def linear(sigma_start: float = 1.0, base_timesteps: int = 1000) -> Schedule:
    "reference: scheduling.py:281-308 (flow-matching space for sigma_start <= 1, else variance preserving)"

    def fn(t: np.ndarray) -> np.ndarray:
        s = t * sigma_start
        if sigma_start <= 1:
            sig, alp = np.asarray(s), 1 - np.asarray(s)
        else:
            th = np.atan(s)
            sig, alp = np.sin(th), np.cos(th)
        return np.stack([t * abs(base_timesteps), sig, alp], axis=1)

    return Schedule(fn, "linear")


def scaled(beta_start: float = 0.00085, beta_end: float = 0.012, beta_scale: float = 2, base_timesteps: int = 1000) -> Schedule:
    "reference: scheduling.py:180-245"

    def fn(t: np.ndarray) -> np.ndarray:
        k = beta_scale
        T = abs(base_timesteps)
        rs = beta_start ** (1 / k)
        re = beta_end ** (1 / k)
        slope = re - rs
        ib = ((rs + slope * t) ** (k + 1) - rs ** (k + 1)) / (slope * (k + 1))
        ib2 = ((rs + slope * t) ** (2 * k + 1) - rs ** (2 * k + 1)) / (slope * (2 * k + 1))
        acp = np.exp(-(T * (ib + ib2 / 2)))
        with np.errstate(divide="ignore"):
            sig = np.sqrt((1 - acp) / acp)
        th = np.atan(sig)
        return np.stack([t * T, np.sin(th), np.cos(th)], 1)

    return Schedule(fn, "scaled")


def flow_shift(base: Schedule, shift: float = 3.0) -> Schedule:
    "reference: scheduling.py:583-592"
    return Schedule(lambda t: base.fn(shift * t / (1 + (shift - 1) * t)), f"flowshift({base.name})")


def hyper(base: Schedule, scale: float = 2, tail: bool = True) -> Schedule:
    "reference: scheduling.py:595-614"

    def fn(t: np.ndarray) -> np.ndarray:
        pts = np.concatenate([[1], t]) * (scale - (-scale * tail)) + (-scale * tail)
        pts = np.sinh(pts) if scale < 0 else np.tanh(pts / math.sqrt(2))
        lo = -pts[0] * tail
        return base.fn((pts[1:] - lo) / (pts[0] - lo))

    return Schedule(fn, f"hyper({base.name})")


def sinner(base: Schedule, count: float = -2, scale: float = 2) -> Schedule:
    "reference: scheduling.py:617-664"

    def fn(t: np.ndarray) -> np.ndarray:
        x = count * 2 ** math.copysign(1, count)
        n = (abs(x) + 1) ** math.copysign(1, x) + 1
        u = np.concatenate([[0, 1], 1 - t])
        period = u * (math.pi * n)
        if scale >= 0:
            period += math.pi
        s = abs(scale) ** -1 + 1
        pts = np.sin(period) + period * s
        return base.fn((pts[2:] - pts[1]) / (pts[0] - pts[1]))

    return Schedule(fn, f"sinner({base.name})")


# --------------------------------------------------------------------------------------------
# model algebra (reference: skrample/sampling/models.py)


def zeta_ts(frm: Pt, to: Pt, eta: float, eps: float = 1e-8) -> float:
    "reference: models.py:30-38"
    if abs(eta) < eps or abs(to.sigma) < eps:
        return 0
    ratio = (frm.alpha * to.sigma) / (to.alpha * frm.sigma)
    var = (to.sigma**2) * (1.0 - ratio**2)
    return eta * math.sqrt(max(0.0, var))


def eta_transform(frm: Pt, to: Pt, eta: float) -> tuple[Pt, Pt]:
    "reference: models.py:44-51"
    z = zeta_ts(frm, to, eta)
    if z != 0:
        to = Pt(to.timestep, math.sqrt(max(0.0, to.sigma**2 - z**2)), to.alpha)
    return frm, to


@dataclass(frozen=True)
class Model:
    "kind in {data, noise, flow, velocity, scalex}; reference: models.py:86-212"

    kind: str
    bias: float = 3

    def xs(self, p: Pt) -> float:
        "reference: models.py:192-196"
        return math.exp(-math.log10(abs(self.bias) + 1) * (p.sigma if self.bias < 0 else p.alpha))

    def to_x(self, sample: Any, out: Any, p: Pt) -> Any:
        _, s, a = p
        if self.kind == "data":
            return out
        if self.kind == "noise":
            return (sample - s * out) / a
        if self.kind == "flow":
            return (sample - s * out) / (a + s)
        if self.kind == "velocity":
            return a * sample - s * out
        return out * self.xs(p)

    def from_x(self, sample: Any, x: Any, p: Pt) -> Any:
        _, s, a = p
        if self.kind == "data":
            return x
        if self.kind == "noise":
            return (sample - a * x) / s
        if self.kind == "flow":
            return (sample - (a + s) * x) / s
        if self.kind == "velocity":
            return (a * sample - x) / s
        return x / self.xs(p)

    def gamma(self, frm: Pt, to: Pt, eta: float = 0) -> float:
        if self.kind == "noise":
            return to.alpha / frm.alpha
        f, t = eta_transform(frm, to, eta)
        if self.kind in ("data", "scalex"):
            return t.sigma / f.sigma
        if self.kind == "flow":
            return (t.sigma + t.alpha) / (f.sigma + f.alpha)
        return (t.sigma / f.sigma) * (1 - f.alpha * f.alpha) + t.alpha * f.alpha

    def delta(self, frm: Pt, to: Pt, eta: float = 0) -> float:
        f, t = eta_transform(frm, to, eta)
        if self.kind == "data":
            return t.alpha - f.alpha * t.sigma / f.sigma
        if self.kind == "noise":
            return t.sigma - (t.alpha * f.sigma) / f.alpha
        if self.kind == "flow":
            return (f.alpha * t.sigma - t.alpha * f.sigma) / (f.alpha + f.sigma)
        if self.kind == "velocity":
            return f.alpha * t.sigma - t.alpha * f.sigma
        return (t.alpha - f.alpha * t.sigma / f.sigma) * self.xs(f)

    def forward(self, sample: Any, out: Any, frm: Pt, to: Pt, noise: Any = None, eta: float = 0) -> Any:
        "reference: models.py:53-67"
        g, d = self.gamma(frm, to, eta), self.delta(frm, to, eta)
        if noise is not None and (z := zeta_ts(frm, to, eta)) != 0:
            return fold_sumprod((sample, out, noise), (g, d, z))
        return fold_sumprod((sample, out), (g, d))

    def backward(self, sample: Any, result: Any, frm: Pt, to: Pt, noise: Any = None, eta: float = 0) -> Any:
        "reference: models.py:69-83"
        g, d = self.gamma(frm, to, eta), self.delta(frm, to, eta)
        if noise is not None and (z := zeta_ts(frm, to, eta)) != 0:
            return (result - sample * g - noise * z) / d
        return (result - sample * g) / d


DATA = Model("data")


def convert(frm: Model, to: Model, sample: Any, out: Any, p: Pt) -> Any:
    "reference: models.py:220-224 (identity only for the same object)"
    if frm is to:
        return out
    return to.from_x(sample, frm.to_x(sample, out, p), p)


# --------------------------------------------------------------------------------------------
# structured samplers (reference: skrample/sampling/structured.py)


@dataclass
class Rec:
    "One finished step (the reference's SKSamples). reference: structured.py:16-40"

    sample: Any
    prediction: Any
    step: St
    noise: Any
    final: Any = None


def effective_order(order: int, max_order: int, step: St, n_previous: int) -> int:
    "reference: structured.py:137-149"
    pos = step.position()
    return max(1, min(max_order, round(pos + 1), order, n_previous + 1, round(step.amount() - pos)))


def _lam(p: Pt) -> float:
    return ln(divf(p.alpha, p.sigma))


def _predictions(cur: Rec, model: Model, deriv: Model | None, sch: Schedule, previous: Sequence[Rec], k: int) -> tuple[list[Any], Model]:
    "Shared head of DPM/Adams/UniP. reference: structured.py:207-220, 304-317, 356-371"
    frm = sch.ipoints(cur.step)[0]
    if deriv is not None:
        preds = [convert(model, deriv, cur.sample, cur.prediction, frm)]
        hist = previous[len(previous) - (k - 1) :] if k > 1 else []
        preds += [convert(model, deriv, p.sample, p.prediction, sch.ipoints(p.step)[0]) for p in reversed(hist)]
        return preds, deriv
    hist = previous[len(previous) - (k - 1) :] if k > 1 else []
    return [cur.prediction, *[p.prediction for p in reversed(hist)]], model


def euler_step(cur: Rec, model: Model, sch: Schedule, eta: float = 0) -> Any:
    "reference: structured.py:167-180"
    frm, to = sch.ipoints(cur.step)
    return model.forward(cur.sample, cur.prediction, frm, to, cur.noise, eta)


def dpm_step(cur: Rec, model: Model, sch: Schedule, previous: Sequence[Rec], order: int = 2, eta: float = 0, deriv: Model | None = DATA) -> Any:
    "reference: structured.py:195-283"
    frm, to = sch.ipoints(cur.step)
    k = effective_order(order, 3, cur.step, len(previous))
    preds, fmodel = _predictions(cur, model, deriv, sch, previous, k)
    pred = preds[0]
    if k >= 2:
        lam, lam_next = _lam(frm), _lam(to)
        h = abs(lam_next - lam)
        lam_prev = _lam(sch.ipoint(previous[-1].step.time_from))
        r = (lam - lam_prev) / h
        d10 = (1.0 / r) * (pred - preds[1])
        if k >= 3:
            lam_prev2 = _lam(sch.ipoint(previous[-2].step.time_from))
            r2 = (lam_prev - lam_prev2) / h
            d11 = (1.0 / r2) * (preds[1] - preds[2])
            d1 = d10 + (r / (r + r2)) * (d10 - d11)
            d2 = (1.0 / (r + r2)) * (d10 - d11)
            hh = -h
            e = math.expm1(hh)
            c1 = (e / hh - 1.0) / e if e != 0 else 0
            c2 = ((e - hh) / hh**2 - 0.5) / e if e != 0 else 0
            pred = pred + c1 * d1 + c2 * d2
        else:
            pred = pred + 0.5 * d10
    return fmodel.forward(cur.sample, pred, frm, to, cur.noise, eta)


def adams_step(cur: Rec, model: Model, sch: Schedule, previous: Sequence[Rec], order: int = 2, eta: float = 0, deriv: Model | None = DATA) -> Any:
    "reference: structured.py:294-330"
    frm, to = sch.ipoints(cur.step)
    k = effective_order(order, 9, cur.step, len(previous))
    preds, fmodel = _predictions(cur, model, deriv, sch, previous, k)
    weighted = fold_sumprod(preds[:k], bashforth(k))
    return fmodel.forward(cur.sample, weighted, frm, to, cur.noise, eta)


def uni_solve(
    cur: Rec,
    model: Model,
    sch: Schedule,
    previous: Sequence[Rec],
    order: int = 2,
    eta: float = 0,
    deriv: Model | None = DATA,
    fast_solve: bool = False,
    prediction_next: Any = None,
) -> Any:
    "UniP, or UniC when prediction_next is given. reference: structured.py:344-436"
    frm, to = sch.ipoints(cur.step)
    k = effective_order(order, 9, cur.step, len(previous))
    preds, fmodel = _predictions(cur, model, deriv, sch, previous, k)
    if deriv is not None and prediction_next is not None:
        prediction_next = convert(model, deriv, cur.sample, prediction_next, frm)
    pred = preds[0]
    lam, lam_next = _lam(frm), _lam(to)
    h = abs(lam_next - lam)
    hh = -h
    big_b = math.expm1(hh)
    rks: list[float] = []
    d1s: list[Any] = []
    for n in range(1, k):
        lam_n = _lam(sch.ipoints(previous[-n].step)[0])
        rk = (lam_n - lam) / h
        rks.append(rk if math.isfinite(rk) else 0)
        d1s.append((preds[n] - pred) / rk)
    if prediction_next is not None:
        rks.append(1.0)
        check = 1
        d1s.append(prediction_next - pred)
    else:
        check = 2
    if not rks or (k == check and fast_solve):
        rhos = [0.5]
    else:
        phi = big_b / hh - 1
        rows, rhs = [], []
        for n in range(1, len(rks) + 1):
            rows.append([math.pow(v, n - 1) for v in rks])
            rhs.append(phi * math.factorial(n) / big_b)
            phi = phi / hh - 1 / math.factorial(n + 1)
        rhos = np.linalg.solve(rows, rhs).tolist()
    acc: Any = 0
    for rho, d1 in zip(rhos[: len(d1s)], d1s, strict=True):
        acc = acc + rho * d1
    pred = pred + acc
    return fmodel.forward(cur.sample, pred, frm, to, cur.noise, eta)


def unipc_step(
    cur: Rec,
    model: Model,
    sch: Schedule,
    previous: Sequence[Rec],
    order: int = 2,
    eta: float = 0,
    deriv: Model | None = DATA,
    fast_solve: bool = False,
) -> Rec:
    "reference: structured.py:469-497 (default predictor = UniP with the same settings)"
    frm = sch.ipoints(cur.step)[0]
    inner = model
    if deriv is not None:
        cur = Rec(cur.sample, convert(model, deriv, cur.sample, cur.prediction, frm), cur.step, cur.noise)
        inner = deriv
    if previous:
        corrected = uni_solve(previous[-1], inner, sch, previous[:-1], order, eta, deriv, fast_solve, prediction_next=cur.prediction)
        cur = Rec(corrected, cur.prediction, cur.step, cur.noise)
    cur.final = uni_solve(cur, inner, sch, previous, order, eta, deriv, fast_solve)
    return cur


def spowf(x: Any, f: float) -> Any:
    """reference: common.py:187-190.  torch multiplies the float tensor by an int64 sign tensor and stays in the float
    type; NumPy would promote float32 * int64 to float64, so the sign is built in the value's own dtype.  Exponents torch
    special-cases (2, 3, 0.5, ...: exact products / square roots) are the same special cases in NumPy; a general
    exponent goes through each library's own powf (<= 1 ulp apart, tests state the bound)."""
    sign = np.where(x < 0, -1, 1)
    if isinstance(x, np.ndarray):
        sign = sign.astype(x.dtype)
        if f == 3:
            m = np.abs(x)
            return m * m * m * sign  # torch: pow(x, 3) = x*x*x
    return np.abs(x) ** f * sign


def spc_step(
    cur: Rec,
    model: Model,
    sch: Schedule,
    previous: Sequence[Rec],
    deriv: Model | None = DATA,
    corrector_order: int = 4,
    bias: float = 0,
    power: float = 1,
    adaptive: bool = True,
    invert: bool = False,
) -> Rec:
    "Default SPC: Euler predictor, Adams(4) corrector. reference: structured.py:527-577"
    frm = sch.ipoints(cur.step)[0]
    inner = model
    if deriv is not None:
        cur = Rec(cur.sample, convert(model, deriv, cur.sample, cur.prediction, frm), cur.step, cur.noise)
        inner = deriv
    if previous:
        nxt = [*(p.prediction for p in previous[1:]), cur.prediction]
        shifted = [Rec(p.sample, q, p.step, p.noise, p.final) for p, q in zip(previous, nxt, strict=True)]
        last = shifted.pop()
        corrected = adams_step(last, inner, sch, shifted, corrector_order, 0, DATA)
        p, c = (frm.sigma, frm.alpha) if adaptive else (0, 0)
        p, c = softmax2(p - bias, c + bias)
        if invert:
            p, c = c, p
        if abs(power - 1) > 1e-8:
            mixed = spowf(spowf(cur.sample, power) * p + spowf(corrected, power) * c, 1 / power)
        else:
            mixed = cur.sample * p + corrected * c
        cur = Rec(mixed, cur.prediction, cur.step, cur.noise)
    cur.final = euler_step(cur, inner, sch, 0)
    return cur


# --------------------------------------------------------------------------------------------
# Runge-Kutta (reference: skrample/sampling/functional.py, tableaux/)


class Tableau(NamedTuple):
    "stages = ((c, (a...)), ...); weights rows. reference: tableaux/common.py:7-25"
    stages: tuple
    weights: tuple


def rk2_tableau(c1: float) -> Tableau:
    "reference: tableaux/providers.py:14-22"
    return Tableau(((0.0, ()), (c1, (c1,))), ((1 - 1 / (2 * c1), 1 / (2 * c1)),))


HEUN = rk2_tableau(1.0)


def ees25_tableau(x: float) -> Tableau:
    "reference: tableaux/providers.py:86-97"
    return Tableau(
        (
            (0.0, ()),
            ((1 + 2 * x) / (4 * (1 - x)), ((1 + 2 * x) / (4 * (1 - x)),)),
            (3 / (4 * (1 - x)), ((4 * x - 1) ** 2 / (4 * (x - 1) * (1 - 4 * x**2)), (1 - x) / (1 - 4 * x**2))),
        ),
        ((x, 1 / 2, 1 / 2 - x),),
    )


def ees27_tableau(x: float) -> Tableau:
    "reference: tableaux/providers.py:100-130"
    v2 = math.sqrt(2)
    a_ = (2 * x + v2) / ((2 * x - 1) * (-2 * x - v2 + 1))
    b_ = 1 / ((2 * x - 1) * (1 - v2 - 2 * x) * (2 - v2 - 2 * x))
    a2 = ((-2 + v2 * (1 - 2 * x)) / (4 * (x - 1)),)
    a3 = ((((2 * x + v2 - 2) * (4 * x + v2 - 2)) / (4 * v2 * (x - 1))) * a_, (0.5 * (-1 + v2)) * a_)
    a4 = (
        ((2 * x - v2) * (-40 * x**4 + (80 - 40 * v2) * x**3 - (88 - 60 * v2) * x**2 + (48 - 34 * v2) * x + 7 * v2 - 10))
        / (4 * (x - 1) * (2 * x**2 - 1))
        * b_,
        (2 - v2) * x * (x - 1) * (4 * x + v2 - 2) * b_,
        ((2 - v2) * (2 * x - v2) * (2 + v2 - 2 * x) * (x - 1) * (2 * x - 1)) / (4 * (2 * x**2 - 1) * (2 * x**2 - 4 * x + 1)),
    )
    return Tableau(
        ((0.0, ()), (math.fsum(a2), a2), (math.fsum(a3), a3), (math.fsum(a4), a4)),
        ((x, 1 / 2 * (2 - v2) - (1 - v2) * x, (1 - v2) * (x - 1), 1 / 2 * (2 - v2) - x),),
    )


def step_tableau(
    tab: Tableau,
    sample: Any,
    net: Callable[[Any, float, float, float], Any],
    model: Model,
    sch: Schedule,
    step: St,
    deriv: Model | None = None,
    noise: Any = None,
    eta: float = 0,
    epsilon: float = 1e-8,
    store: Callable[[Any], Any] | None = None,
) -> list[Any]:
    """reference: functional.py:55-105.  ``store`` models latents kept in a 16-bit storage type with fp32 arithmetic (the
    reference wrapper's compute_scale=float32 numerics, diffusers.py:575-599): every tensor that is handed to the
    network or returned - the stage inputs and the results - is rounded once by ``store``; derivatives stay fp32."""
    keep = store if store is not None else (lambda v: v)
    nodes, weights = tab.stages, tab.weights
    if deriv is not None:
        raw, src = net, model

        def net(x: Any, t: float, s: float, a: float) -> Any:  # noqa: F811 - wrap_model_call, models.py:232-239
            return convert(src, deriv, x, raw(x, t, s, a), Pt(t, s, a))

        model = deriv
    ks: list[Any] = []
    s0, s1, *fracs = sch.ipoints([*step, *(step[0] + c * (step[1] - step[0]) for c, _ in nodes)])
    for frac, (_, coeffs) in zip(fracs, nodes, strict=True):
        if coeffs:
            x = model.forward(sample, fold_sumprod(ks, coeffs) / math.fsum(coeffs), s0, frac)
        else:
            x = sample
        if abs(frac.timestep) < epsilon or abs(frac.sigma) < epsilon:
            ks.append(model.backward(sample, x, s0, s1))
        else:
            ks.append(net(keep(x) if coeffs else x, *frac))
    return [keep(model.forward(sample, fold_sumprod(ks, w), s0, s1, noise, eta)) for w in weights]


def dynasaurk_tableau(step: St, order: int = 2, per_step_decay: float = math.log(0.5) / -2, total_step_decay: float = math.log(0.5) / -20) -> Tableau:
    "reference: functional.py:302-328"
    if order >= 4:
        high, low, tf = 1 / 4 * (2 - math.sqrt(2)), 1 / 14 * (5 - 3 * math.sqrt(2)), ees27_tableau
    elif order >= 3:
        high, low, tf = 0.25, 0.1, ees25_tableau
    else:
        high, low, tf = 1, 0.5, rk2_tableau
    stages = len(tf((high + low) / 2).stages)
    st = step.normal().clamp()
    g = math.exp((-total_step_decay * st.amount() - per_step_decay * st.position()) * stages)
    g = abs(0 - min(max(g, 0), 1))
    return tf(g * high + (1 - g) * low)


# --------------------------------------------------------------------------------------------
# noise (reference: skrample/pytorch/noise.py) - transforms of *supplied* normal draws


def point_add_noise(p: Pt, sample: Any, noise: Any) -> Any:
    "reference: common.py:32-33"
    return sample * p.alpha + noise * p.sigma


def point_remove_noise(p: Pt, sample: Any, noise: Any) -> Any:
    "reference: common.py:35-40"
    return (sample - noise * p.sigma) / p.alpha


def round_bf16(x: np.ndarray) -> np.ndarray:
    "float32 -> nearest-even bfloat16, returned as float32 values."
    bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = (bits + 0x7FFF + ((bits >> 16) & 1)) & 0xFFFF0000
    nan = np.isnan(x)
    out = rounded.astype(np.uint32).view(np.float32)
    return np.where(nan, x, out)


def bilinear_axis_index(out_size: int, in_size: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    "F.interpolate(align_corners=False) source indices/weights along one axis (float32 index math like ATen)."
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = np.maximum((dst + np.float32(0.5)) * scale - np.float32(0.5), np.float32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    w1 = (src - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, w1


def colored_exponent(step: St | None, color_start: float = 0.25, color_end: float = -2, color_curve: float = 2) -> float:
    "reference: noise.py:410-420"
    if step is None:
        return color_start
    if color_curve == math.inf:
        return color_end
    st = step.normal().clamp()
    t = st.time_to
    x = -color_curve
    shift = (abs(x) + 1) ** math.copysign(1, x)
    t = shift / (shift + (divf(1, t) - 1))
    return (1 - t) * color_start + t * color_end


def colorize(white: np.ndarray, exponent: float, energy: float | None = None) -> np.ndarray:
    "reference: noise.py:338-405 (NumPy pocketfft instead of torch.fft; compare with a tolerance)"
    wstd = white.std(ddof=1)
    if exponent == 0.0:
        return white if energy is None or wstd < 1e-8 else white * (energy / wstd)
    w = np.squeeze(white).astype(np.float32 if white.dtype != np.float64 else np.float64)
    spec = np.fft.rfftn(w)
    axes = []
    nd = w.ndim
    for i, dim in enumerate(w.shape):
        if i == nd - 1:
            axes.append(np.arange(dim // 2 + 1) / dim)
        else:
            axes.append(np.abs(np.fft.fftfreq(dim, d=1.0)))
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1)
    radius = np.sqrt((grid**2).sum(-1))
    rmax = radius.max()
    if rmax > 0:
        radius = radius / rmax
    n_eff = sum(w.shape) / len(w.shape) if w.shape else 1.0
    eps_clip = 0.5 / max(n_eff, 4.0)
    weights = np.maximum(radius, eps_clip) ** (-exponent / 2.0)
    colored = np.fft.irfftn(spec * weights, s=w.shape, axes=tuple(range(nd)))
    cstd = colored.std(ddof=1)
    if cstd > 1e-8:
        colored = colored * (wstd / cstd if energy is None else energy / cstd)
    return colored.reshape(white.shape).astype(white.dtype)


def pyramid_level_shapes(shape: Sequence[int], mask: Sequence[bool], ratios: Sequence[float]) -> list[tuple[int, ...]]:
    "Level shapes for given per-level ratios r (reference draws r = rand()*2+2). reference: noise.py:158-162,197-198"
    running = list(shape)
    out: list[tuple[int, ...]] = []
    for level, r in enumerate(ratios):
        running = [max(1, int(s / (r**level))) if m else s for m, s in zip(mask, running)]
        out.append(tuple(running))
        if any(s <= 1 for m, s in zip(mask, running) if m):
            break
    return out


def upsample_linear(level: np.ndarray, shape: Sequence[int], mask: Sequence[bool]) -> np.ndarray:
    "F.interpolate(mode linear/bilinear, align_corners=False) of `level` to `shape` along the masked axes."
    out = level.astype(np.float32)
    for axis in reversed([d for d, m in enumerate(mask) if m]):  # innermost resized axis first
        i0, i1, w1 = bilinear_axis_index(shape[axis], out.shape[axis])
        lo = np.take(out, i0, axis=axis)
        hi = np.take(out, i1, axis=axis)
        wshape = [1] * out.ndim
        wshape[axis] = -1
        w = w1.reshape(wshape)
        out = (np.float32(1) - w) * lo + w * hi
    return out


def pyramid_compose(base: np.ndarray, levels: Sequence[np.ndarray], mask: Sequence[bool], strength: float = 0.3, depth: int = 99) -> np.ndarray:
    "(base + sum_l strength^l * upsample(level_l)) / std. reference: noise.py:146-207"
    top = len(levels) - 1
    skip = min(top, max(0, top - depth))
    total = np.zeros(base.shape, dtype=np.float32)
    for l, level in enumerate(levels):
        if l < skip:
            continue
        total = total + upsample_linear(level, base.shape, mask) * np.float32(strength**l)
    noise = base.astype(np.float32) + total
    return noise / noise.std(ddof=1)


# --------------------------------------------------------------------------------------------
# Brownian interval noise (reference: skrample/pytorch/noise.py:210-252)
#
# PARITY UNPINNED for the values: the reference delegates to the third-party module torchsde
# (``torchsde>=0.2.6``, pyproject.toml:22; not under /root/reference, not installed), whose interval tree
# derives its draws from numpy SeedSequence spawning.  What the reference's call site fixes is the contract:
# ``tree(t0, t1) / sqrt(t1 - t0)`` is a standard normal tensor that is a deterministic function of
# (entropy, t0, t1), increments over adjoining intervals add up, increments over disjoint intervals are
# independent.  The product keeps that contract with its own counter-based construction (a Levy bridge tree
# over dyadic intervals of 0..1 with Philox4x32-10 node draws); this restates that published construction in
# float64 with ABSOLUTE path values W(t) - an independent route to the same numbers the kernel reaches through
# relative increments in float32 - so the tree indexing, bridge variances and stream layout are checked
# element by element, and the contract itself through statistics.


def philox4x32_10(counter: np.ndarray, key: tuple[int, int]) -> np.ndarray:
    "Philox4x32-10 (Salmon et al. 2011) on rows of four uint32 counters; returns rows of four uint32."
    c = [counter[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m0, m1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & mask]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack(c, axis=1).astype(np.uint32)


def philox_normals(seed: int, stream: int, numel: int) -> np.ndarray:
    """The standard normals of Philox stream (seed, stream): counter = (element // 4, stream), two Box-Muller
    pairs per block, uniforms (x + 0.5) / 2^32 formed in float32 as the kernels form them."""
    groups = (numel + 3) // 4
    index = np.arange(groups, dtype=np.uint64)
    counter = np.stack(
        [index & np.uint64(0xFFFFFFFF), index >> np.uint64(32), np.full(groups, stream & 0xFFFFFFFF, np.uint64), np.full(groups, stream >> 32, np.uint64)],
        axis=1,
    )
    bits = philox4x32_10(counter, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = (bits.astype(np.float32) * np.float32(2.0**-32) + np.float32(2.0**-33)).astype(np.float64)
    radius0, radius1 = np.sqrt(-2.0 * np.log(u[:, 0])), np.sqrt(-2.0 * np.log(u[:, 2]))
    angle0, angle1 = 2.0 * np.pi * u[:, 1], 2.0 * np.pi * u[:, 3]
    z = np.stack([radius0 * np.sin(angle0), radius0 * np.cos(angle0), radius1 * np.sin(angle1), radius1 * np.cos(angle1)], axis=1)
    return z.reshape(-1)[:numel]


BROWNIAN_TREE, BROWNIAN_LEAF = 1 << 63, 1 << 62


def brownian_path(seed: int, t: float, depth: int, numel: int) -> np.ndarray:
    "W(t) of the seed's Brownian path (W(0) = 0, W(1) = the root draw), by midpoint bridges down to the leaf."
    left, right = 0.0, 1.0
    w_left, w_right = np.zeros(numel), philox_normals(seed, BROWNIAN_TREE, numel)
    node = 1
    for _ in range(depth):
        mid = 0.5 * (left + right)
        w_mid = 0.5 * (w_left + w_right) + 0.5 * math.sqrt(right - left) * philox_normals(seed, BROWNIAN_TREE | node, numel)
        if t >= mid:
            left, w_left, node = mid, w_mid, 2 * node + 1
        else:
            right, w_right, node = mid, w_mid, 2 * node
    width = right - left
    deviation = math.sqrt((t - left) * (right - t) / width)
    return w_left + (t - left) / width * (w_right - w_left) + deviation * philox_normals(seed, BROWNIAN_TREE | BROWNIAN_LEAF | node, numel)


def brownian_increment(seed: int, t0: float, t1: float, depth: int, numel: int) -> np.ndarray:
    "``tree(t0, t1) / sqrt(t1 - t0)`` of noise.py:244-245 for the construction above."
    return (brownian_path(seed, t1, depth, numel) - brownian_path(seed, t0, depth, numel)) / math.sqrt(t1 - t0)
