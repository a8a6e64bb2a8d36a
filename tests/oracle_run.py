"""Drive ``oracle/skrample_oracle.py`` from the shared case table (test helper)."""

from __future__ import annotations

from typing import Any

import numpy as np

import cases
from oracle import skrample_oracle as O

MODELS = {
    "DataModel": O.Model("data"),
    "NoiseModel": O.Model("noise"),
    "FlowModel": O.Model("flow"),
    "VelocityModel": O.Model("velocity"),
    "ScaleX": O.Model("scalex"),
}


def schedule(name: str) -> O.Schedule:
    if name == "scaled":
        return O.scaled()
    if name == "flow":
        return O.flow_shift(O.linear())
    if name == "linear":
        return O.linear()
    if name == "hyper_linear":
        return O.hyper(O.linear())
    if name == "hyper_scaled":
        return O.hyper(O.scaled())
    if name == "sinner_linear":
        return O.sinner(O.linear())
    raise KeyError(name)


def supported(case: dict) -> bool:
    if case["schedule"] not in ("scaled", "flow", "linear", "hyper_linear", "hyper_scaled", "sinner_linear"):
        return False
    kw = case["kw"]
    if case["sampler"] == "UniPC" and "predictor" in kw:
        return False
    if case["sampler"] == "SPC" and ("predictor" in kw or "corrector" in kw):
        return False
    return True


def require(case: dict) -> tuple[bool, int]:
    "(require_noise, require_previous) of the case's sampler (reference: structured.py:51-59,134-135,460-467,518-524)."
    kw = case["kw"]
    name = case["sampler"]
    noise = abs(kw.get("stochasticity", 0)) > 1e-8
    if name == "Euler":
        return noise, 0
    if name == "SPC":
        return False, 4
    order = kw.get("order", 2)
    cap = 3 if name == "DPM" else 9
    prev = max(min(order, cap), 1) - 1
    if name == "UniPC":
        prev += 1
    return noise, prev


def one_step(case: dict, cur: O.Rec, previous: list[O.Rec], model: O.Model, sch: O.Schedule) -> O.Rec:
    kw = case["kw"]
    name = case["sampler"]
    eta = kw.get("stochasticity", 0)
    deriv: Any = O.DATA
    if "derivative_transform" in kw:
        deriv = None if kw["derivative_transform"] is None else O.Model(MODELS[kw["derivative_transform"]].kind)  # a distinct object, like the reference case
    order = kw.get("order", 2)
    if name == "Euler":
        cur.final = O.euler_step(cur, model, sch, eta)
    elif name == "DPM":
        cur.final = O.dpm_step(cur, model, sch, previous, order, eta, deriv)
    elif name == "Adams":
        cur.final = O.adams_step(cur, model, sch, previous, order, eta, deriv)
    elif name == "UniP":
        cur.final = O.uni_solve(cur, model, sch, previous, order, eta, deriv, kw.get("fast_solve", False))
    elif name == "UniPC":
        cur = O.unipc_step(cur, model, sch, previous, order, eta, deriv, kw.get("fast_solve", False))
    elif name == "SPC":
        cur = O.spc_step(cur, model, sch, previous, deriv, 4, kw.get("bias", 0), kw.get("power", 1), kw.get("adaptive", True), kw.get("invert", False))
    else:
        raise KeyError(name)
    return cur


def run_structured(case: dict, np_dtype: Any = None) -> O.Rec:
    "Whole trajectory of a case through the oracle; returns the last step's record."
    np_dtype = np_dtype or {"f32": np.float32, "f64": np.float64}[case["dtype"]]
    model = MODELS[case["model"]]
    sch = schedule(case["schedule"])
    need_noise, need_prev = require(case)
    x0, outs, noises = cases.trajectory_inputs(case)
    x = x0.astype(np_dtype)
    previous: list[O.Rec] = []
    rec = None
    for n in range(case["steps"]):
        cur = O.Rec(x, outs[n].astype(np_dtype), O.St.from_int(n, case["steps"]), noises[n].astype(np_dtype) if need_noise else None)
        rec = one_step(case, cur, previous, model, sch)
        previous.append(rec)
        previous = previous[max(len(previous) - need_prev, 0) :]
        x = rec.final
    assert rec is not None
    return rec
