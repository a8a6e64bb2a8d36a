"""Pin the oracle: golden tensors from the reference + the reference's own known-answer tables."""

from __future__ import annotations

import json
import math
import random
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle_run
from oracle import skrample_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
STRUCTURED = np.load(GOLDEN / "structured.npz")
STRUCTURED_INDEX = json.loads((GOLDEN / "structured.json").read_text())


@pytest.mark.parametrize("case", [c for c in STRUCTURED_INDEX if oracle_run.supported(c)], ids=lambda c: c["id"])
def test_oracle_matches_reference_structured(case: dict) -> None:
    "Bit-exact: the oracle's fp32/fp64 arithmetic is the reference's torch-CPU arithmetic (bound stated for SPC power != 1)."
    rec = oracle_run.run_structured(case)
    for field in ("final", "sample", "prediction"):
        cases.assert_matches(np.asarray(getattr(rec, field)), STRUCTURED[f"{case['id']}/{field}"], case, field)


def test_fixture_table_is_current() -> None:
    assert [c["id"] for c in STRUCTURED_INDEX] == [c["id"] for c in cases.STRUCTURED_CASES]


def test_bashforth_textbook() -> None:
    "reference: tests/miscellaneous.py:9-13"
    np.testing.assert_allclose(O.bashforth(1), [1])
    np.testing.assert_allclose(O.bashforth(2), [3 / 2, -1 / 2])
    np.testing.assert_allclose(O.bashforth(3), [23 / 12, -4 / 3, 5 / 12])
    np.testing.assert_allclose(O.bashforth(4), [55 / 24, -59 / 24, 37 / 24, -3 / 8])


# reference: tests/self_scheduling.py:30-45 (MEASURED_SCHEDULE_RESULTS, 7 steps + trailing zero)
KNOWN_SCHEDULES = {
    "hyper_linear": [[1000.0, 1.0, 0.0], [914.5217069182931, 0.9145217069182932, 0.08547829308170685], [747.2650736871641, 0.7472650736871641, 0.2527349263128359], [500.0, 0.5, 0.5], [252.73492631283594, 0.25273492631283595, 0.747265073687164], [85.47829308170682, 0.08547829308170682, 0.9145217069182932], [0.0, 0.0, 1.0]],
    "sinner_linear": [[1000.0, 1.0, 0.0], [941.6266231083632, 0.9416266231083632, 0.0583733768916368], [831.9166561097303, 0.8319166561097303, 0.1680833438902697], [649.6057871357956, 0.6496057871357956, 0.35039421286420436], [415.9583280548651, 0.4159583280548651, 0.5840416719451349], [182.31086897393482, 0.1823108689739348, 0.8176891310260652], [-0.0, -0.0, 1.0]],
    "hyper_scaled": [[1000.0, 0.9976644071919207, 0.06830615362025201], [914.5217069182931, 0.9938382934567664, 0.11083973321396125], [747.2650736871641, 0.9706419180177607, 0.24052913957939248], [500.0, 0.8500395169558599, 0.5267189189818877], [252.73492631283594, 0.5739728673748606, 0.8188743172901936], [85.47829308170682, 0.2951996917929435, 0.9554355770879328], [0.0, 0.0, 1.0]],
}


@pytest.mark.parametrize("name", KNOWN_SCHEDULES)
def test_oracle_known_schedules(name: str) -> None:
    "reference: tests/self_scheduling.py:48-58 (points at linspace(1, 0, 7), rtol 1e-5)"
    got = oracle_run.schedule(name).points_np(np.linspace(1, 0, 7))
    np.testing.assert_allclose(got, np.asarray(KNOWN_SCHEDULES[name]), rtol=1e-5, atol=1e-12)


# reference: tests/self_sampling.py:57-82 (MEASURED_SAMPLER_RESULTS; generator scripts/sampling_self_measure.py:34-51)
KNOWN_TRAJECTORIES = {
    ("RKUltra", "linear", "data"): [0.5780598392186804, 0.47984288272550235, 0.36652751218728047, 0.44357997939897875, 0.6844909932628795, 0.45987338592721855, 0.5951793487736575],
    ("RKUltra", "linear", "flow"): [0.6541411854539834, 0.6575909905968154, 0.6158597768420465, 0.4742539272590272, 0.2790714207371206, 0.29258072035226296, 0.2630653895681999],
    ("RKUltra", "linear", "velocity"): [0.6124513781869325, 0.5500072845558036, 0.43877573589101987, 0.25002813549326974, 0.06398119136049375, 0.09509440076489664, 0.07264199451746078],
    ("RKUltra", "scaled", "data"): [0.6300931871907277, 0.6176224267634932, 0.6210501754229566, 0.8150691665877265, 1.103791875801471, 0.995328115543478, 1.130634078389917],
    ("RKUltra", "scaled", "flow"): [0.664132347336799, 0.708288578272592, 0.7392745926547724, 0.5874191175220908, 0.3507488804008957, 0.30995208778899164, 0.15800230361404097],
    ("RKUltra", "scaled", "velocity"): [0.6466992262396786, 0.647570542272716, 0.605409393836501, 0.3622010992613355, 0.11677787277282146, 0.16251186182339153, 0.0791635631891336],
    ("DynasauRK", "linear", "data"): [0.6437535215196428, 0.5039599707457046, 0.5031843850880807, 0.15216380075405084, 0.6699287344780329, 0.5428476583359604, 0.7539033472799924],
    ("DynasauRK", "linear", "flow"): [0.5920477557874171, 0.6376176537943613, 0.5211684005926369, 0.5835226791504955, 0.33260467455012566, 0.3143103644489387, 0.2838967980136892],
    ("DynasauRK", "linear", "velocity"): [0.5502356298556916, 0.5363255354665724, 0.3523977819326804, 0.3831554921411819, 0.12488687297136288, 0.11514939841555247, 0.09085234713376186],
    ("DynasauRK", "scaled", "data"): [0.6617890337032405, 0.6243956478591233, 0.7276437420068448, 0.5315971149946406, 1.0340724369280494, 1.00740676609858, 1.20067238415457],
    ("DynasauRK", "scaled", "flow"): [0.6351824097648943, 0.7034332655495305, 0.6587342249495243, 0.749741903872867, 0.40394335261953507, 0.31409084908338075, 0.16893219195756154],
    ("DynasauRK", "scaled", "velocity"): [0.615127136722084, 0.645619241507403, 0.50729139993555, 0.6186496466769835, 0.2177338664597791, 0.19038923706018654, 0.10760075365235315],
    ("Adams", "linear", "data"): [0.5823892132380544, 0.45238300627281497, 0.3893269179260654, 0.22944591590064134, 1.0260936490800747, 0.47614703345685516, 0.6114529963032942],
    ("Adams", "linear", "flow"): [0.652357160411046, 0.6865655116121595, 0.5897498257234484, 0.5503720966982281, 0.12391546260993933, 0.25062205385363334, 0.2240010031899688],
    ("Adams", "linear", "velocity"): [0.6082499371443788, 0.5759467522266517, 0.3996314973648122, 0.34379120213634495, -0.09453883964299331, 0.10484655646602958, 0.08104875615666654],
    ("Adams", "scaled", "data"): [0.6313689484502868, 0.5963880658262994, 0.639788604742546, 0.6078314580200911, 1.390473704962506, 1.0096464267112562, 1.1449523895576952],
    ("Adams", "scaled", "flow"): [0.6634599890325708, 0.7290803151980236, 0.7161887251367887, 0.6899822151363814, 0.12091613842652106, 0.3321052690917803, 0.171353101758476],
    ("Adams", "scaled", "velocity"): [0.6453068566447652, 0.6685746457837283, 0.5608953018885953, 0.5153161547316122, -0.1860890876184754, 0.22329686513620867, 0.12198019090845079],
    ("SPC", "linear", "data"): [0.5823892132380544, 0.4586771307892887, 0.3741928361603825, 0.20446379350926672, 1.0037336805147574, 0.4694307383584173, 0.6047367012048563],
    ("SPC", "linear", "flow"): [0.652357160411046, 0.6839718628050788, 0.6135188719216544, 0.5811791631975856, 0.14133992123072664, 0.2635454619542076, 0.23603296343782088],
    ("SPC", "linear", "velocity"): [0.6082499371443788, 0.5698386190526368, 0.41648461030717204, 0.35170696362721426, -0.11333390686698243, 0.07326063755909841, 0.05382038742020298],
    ("SPC", "scaled", "data"): [0.6313689484502868, 0.5990537809889829, 0.6145638214778157, 0.5604015372368046, 1.3984895479002883, 0.9429774144798273, 1.0782833773262663],
    ("SPC", "scaled", "flow"): [0.6634599890325708, 0.7276754161913265, 0.7441508235578342, 0.7271492552656235, 0.10823100791576633, 0.3722253198512067, 0.19553178254985126],
    ("SPC", "scaled", "velocity"): [0.6453068566447652, 0.6656652764064962, 0.5865910687657327, 0.5294679004438534, -0.24635072058839225, 0.23832813302233824, 0.13256813474580895],
}


def _known_trajectory(sampler: str, base: str, kind: str) -> list[float]:
    """The reference's ``capture`` recipe (tests/self_sampling.py:34-51) restated on the oracle:
    generate_model with model x - sin(t), Hyper(schedule), rng = random.random() seeded 42, 7 steps."""
    steps = 7
    sch = O.hyper(O.linear() if base == "linear" else O.scaled())
    model = O.Model(kind)
    random.seed(42)

    def net(x: float, t: float, s: float, a: float) -> float:
        return x - math.sin(t)

    sample = random.random()  # rng(None): functional.py:139-140
    seen: list[float] = []
    if sampler in ("Adams", "SPC"):
        # interface.py:37-59 - structured samplers consume no rng unless stochastic (they are not here)
        pts = sch.schedule(steps)
        previous: list[O.Rec] = []
        need_prev = 1 if sampler == "Adams" else 4
        for n, pt in enumerate(pts):
            cur = O.Rec(sample, net(sample, *pt), O.St.from_int(n, steps), None)
            if sampler == "Adams":
                cur.final = O.adams_step(cur, model, sch, previous, 2, 0, O.DATA)
            else:
                cur = O.spc_step(cur, model, sch, previous)
            previous.append(cur)
            previous = previous[max(len(previous) - need_prev, 0) :]
            sample = cur.final
            seen.append(sample)
        return seen
    for n in range(steps):
        step = O.St.from_int(n, steps)
        tab = O.HEUN if sampler == "RKUltra" else O.dynasaurk_tableau(step, 2)
        noise = random.random()  # rng(step) is drawn every step even for an ODE: functional.py:266,347
        sample = O.step_tableau(tab, sample, net, model, sch, step, O.DATA, noise, 0)[0]
        seen.append(sample)
    return seen


@pytest.mark.parametrize("key", KNOWN_TRAJECTORIES, ids=lambda k: "-".join(k))
def test_oracle_known_trajectories(key: tuple[str, str, str]) -> None:
    "reference: tests/self_sampling.py:86-104 (tolerance 1e-3 percent)"
    got = np.asarray(_known_trajectory(*key))
    want = np.asarray(KNOWN_TRAJECTORIES[key])
    assert np.all(np.abs(got - want) <= 1e-5 * np.abs(want)), (got, want)
