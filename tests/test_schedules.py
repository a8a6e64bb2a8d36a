"""sigma / alpha / timestep values are bit-identical to the reference (north star: 'indexing bit-exact')."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

import cases
from skrample_b200 import scheduling

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "schedules.json").read_text())


def _decode(rows: list[list[str]]) -> np.ndarray:
    return np.asarray([[float.fromhex(v) for v in row] for row in rows], dtype=np.float64)


@pytest.mark.parametrize("n", range(len(cases.SCHEDULE_CASES)), ids=lambda n: str(cases.SCHEDULE_CASES[n]))
def test_points_bit_exact(n: int) -> None:
    sch = cases.make_schedule_stack(scheduling, cases.SCHEDULE_CASES[n])
    want = GOLDEN[str(n)]
    with np.errstate(all="ignore"):
        assert np.array_equal(sch.points_np(cases.SCHEDULE_TIMES), _decode(want["points"]), equal_nan=True)
        assert np.array_equal(sch.ipoints_np(cases.SCHEDULE_TIMES), _decode(want["ipoints"]), equal_nan=True)
        assert np.array_equal(sch.schedule_np(7), _decode(want["schedule7"]), equal_nan=True)
        assert np.array_equal(sch.schedule_np(25), _decode(want["schedule25"]), equal_nan=True)
        # batch == single point, exactly (reference: tests/self_scheduling.py:70-78)
        single = np.asarray([sch.point(t) for t in cases.SCHEDULE_TIMES], dtype=np.float64)
        assert np.array_equal(sch.points_np(cases.SCHEDULE_TIMES), single, equal_nan=True)


def test_timestep_lookup_is_exact() -> None:
    "The wrapper finds a step by exact float match of the timestep (reference: diffusers.py:565-567)."
    from skrample_b200.diffusers import SkrampleWrapperScheduler
    from skrample_b200.sampling.structured import DPM

    w = SkrampleWrapperScheduler(DPM(), scheduling.Karras(scheduling.Scaled()))
    w.set_timesteps(25)
    for n, t in enumerate(w.timesteps):
        assert w._index_of(t) == n
        assert w._index_of(t.item()) == n


def test_modifier_split_and_stack_roundtrip() -> None:
    sch = scheduling.Hyper(scheduling.FlowShift(scheduling.Sinner(scheduling.Karras(scheduling.Scaled()))))
    mods, sub, base = sch.all_split
    assert [type(m).__name__ for m in mods] == ["Hyper", "FlowShift", "Sinner"]
    assert type(sub).__name__ == "Karras" and type(base).__name__ == "Scaled"
    assert scheduling.ScheduleModifier.stack(mods, sub, base) == sch
    before, hit, after, _, _ = sch.find_split(scheduling.FlowShift)
    assert [type(m).__name__ for m in before] == ["Hyper"] and [type(m).__name__ for m in after] == ["Sinner"]
    assert hit.shift == 3.0
