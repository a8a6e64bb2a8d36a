"""GraphedTrajectory: a captured trajectory replays the bits of the eager structured-sampler loop
(reference loop: skrample/sampling/interface.py:23-59)."""

from __future__ import annotations

import pytest
import torch

from skrample_b200 import scheduling
from skrample_b200.common import Step
from skrample_b200.sampling import models, structured


def test_graphed_trajectory_needs_device_tensors() -> None:
    from skrample_b200.graphs import GraphedTrajectory

    with pytest.raises(ValueError, match="CUDA"):
        GraphedTrajectory(structured.Euler(), models.FlowModel(), scheduling.Linear(), steps=4, like=torch.zeros(8))


CASES = [
    ("unipc3-sde-bf16", structured.UniPC(order=3, stochasticity=1), models.NoiseModel(), scheduling.Scaled(), torch.bfloat16),
    ("adams4-data-f32", structured.Adams(order=4), models.DataModel(), scheduling.Scaled(), torch.float32),  # identity conversion: history aliases the prediction ring
    ("euler-sde-flow-f32", structured.Euler(stochasticity=1), models.FlowModel(), scheduling.FlowShift(scheduling.Linear()), torch.float32),
    ("dpm3-sde-f16", structured.DPM(order=3, stochasticity=1), models.VelocityModel(), scheduling.Scaled(), torch.float16),
    ("spc-f32", structured.SPC(), models.NoiseModel(), scheduling.Scaled(), torch.float32),
]


@pytest.mark.gpu
@pytest.mark.parametrize(("name", "sampler", "model", "schedule", "dtype"), CASES, ids=[c[0] for c in CASES])
def test_graph_replay_equals_eager_loop(name: str, sampler, model, schedule, dtype: torch.dtype) -> None:  # noqa: ANN001
    from skrample_b200 import native
    from skrample_b200.graphs import GraphedTrajectory

    steps, shape = 9, (2, 4, 40, 33)  # ragged: exercises the tail path inside the graph too
    g = torch.Generator(device="cuda").manual_seed(11)
    traj = GraphedTrajectory(sampler, model, schedule, steps, like=torch.empty(shape, device="cuda", dtype=dtype))
    assert len(traj) == steps

    for attempt in range(2):  # a second trajectory through the same graphs, on new data
        x0 = torch.randn(shape, device="cuda", generator=g).to(dtype)
        preds = [torch.randn(shape, device="cuda", generator=g).to(dtype) for _ in range(steps)]
        noises = [torch.randn(shape, device="cuda", generator=g).to(dtype) for _ in range(steps)]

        x, previous, eager = x0, [], []
        for n in range(steps):
            res = sampler.sample(x, preds[n], Step.from_int(n, steps), model, schedule, noises[n] if sampler.require_noise else None, previous)
            previous = ([*previous, res])[-sampler.require_previous :] if sampler.require_previous else []
            x = res.final
            eager.append(x.clone())

        launches = native.launch_count()
        traj.start(x0)
        for n in range(steps):
            assert traj.position == n
            if n % 2:  # both ways of handing inputs over
                traj.prediction().copy_(preds[n])
                if sampler.require_noise:
                    traj.noise().copy_(noises[n])
                out = traj.step()
            else:
                out = traj.step(preds[n], noises[n] if sampler.require_noise else None)
            assert out.dtype == dtype
            assert torch.equal(out, eager[n]), f"attempt {attempt} step {n}: max diff {(out.float() - eager[n].float()).abs().max().item()}"
        assert native.launch_count() == launches, "replay must not go through the Python launch path"
        with pytest.raises(IndexError):
            traj.step()


@pytest.mark.gpu
def test_graphed_trajectory_points_and_include() -> None:
    from skrample_b200.graphs import GraphedTrajectory

    schedule = scheduling.Scaled()
    traj = GraphedTrajectory(structured.Euler(), models.NoiseModel(), schedule, 10, like=torch.empty(64, device="cuda"), include=slice(4, None))
    assert len(traj) == 6
    assert traj.point(0) == schedule.ipoint(0.4)
    with pytest.raises(ValueError, match="no noise"):
        traj.noise()
