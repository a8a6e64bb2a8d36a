"""Host cost of the pieces of one plan-cache hit of ``sampler.sample`` on the GPU box (development aid).

Every piece is timed back to back (hot caches, the stream drained only between batches of 25), so the numbers are the
enqueue costs a pipelined caller pays; tools/e2e_breakdown.py shows the same phases inside a synchronous loop.

    python tools/hit_breakdown.py [workload]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from skrample_b200 import native
from skrample_b200.common import Step
from skrample_b200.sampling import plan, structured
from skrample_b200.sampling.structured import SampleInput

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
spec = bench.WORKLOADS[next((a for a in sys.argv[1:] if not a.startswith("--")), bench.DEFAULT_WORKLOAD)]
traj = bench.Trajectory(spec, dev, seed=4321, noise="none")
traj.record()
T = bench.STEPS_PER_TRAJECTORY
steps = [Step.from_int(n, T) for n in range(T)]
sampler, model, schedule = traj.sampler, traj.model, traj.schedule
keep = max(sampler.require_previous, 1)
REPS = 200


def timed(label: str, body, per: int = T) -> None:  # noqa: ANN001
    for _ in range(5):
        body()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(REPS):
        body()
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:58s} {dt / (REPS * per) * 1e6:7.2f} us", flush=True)


def noise_only() -> None:
    for n in range(T):
        traj.noise_source.auto(steps[n])


def full() -> None:
    x, previous = traj.x0, []
    for n in range(T):
        noise = traj.noise_source.auto(steps[n])
        res = sampler.sample(x, traj.predictions[n], steps[n], model, schedule, noise, previous)
        previous = (previous + [res])[-keep:]
        x = res.final


fixed_noise = [traj.noise_source.auto(steps[n]) for n in range(T)]


def sample_only() -> None:
    x, previous = traj.x0, []
    for n in range(T):
        res = sampler.sample(x, traj.predictions[n], steps[n], model, schedule, fixed_noise[n], previous)
        previous = (previous + [res])[-keep:]
        x = res.final


# a steady-state step captured for the piecewise timings (every step of the trajectory has its plan after one pass)
sample_only()
x, previous = traj.x0, []
for n in range(6):
    res = sampler.sample(x, traj.predictions[n], steps[n], model, schedule, fixed_noise[n], previous)
    previous = (previous + [res])[-keep:]
    x = res.final
n = 6
packed = SampleInput(x, traj.predictions[n], steps[n], fixed_noise[n])
forced = (x.dtype, False)


def key_and_lookup() -> None:
    for _ in range(T):
        key = plan.key_for(sampler, packed, model, schedule, previous, forced)
        plan.lookup(key, sampler, model, schedule)


key = plan.key_for(sampler, packed, model, schedule, previous, forced)
hit = plan.lookup(key, sampler, model, schedule)
fast = native._fast_module()


def hit_only() -> None:
    for _ in range(T):
        fast.hit(hit.roles, hit.compiled.n_inputs, hit.compiled.n_philox, packed, previous, hit.compiled.fast, False)


def empties() -> None:
    for _ in range(T):
        for _ in hit.compiled.out_specs:
            torch.empty_like(x)


def packing() -> None:
    for _ in range(T):
        SampleInput(x, traj.predictions[n], steps[n], fixed_noise[n])


print(f"workload {spec['name'] if 'name' in spec else ''} outputs per step: {len(hit.compiled.out_specs)}; philox draws: {hit.compiled.n_philox}; inputs: {hit.compiled.n_inputs}")
timed("noise keys (BatchTensorNoise.auto)", noise_only)
timed("sampler.sample (hit) with its bookkeeping", sample_only)
timed("noise keys + sampler.sample", full)
timed("  SampleInput(...)", packing)
timed("  plan.key_for + plan.lookup", key_and_lookup)
timed("  _fast.hit (bind, key tables, allocate outputs, launch)", hit_only)
timed("  torch.empty_like x outputs", empties)
