"""Where one end-to-end step of bench.py's e2e leg spends its host time (development aid).

Phases, timed with perf_counter and no extra synchronisation: H2D enqueue of the prediction, the noise keys / draw,
sampler.sample (plan lookup + launch), D2H enqueue, and the final wait for the stream.  Prints one table row per
variant: the eager public API and GraphedTrajectory.

    python tools/e2e_breakdown.py [workload] [--profile]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from skrample_b200.common import Step
from skrample_b200.graphs import GraphedTrajectory

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
spec = bench.WORKLOADS[next((a for a in sys.argv[1:] if not a.startswith("--")), bench.DEFAULT_WORKLOAD)]
traj = bench.Trajectory(spec, dev, seed=4321, noise="none")  # the noise is drawn here, as its own phase
traj.record()
host_pred = [p.cpu().pin_memory() for p in traj.predictions]
pred_dev = torch.empty(spec["shape"], dtype=traj.dtype, device=dev)
result_host = torch.empty(spec["shape"], dtype=traj.dtype).pin_memory()
names = ["h2d", "noise", "sample", "d2h", "wait"]
need_noise = traj.sampler.require_noise
N = 2000


def eager(acc: dict | None) -> None:
    t = [time.perf_counter()]
    pred_dev.copy_(host_pred[traj.n], non_blocking=True)
    t.append(time.perf_counter())
    noise = traj.noise_source.auto(Step.from_int(traj.n, bench.STEPS_PER_TRAJECTORY)) if need_noise else None
    t.append(time.perf_counter())
    final = traj.step(pred_dev, noise)
    t.append(time.perf_counter())
    result_host.copy_(final, non_blocking=True)
    t.append(time.perf_counter())
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())
    if acc is not None:
        for name, a, b in zip(names, t, t[1:]):
            acc[name] += b - a


graphed = GraphedTrajectory(traj.sampler, traj.model, traj.schedule, bench.STEPS_PER_TRAJECTORY, like=traj.x0)
graphed.start(traj.x0)


def replayed(acc: dict | None) -> None:
    t = [time.perf_counter()]
    if graphed.position == len(graphed):
        graphed.start(traj.x0)
    graphed.prediction().copy_(host_pred[graphed.position], non_blocking=True)
    t.append(time.perf_counter())
    if need_noise:
        traj.noise_source.generate_into(graphed.noise(), None)
    t.append(time.perf_counter())
    final = graphed.step()
    t.append(time.perf_counter())
    result_host.copy_(final, non_blocking=True)
    t.append(time.perf_counter())
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())
    if acc is not None:
        for name, a, b in zip(names, t, t[1:]):
            acc[name] += b - a


print("| variant | us/step | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for label, one in (("sampler.sample (plan-cache hit)", eager), ("GraphedTrajectory.step", replayed)):
    for _ in range(100):
        one(None)
    traj.reset()
    acc = dict.fromkeys(names, 0.0)
    t0 = time.perf_counter()
    for _ in range(N):
        one(acc)
    total = time.perf_counter() - t0
    print(f"| {label} | {total / N * 1e6:.1f} | " + " | ".join(f"{acc[k] / N * 1e6:.1f}" for k in names) + " |")

if "--profile" in sys.argv:
    import cProfile
    import pstats

    traj.reset()
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(N):
        eager(None)
    prof.disable()
    pstats.Stats(prof).sort_stats("tottime").print_stats(25)
