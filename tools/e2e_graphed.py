"""End-to-end step time through GraphedTrajectory with host buffers (development aid; compare tools/e2e_breakdown.py)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from skrample_b200.graphs import GraphedTrajectory

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
spec = bench.WORKLOADS[next((a for a in sys.argv[1:] if not a.startswith("--")), bench.DEFAULT_WORKLOAD)]
traj = bench.Trajectory(spec, dev, seed=4321)
traj.record()
host_pred = [p.cpu().pin_memory() for p in traj.predictions]
result_host = torch.empty(spec["shape"], dtype=traj.dtype).pin_memory()
graphed = GraphedTrajectory(traj.sampler, traj.model, traj.schedule, bench.STEPS_PER_TRAJECTORY, like=traj.x0)
names = ["h2d", "noise", "step", "d2h", "wait"]
acc = dict.fromkeys(names, 0.0)
noise32 = torch.empty(spec["shape"], device=dev, dtype=torch.float32)


def one(record: bool) -> None:
    if graphed.position == len(graphed):
        graphed.start(traj.x0)
    n = graphed.position
    t = [time.perf_counter()]
    graphed.prediction().copy_(host_pred[n], non_blocking=True)
    t.append(time.perf_counter())
    if traj.sampler.require_noise:
        graphed.noise().copy_(traj.noise_source.generate(None))
    t.append(time.perf_counter())
    final = graphed.step()
    t.append(time.perf_counter())
    result_host.copy_(final, non_blocking=True)
    t.append(time.perf_counter())
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())
    if record:
        for name, a, b in zip(names, t, t[1:]):
            acc[name] += b - a


graphed.start(traj.x0)
for _ in range(50):
    one(False)
graphed.start(traj.x0)
N = 500
t0 = time.perf_counter()
for _ in range(N):
    one(True)
total = time.perf_counter() - t0
print(f"graphed us/step {total / N * 1e6:.1f}  " + "  ".join(f"{k}={v / N * 1e6:.1f}" for k, v in acc.items()))
