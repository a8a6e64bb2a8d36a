"""Where one end-to-end step of bench.py's e2e leg spends its host time (development aid).

Phases, timed with perf_counter and no extra synchronisation: H2D enqueue of the prediction, noise generation
(enqueue), sampler.sample (plan lookup + launch), D2H enqueue, and the final wait for the stream."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
spec = bench.WORKLOADS[next((a for a in sys.argv[1:] if not a.startswith("--")), bench.DEFAULT_WORKLOAD)]
traj = bench.Trajectory(spec, dev, seed=4321)
traj.record()
host_pred = [p.cpu().pin_memory() for p in traj.predictions]
result_host = torch.empty(spec["shape"], dtype=traj.dtype).pin_memory()
names = ["h2d", "noise", "sample", "d2h", "wait"]
acc = dict.fromkeys(names, 0.0)


def one(record: bool) -> None:
    t = [time.perf_counter()]
    pred = host_pred[traj.n].to(dev, non_blocking=True)
    t.append(time.perf_counter())
    noise = None
    if traj.sampler.require_noise:
        noise = traj.noise_source.auto(None) if bench.SUPPLIED_NOISE else traj.noise_source.lazy(None)
    t.append(time.perf_counter())
    final = traj.step(pred, noise)
    t.append(time.perf_counter())
    result_host.copy_(final, non_blocking=True)
    t.append(time.perf_counter())
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())
    if record:
        for name, a, b in zip(names, t, t[1:]):
            acc[name] += b - a


for _ in range(50):
    one(False)
traj.reset()
N = 500
t0 = time.perf_counter()
for _ in range(N):
    one(True)
total = time.perf_counter() - t0
print(f"us/step {total / N * 1e6:.1f}  " + "  ".join(f"{k}={v / N * 1e6:.1f}" for k, v in acc.items()))
