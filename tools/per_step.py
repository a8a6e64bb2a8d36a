"""Per-step device time of one bench workload's 25-step trajectory (development aid): which steps of the order ramp
cost what, with the kernel shape each runs on.  CUDA events between consecutive launches, several replicas interleaved
so inputs come from HBM, median over repetitions.

    python tools/per_step.py [workload] [--contracted]
"""
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from skrample_b200 import native

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
name = next((a for a in sys.argv[1:] if not a.startswith("--")), "unipc3_sde_flux_bf16")
if "--contracted" in sys.argv:
    native.set_arithmetic("contracted")
spec = bench.WORKLOADS[name]
peak, _ = bench.measured_peak()
per_bytes = bench.step_bytes(spec, dev, "supplied")
replicas = 6
trajs = [bench.Trajectory(spec, dev, seed=50 + i, keep=25, noise="supplied") for i in range(replicas)]
for t in trajs:
    t.record()
shapes = []
real = native.CompiledProgram.specialise


def spy(self, signature):  # noqa: ANN001, ANN201
    plan = real(self, signature)
    shapes.append(plan.shape_name)
    return plan


native.CompiledProgram.specialise = spy
for t in trajs[:1]:
    for _ in range(25):
        t.step()
native.CompiledProgram.specialise = real
for t in trajs[1:]:
    for _ in range(25):
        t.step()
torch.cuda.synchronize()
# one CUDA graph per step index holding that step of every replica: device time without host launch gaps
stream = torch.cuda.Stream(device=dev)
graphs = []
with torch.cuda.stream(stream):
    for n in range(25):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for t in trajs:
                t.step()
        graphs.append(graph)
    times = [[] for _ in range(25)]
    events = [torch.cuda.Event(enable_timing=True) for _ in range(26)]
    for rep in range(15):
        events[0].record()
        for n in range(25):
            graphs[n].replay()
            events[n + 1].record()
        torch.cuda.synchronize()
        if rep >= 3:
            for n in range(25):
                times[n].append(events[n].elapsed_time(events[n + 1]) * 1e3 / replicas)
print(f"{name}: step, shape, MB, median us per launch (graph replay), TB/s, of peak")
for n in range(25):
    us = statistics.median(times[n])
    shape = shapes[n] if n < len(shapes) else "?"
    print(f"{n:2d} {shape:28s} {per_bytes[n] / 1e6:7.1f} {us:7.2f} {per_bytes[n] / us / 1e6:6.2f} {per_bytes[n] / us / 1e3 / peak:5.2f}")
total_us = sum(statistics.median(t) for t in times)
print(f"trajectory: {sum(per_bytes) / 1e6:.0f} MB in {total_us:.1f} us = {sum(per_bytes) / total_us / 1e6:.2f} TB/s ({sum(per_bytes) / total_us / 1e3 / peak:.3f} of peak)")
