"""Device-side time of one explicit Runge-Kutta step (BASELINE.json configs[2]): RKUltra(order=4) = EES(2,7) 4 stages,
FlowShift(Linear(), 3), FlowModel, 16x16x128x128 latents.  The network is a table of pre-recorded outputs (zero cost), so
the graph holds only this library's launches; 16 interleaved latents keep the inputs out of L2.  Development aid."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from skrample_b200 import native, scheduling
from skrample_b200.common import Step
from skrample_b200.sampling import functional, models

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[next((a for a in sys.argv[1:] if a in ("bf16", "f32")), "bf16")]
order = int(next((a for a in sys.argv[1:] if a.isdigit()), "4"))
shape = (16, 16, 128, 128)
replicas, steps = 16, 25
sampler = functional.RKUltra(order=order)
schedule, model_transform = scheduling.FlowShift(scheduling.Linear(), shift=3.0), models.FlowModel()
g = torch.Generator(device=dev).manual_seed(0)
latents = [torch.randn(shape, device=dev, generator=g).to(dtype) for _ in range(replicas)]
outputs = [[torch.randn(shape, device=dev, generator=g).to(dtype) * 0.3 for _ in range(8)] for _ in range(replicas)]


def run() -> None:
    for r in range(replicas):
        calls = iter(range(100))
        table = outputs[r]
        x = latents[r]
        for n in (3, 12):  # two interior steps of the schedule per replica
            x = sampler.step(x, lambda s, t, sig, alp: table[next(calls) % 8], model_transform, schedule, Step.from_int(n, steps))


stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    run()
    torch.cuda.synchronize()
    before = native.launch_count()
    native.ACCOUNT["bytes"] = 0
    native.ACCOUNT["on"] = True
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        run()
    native.ACCOUNT["on"] = False
    launches = native.launch_count() - before
torch.cuda.synchronize()
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
rk_steps = replicas * 2
print(
    f"RKUltra({order}) {dtype}: {ms * 1e3 / rk_steps:.1f} us per RK step, {launches / rk_steps:.1f} launches/step, "
    f"{ms * 1e3 / launches:.1f} us/launch, {native.ACCOUNT['bytes'] / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic"
)

# eager wall time of the same step (host emission + launches), to see how far the host is behind the device
import time

x = latents[0]
table = outputs[0]
calls = iter(range(1 << 30))
model = lambda s, t, sig, alp: table[next(calls) % 8]  # noqa: E731
for _ in range(20):
    sampler.step(x, model, model_transform, schedule, Step.from_int(3, steps))
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 200
for _ in range(n):
    sampler.step(x, model, model_transform, schedule, Step.from_int(3, steps))
torch.cuda.synchronize()
print(f"eager: {(time.perf_counter() - t0) / n * 1e6:.1f} us per RK step wall")
