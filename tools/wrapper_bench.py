"""Host + device time of one diffusers-wrapper step (development aid): SkrampleWrapperScheduler around UniPC-3 SDE on an
SDXL-sized bf16 latent, timesteps taken from scheduler.timesteps (device tensor views), per-item CUDA generators."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from skrample_b200 import diffusers, scheduling
from skrample_b200.sampling import structured

dev = torch.device("cuda", 0)
shape = (8, 4, 128, 128)
wrapper = diffusers.SkrampleWrapperScheduler(sampler=structured.UniPC(order=3, stochasticity=1), schedule=scheduling.Scaled())
generators = [torch.Generator(device=dev).manual_seed(i) for i in range(shape[0])]
outputs = [torch.randn(shape, device=dev).bfloat16() for _ in range(4)]


def trajectory(steps: int = 25) -> torch.Tensor:
    wrapper.set_timesteps(steps, device=dev)
    x = torch.randn(shape, device=dev).bfloat16()
    for n, t in enumerate(wrapper.timesteps):
        x = wrapper.step(outputs[n % 4], t, x, generator=generators, return_dict=False)[0]
    return x


for _ in range(3):
    trajectory()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    trajectory()
torch.cuda.synchronize()
print(f"wrapper.step: {(time.perf_counter() - t0) / (reps * 25) * 1e6:.1f} us per step (wall, device-resident inputs)")

if "--profile" in sys.argv:
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    for _ in range(8):
        trajectory()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
