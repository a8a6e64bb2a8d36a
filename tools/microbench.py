"""Quick device-side bandwidth probe of the step kernel (development aid, not the bench contract)."""

from __future__ import annotations

import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from skrample_b200 import native, scheduling  # noqa: E402
from skrample_b200.common import Step  # noqa: E402
from skrample_b200.sampling import models, structured  # noqa: E402


def time_step(sampler, model, schedule, shape, dtype, steps=25, reps=5):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    ring = 6
    xs = [torch.randn(shape, device=dev, generator=g).to(dtype) for _ in range(ring)]
    outs = [torch.randn(shape, device=dev, generator=g).to(dtype) for _ in range(ring)]
    noises = [torch.randn(shape, device=dev, generator=g).to(dtype) for _ in range(ring)]
    best = 1e9
    total_bytes = 0
    for rep in range(reps):
        prev = []
        x = xs[0]
        per_step = []
        for n in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = sampler.sample(x, outs[n % ring], Step.from_int(n, steps), model, schedule, noises[n % ring] if sampler.require_noise else None, prev)
            e1.record()
            per_step.append((e0, e1))
            prev = (prev + [res])[-sampler.require_previous :] if sampler.require_previous else []
            x = res.final
        torch.cuda.synchronize()
        t = sum(a.elapsed_time(b) for a, b in per_step)
        best = min(best, t)
    return best / steps


def main() -> None:
    print(torch.cuda.get_device_name(0))
    flow = scheduling.FlowShift(scheduling.Linear())
    for name, sampler, passes in [
        ("euler_sde", structured.Euler(stochasticity=1), 4),
        ("dpm2", structured.DPM(order=2), 5),
        ("adams9_sde", structured.Adams(order=9, stochasticity=1), 13),
        ("unipc3_sde", structured.UniPC(order=3, stochasticity=1), 11),
    ]:
        for shape in [(1, 4, 128, 128), (16, 16, 128, 128), (64, 16, 128, 128), (256, 16, 128, 128)]:
            for dtype in (torch.float32, torch.bfloat16):
                n = 1
                for s in shape:
                    n *= s
                try:
                    ms = time_step(sampler, models.FlowModel(), flow, shape, dtype)
                except torch.OutOfMemoryError:
                    continue
                gb = passes * n * (4 if dtype == torch.float32 else 2) / 1e9
                print(f"{name:12s} {str(shape):22s} {str(dtype):15s} {ms * 1e3:9.1f} us/step  ~{gb / (ms / 1e3):8.0f} GB/s (nominal {passes} passes)")


if __name__ == "__main__":
    main()
