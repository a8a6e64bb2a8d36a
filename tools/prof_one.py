"""Run a few steps of one sampler configuration (profiling target for ncu)."""

from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from skrample_b200 import scheduling  # noqa: E402
from skrample_b200.common import Step  # noqa: E402
from skrample_b200.sampling import models, structured  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sampler", default="euler")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--lazy-noise", action="store_true", help="hand the sampler Philox keys (BatchTensorNoise.auto): the step draws its noise itself where it chooses to")
args = ap.parse_args()

sampler = {
    "euler": structured.Euler(stochasticity=1),
    "adams9": structured.Adams(order=9, stochasticity=1),
    "unipc3": structured.UniPC(order=3, stochasticity=1),
    "dpm2": structured.DPM(order=2),
}[args.sampler]
dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
shape = (args.batch, 16, 128, 128)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(shape, device="cuda", generator=g).to(dtype)
outs = [torch.randn(shape, device="cuda", generator=g).to(dtype) for _ in range(2)]
noise = torch.randn(shape, device="cuda", generator=g).to(dtype)
flow = scheduling.FlowShift(scheduling.Linear())
prev: list = []
source = None
if args.lazy_noise:
    from skrample_b200.pytorch import noise as sk_noise

    source = sk_noise.BatchTensorNoise.from_batch_inputs(sk_noise.Random, shape[1:], [torch.Generator(device="cuda").manual_seed(100 + i) for i in range(shape[0])])
for n in range(args.steps):
    if source is not None:
        noise = source.auto(Step.from_int(n, 25))
    res = sampler.sample(x, outs[n % 2], Step.from_int(n, 25), models.FlowModel(), flow, noise if sampler.require_noise else None, prev)
    prev = (prev + [res])[-sampler.require_previous :] if sampler.require_previous else []
    x = res.final
torch.cuda.synchronize()
print("done", float(x.float().abs().mean()))
