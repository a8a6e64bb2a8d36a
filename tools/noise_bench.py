"""Device time of the noise generators on the BASELINE.json configs[3] unit shape (16x21x90x160 video latent) and on a
Flux-sized batch.  CUDA events around `generate`, after warm-up; algorithmic bytes = the tensor written once.
Development aid; the numbers are quoted in profiles/."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from skrample_b200.common import Step
from skrample_b200.pytorch import noise

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
only = set(sys.argv[1:])  # e.g. `noise_bench.py Pyramid` times only that generator


def timed(make, reps=30):
    src = make()
    step = Step.from_int(5, 25)
    for _ in range(20):  # the first calls in a process pay for cuFFT plans, allocator growth and lazy module loading
        src.generate(step)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        out = src.generate(step)
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / reps * 1e3, out


for label, unit in (("video 16x21x90x160", (16, 21, 90, 160)), ("flux item 16x128x128", (16, 128, 128))):
    for dtype in (torch.float32, torch.bfloat16):
        rows = [
            ("Random", lambda: noise.Random.from_inputs(unit, torch.Generator(device=dev).manual_seed(1), dtype=dtype)),
            ("Offset", lambda: noise.Offset.from_inputs(unit, torch.Generator(device=dev).manual_seed(1), noise.OffsetProps(), dtype=dtype)),
            ("Pyramid", lambda: noise.Pyramid.from_inputs(unit, torch.Generator(device=dev).manual_seed(1), noise.PyramidProps(), dtype=dtype)),
            ("Colored", lambda: noise.Colored.from_inputs(unit, torch.Generator(device=dev).manual_seed(1), noise.ColoredProps(), dtype=dtype)),
            ("Brownian", lambda: noise.Brownian.from_inputs(unit, torch.Generator(device=dev).manual_seed(1), noise.BrownianProps(), dtype=dtype)),
        ]
        for name, make in rows:
            if only and name not in only:
                continue
            us, out = timed(make)
            nbytes = out.numel() * out.element_size()
            print(f"{label:22s} {str(dtype).replace('torch.', ''):9s} {name:8s} {us:9.1f} us  {nbytes / us / 1e3:8.1f} GB/s written  ({out.numel() / us:8.1f} Melem/s)")
if only:
    sys.exit(0)
batch = noise.BatchTensorNoise.from_batch_inputs(noise.Random, (16, 128, 128), [torch.Generator(device=dev).manual_seed(i) for i in range(16)], dtype=torch.float32)
us, out = timed(lambda: batch)
print(f"{'flux batch 16x16x128x128':22s} float32   Random x16 {us:7.1f} us  {out.numel() * 4 / us / 1e3:8.1f} GB/s written")
