"""How fast does the block kernel move UniPC-3's traffic when there is almost nothing to compute?  (development aid)

A hand-written step program with UniPC-3 SDE's operand list on a Flux-sized bf16 latent - 4 bf16 + 4 fp32 inputs, two fp32
and one bf16 output, 34 bytes per element - but only a handful of multiply-adds per element, timed like bench.py's rows
(CUDA-graph chain over rotating buffer sets, working set > 2x L2).  If this runs at the HBM roofline, UniPC's distance
from it is arithmetic / issue; if it does not, it is the access pattern.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from skrample_b200 import native
from skrample_b200.sampling import program as pg

dev = torch.device("cuda", 0)
import os

shape = tuple(int(v) for v in os.environ.get("PROBE_SHAPE", "16,16,128,128").split(","))
sets = 8
g = torch.Generator(device=dev).manual_seed(0)


def tensors() -> dict:
    h = lambda: torch.randn(shape, device=dev, generator=g).bfloat16()  # noqa: E731
    f = lambda: torch.randn(shape, device=dev, generator=g)  # noqa: E731
    return dict(x=h(), y=h(), n1=h(), n2=h(), a=f(), b=f(), c=f(), d=f())


data = [tensors() for _ in range(sets)]


variant = next((a for a in sys.argv[1:] if not a.startswith("--")), "unipc")
# unipc: 2 fp32 + 1 bf16 outputs; one-out: the bf16 final only; two-f32: the two fp32 outputs only; half-state: all three bf16
OUT = {"unipc": ("compute", "compute", torch.bfloat16), "one-out": (None, None, torch.bfloat16), "two-f32": ("compute", "compute", None),
       "half-state": (torch.bfloat16, torch.bfloat16, torch.bfloat16)}[variant]


def step(t: dict) -> list:
    p = pg.Program()
    p.load(pg.X, t["x"])
    p.load(pg.P, t["y"])
    if OUT[0] is not None:
        p.store(pg.P, OUT[0])
    p.load(pg.X, t["a"])
    p.acc(0.5, reg=pg.P, first=True)
    p.acc(0.25, t["b"])
    p.acc(0.125, t["c"])
    p.fwd(0.9, 0.1, pg.A, t["n1"], 0.3)
    if OUT[1] is not None:
        p.store(pg.R, OUT[1])
    p.mov(pg.X, pg.R)
    p.acc(0.5, reg=pg.P, first=True)
    p.acc(0.25, t["d"])
    p.fwd(0.8, 0.2, pg.A, t["n2"], 0.1)
    if OUT[2] is not None:
        p.store(pg.R, OUT[2])
    else:
        p.mov(pg.X, pg.R)
    return p.run()


numel = 1
for extent in shape:
    numel *= extent
nbytes = numel * (4 * 2 + 4 * 4 + sum(0 if o is None else (2 if o is torch.bfloat16 else 4) for o in OUT))
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    for t in data:
        step(t)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(graph, stream=stream):
        for _ in range(5):
            for t in data:
                keep.append(step(t))
    for _ in range(3):
        graph.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        graph.replay()
    b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / (20 * 5 * sets) * 1e3
print(f"{variant} {shape}: UniPC-3-shaped reads, trivial arithmetic: {us:.2f} us per launch, {nbytes / us / 1e3:.0f} GB/s ({nbytes / 1e6:.1f} MB)")
