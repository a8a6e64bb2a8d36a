"""A/B timing of bench workloads under the current environment / library build (development aid).

    [SKRAMPLE_B200_LIB=path/to/variant.so] python tools/ab_bench.py [workload ...] [--noise] [--contracted]

Prints microseconds per step and algorithmic GB/s of each workload's graph-replayed chain (bench.chain_time): with the
noise tensors written beforehand by default, with the step's noise drawn in the timed region for --noise.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from skrample_b200 import native

args = sys.argv[1:]
names = [a for a in args if not a.startswith("--")] or ["unipc3_sde_flux_bf16"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if "--contracted" in args:
    native.set_arithmetic("contracted")
for name in names:
    if name not in bench.WORKLOADS:
        print(f"unknown workload {name}; known: {sorted(bench.WORKLOADS)}", file=sys.stderr)
        continue
    r = bench.chain_time(bench.WORKLOADS[name], dev, "auto" if "--noise" in args else "supplied", 0, 25, min_seconds=0.15, blocks=3)
    us = r["ms_per_step"] * 1e3
    print(f"{name}={us:.2f}us/{r['bytes_per_step_avg'] / us / 1e3:.0f}GB/s", flush=True)
