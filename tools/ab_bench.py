"""A/B timing of sweep workloads under the current environment (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
args = sys.argv[1:]
bench.SUPPLIED_NOISE = "--fused" not in args
names = [a for a in args if not a.startswith("--")] or ["unipc3_sde_flux_bf16"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
out = []
for name in names:
    if name not in bench.WORKLOADS:
        print(f"unknown workload {name}; known: {sorted(bench.WORKLOADS)}", file=sys.stderr)
        continue
    r = bench.graph_throughput(bench.WORKLOADS[name], dev, 200, 50, 2 * bench.L2_BYTES)
    us = r["elapsed_ms"] * 1e3 / r["launches"]
    gbs = r["bytes"] / (r["elapsed_ms"] * 1e-3) / 1e9
    out.append(f"{name.replace('_sde','')}={us:.1f}us/{gbs:.0f}GB/s")
    print(out[-1], flush=True)
