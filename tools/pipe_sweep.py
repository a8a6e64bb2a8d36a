"""Pipeline-shape sweep of one bench workload (development aid): SKR_CTAS x SKR_STAGES x arithmetic.

    python tools/pipe_sweep.py [workload]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from skrample_b200 import native

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
name = next((a for a in sys.argv[1:] if not a.startswith("--")), "unipc3_sde_flux_bf16")
spec = bench.WORKLOADS[name]
peak, _ = bench.measured_peak()
for arith in ("exact", "contracted"):
    for ctas, stages in ((0, 0), (4, 2), (3, 2), (3, 3), (2, 3), (2, 4), (2, 5), (1, 8)):
        os.environ["SKR_CTAS"] = str(ctas)
        os.environ["SKR_STAGES"] = str(stages)
        native.reset_switches()
        native.set_arithmetic(arith)
        torch.cuda.empty_cache()
        try:
            r = bench.chain_time(spec, dev, "supplied", 0, 25, min_seconds=0.1, blocks=3)
        except RuntimeError as error:
            print(f"{arith:10s} ctas={ctas} stages={stages}: {str(error)[:80]}")
            continue
        us = r["ms_per_step"] * 1e3
        print(f"{arith:10s} ctas={ctas} stages={stages}: {us:7.2f} us/step  {r['bytes_per_step_avg'] / (us * 1e-6) / 1e9 / peak:.3f} of peak", flush=True)
native.set_arithmetic("exact")
