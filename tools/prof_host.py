"""cProfile of the end-to-end (host-buffer) step loop of bench.py (development aid)."""
import cProfile, pstats, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
spec = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
bench.e2e_throughput(spec, dev, 100, 25)
pr = cProfile.Profile()
pr.enable()
r = bench.e2e_throughput(spec, dev, 400, 25)
pr.disable()
print("us/step", r["elapsed_s"] / 400 * 1e6)
pstats.Stats(pr).sort_stats("tottime").print_stats(32)
