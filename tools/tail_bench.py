"""Cost of a ragged tail: the same UniPC-3 SDE bf16 step on flat latents of 511 tiles, 511 tiles + 992 elements, 512 tiles.
CUDA-graph replay of one 25-step trajectory per size (16 interleaved latents).  Development aid."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for numel in (511 * 1024, 511 * 1024 + 992, 511 * 1024 + 5, 512 * 1024):
    spec = dict(sampler="UniPC", kw={"order": 3, "stochasticity": 1}, schedule="scaled", model="NoiseModel", shape=(1, numel), dtype="bf16")
    r = bench.graph_throughput(spec, dev, 200, 50, 2 * bench.L2_BYTES)
    print(f"numel {numel}: {r['elapsed_ms'] * 1e3 / r['launches']:.2f} us/step")
