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
    spec = dict(sampler="UniPC", kw={"order": 3, "stochasticity": 1}, schedule="scaled", model="NoiseModel", shape=(1, numel), dtype="bf16", noise="Random")
    r = bench.chain_time(spec, dev, "supplied", 0, 25, min_seconds=0.1, blocks=3)
    print(f"numel {numel}: {r['ms_per_step'] * 1e3:.2f} us/step")
