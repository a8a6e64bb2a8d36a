"""Batch sharding across ranks: no collective on the step path, an all_gather only to validate (gloo, world size 2)."""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _trajectory(x: torch.Tensor, outs: list[torch.Tensor], noises: list[torch.Tensor]) -> torch.Tensor:
    from skrample_b200 import scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import models, structured

    sampler = structured.UniPC(order=3, stochasticity=1)
    schedule, model = scheduling.Scaled(), models.NoiseModel()
    previous: list = []
    for n, (out, noise) in enumerate(zip(outs, noises, strict=True)):
        res = sampler.sample(x, out, Step.from_int(n, len(outs)), model, schedule, noise, previous)
        previous = (previous + [res])[-sampler.require_previous :]
        x = res.final
    return x


def _inputs() -> tuple[torch.Tensor, list[torch.Tensor], list[torch.Tensor]]:
    g = torch.Generator().manual_seed(3)
    x = torch.randn((4, 4, 16, 16), generator=g)
    outs = [torch.randn((4, 4, 16, 16), generator=g) * 0.5 for _ in range(6)]
    noises = [torch.randn((4, 4, 16, 16), generator=g) for _ in range(6)]
    return x, outs, noises


def _worker(rank: int, world: int, port: int, result_path: str) -> None:
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, outs, noises = _inputs()
    per = x.shape[0] // world
    mine = slice(rank * per, (rank + 1) * per)
    local = _trajectory(x[mine], [o[mine] for o in outs], [z[mine] for z in noises])  # the step path: no communication
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)  # validation only
    if rank == 0:
        torch.save(torch.cat(gathered), result_path)
    dist.barrier()
    dist.destroy_process_group()


def test_batch_shards_reproduce_the_unsharded_run(tmp_path: Path) -> None:
    path = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), path), nprocs=2, join=True)
    sharded = torch.load(path)
    x, outs, noises = _inputs()
    assert torch.equal(sharded, _trajectory(x, outs, noises))


def _nccl_worker(rank: int, world: int, port: int, result_path: str) -> None:
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from skrample_b200.pytorch import noise as sk_noise

    device = torch.device("cuda", rank)
    x, outs, _ = _inputs()
    per = x.shape[0] // world
    mine = slice(rank * per, (rank + 1) * per)
    # per-item generators keyed by the GLOBAL item index: any sharding draws the same noise for an item
    source = sk_noise.BatchTensorNoise.from_batch_inputs(
        sk_noise.Random, tuple(x.shape[1:]), [torch.Generator(device=device).manual_seed(1000 + i) for i in range(mine.start, mine.stop)]
    )
    noises = [source.generate(None) for _ in outs]
    local = _trajectory(x[mine].to(device), [o[mine].to(device) for o in outs], noises)  # the step path: no communication
    gathered = torch.empty((x.shape[0], *local.shape[1:]), device=device)
    dist.all_gather_into_tensor(gathered, local)  # validation only (NCCL over NVLink)
    if rank == 0:
        torch.save(gathered.cpu(), result_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_batch_shards_reproduce_the_single_gpu_run(tmp_path: Path) -> None:
    "Two ranks, one GPU each, batch-sharded with device-side noise: the gathered result is the one-GPU result bit for bit."
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from skrample_b200.pytorch import noise as sk_noise

    path = str(tmp_path / "gathered.pt")
    mp.spawn(_nccl_worker, args=(2, _free_port(), path), nprocs=2, join=True)
    sharded = torch.load(path)
    device = torch.device("cuda", 0)
    x, outs, _ = _inputs()
    source = sk_noise.BatchTensorNoise.from_batch_inputs(
        sk_noise.Random, tuple(x.shape[1:]), [torch.Generator(device=device).manual_seed(1000 + i) for i in range(x.shape[0])]
    )
    noises = [source.generate(None) for _ in outs]
    whole = _trajectory(x.to(device), [o.to(device) for o in outs], noises)
    assert torch.equal(sharded, whole.cpu())


def test_bench_reference_arm_under_torchrun_prints_once() -> None:
    """The driver launches the reference arm like the product arm (torchrun, one rank per GPU): rank 0 alone measures
    and prints the line, the other ranks exit 0 without output."""
    import json
    import socket
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    with socket.socket() as probe:
        probe.bind(("127.0.0.1", 0))
        port = probe.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "25", "--warmup", "3"]
    done = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [line for line in done.stdout.splitlines() if line.strip()]
    assert len(lines) == 1, done.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0


def test_bench_leaves_a_process_group_without_hanging(tmp_path: Path) -> None:
    """bench.leave_process_group: every rank exits 0 promptly - also when one rank never reaches the teardown (the
    watchdog's case), which a plain destroy_process_group() would wait for."""
    import socket
    import subprocess
    import sys
    import time

    root = Path(__file__).resolve().parent.parent
    script = tmp_path / "leave.py"
    script.write_text(
        "import os, sys, time\n"
        f"sys.path.insert(0, {str(root)!r})\n"
        "import torch, torch.distributed as dist\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "t = torch.ones(1); dist.all_reduce(t); assert t.item() == 2\n"
        "print('rank', dist.get_rank(), 'done', flush=True)\n"
        "if os.environ.get('STRAGGLER') == '1' and dist.get_rank() == 1:\n"
        "    time.sleep(8); os._exit(0)\n"
        "bench.leave_process_group(grace_s=3.0)\n"
        "raise SystemExit('leave_process_group returned')\n"
    )
    for straggler in ("0", "1"):
        with socket.socket() as probe:
            probe.bind(("127.0.0.1", 0))
            port = probe.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
        t0 = time.perf_counter()
        done = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "STRAGGLER": straggler})
        assert done.returncode == 0, done.stderr[-2000:]
        assert done.stdout.count("done") == 2, done.stdout[-500:]
        assert time.perf_counter() - t0 < 120
