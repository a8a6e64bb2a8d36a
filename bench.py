#!/usr/bin/env python
"""Benchmark of the skrample_b200 sampler step (BASELINE.json metric: sampler-step GB/s and latent-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--sweep]

A "step" is ONE solver step of the workload's sampler over one latent batch = one fused kernel launch.
The default workload is BASELINE.json configs[1]: UniPC order 3, stochastic, Random noise (device Philox), Scaled
schedule, epsilon model, SDXL latent 8x4x128x128, bf16 storage / fp32 compute, on one B200.

  value      whole-job latent-steps/s (batch items x solver steps per second; ``sampler_step_GBps`` is the same run
             as algorithmic GB/s), inputs resident in HBM, launches replayed from a CUDA graph (device-side time
             between two events; max over ranks).  Trajectories of several latent
             batches are interleaved so consecutive launches never touch the same buffers and the working set
             (> 2x L2) comes from HBM; the batches are independent requests, so the graph runs them on ``--streams``
             parallel branches (default 4).  ``one_stream`` is the same launches as a single chain.
  e2e        the same steps through the public API (``sampler.sample``) with HOST buffers: per step the model
             prediction is copied from pinned host memory, the noise is drawn on the device and the result is read
             back.  ``--inflight`` independent requests (default 2) are advanced round robin, each waiting for its own
             previous result; ``e2e.one_request`` is the plain synchronous loop.
  roofline   algorithmic bytes per launch / average launch duration of the step kernel (taken on the single chain,
             where a launch's duration is well defined) vs the measured HBM copy
             peak (MEASURED_PEAKS.json); ``sweep`` repeats that for the larger BASELINE shapes.
  cpu_baseline  the CPU oracle port of the reference algorithm (oracle/skrample_oracle.py on torch-CPU tensors,
             all host threads) on a bounded sample of the same workload.

Under torchrun every rank runs the same per-GPU batch on its own GPU (weak scaling, no collective on the step
path); rank 0 prints one JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for extra in (ROOT, ROOT / "tests"):
    if str(extra) not in sys.path:
        sys.path.insert(0, str(extra))

import torch  # noqa: E402

STEPS_PER_TRAJECTORY = 25
L2_BYTES = 126 * 1024 * 1024

WORKLOADS = {
    # BASELINE.json configs[1]
    "unipc3_sde_sdxl_bf16": dict(sampler="UniPC", kw={"order": 3, "stochasticity": 1}, schedule="scaled", model="NoiseModel", shape=(8, 4, 128, 128), dtype="bf16"),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "dpm2_scaled_fp32": dict(sampler="DPM", kw={"order": 2}, schedule="scaled", model="NoiseModel", shape=(1, 4, 128, 128), dtype="f32"),
    # BASELINE.json configs[3], one GPU's shard (one item of 8x16x21x90x160)
    "adams9_sde_video_bf16": dict(sampler="Adams", kw={"order": 9, "stochasticity": 1}, schedule="flow", model="FlowModel", shape=(1, 16, 21, 90, 160), dtype="bf16"),
    "adams9_sde_video_f32": dict(sampler="Adams", kw={"order": 9, "stochasticity": 1}, schedule="flow", model="FlowModel", shape=(1, 16, 21, 90, 160), dtype="f32"),
    # BASELINE.json configs[4] end points of the sweep
    "euler_sde_flow_f32_16": dict(sampler="Euler", kw={"stochasticity": 1}, schedule="flow", model="FlowModel", shape=(16, 16, 128, 128), dtype="f32"),
    "euler_sde_flow_f32_64": dict(sampler="Euler", kw={"stochasticity": 1}, schedule="flow", model="FlowModel", shape=(64, 16, 128, 128), dtype="f32"),
    "euler_sde_flow_f32_256": dict(sampler="Euler", kw={"stochasticity": 1}, schedule="flow", model="FlowModel", shape=(256, 16, 128, 128), dtype="f32"),
    "euler_sde_flow_bf16_256": dict(sampler="Euler", kw={"stochasticity": 1}, schedule="flow", model="FlowModel", shape=(256, 16, 128, 128), dtype="bf16"),
    "unipc3_sde_flux_bf16": dict(sampler="UniPC", kw={"order": 3, "stochasticity": 1}, schedule="flow", model="FlowModel", shape=(16, 16, 128, 128), dtype="bf16"),
    "unipc3_sde_ragged_bf16": dict(sampler="UniPC", kw={"order": 3, "stochasticity": 1}, schedule="scaled", model="NoiseModel", shape=(8, 4, 127, 129), dtype="bf16"),
    "unipc3_sde_flux_f32": dict(sampler="UniPC", kw={"order": 3, "stochasticity": 1}, schedule="flow", model="FlowModel", shape=(16, 16, 128, 128), dtype="f32"),
    "unipc3_sde_flux64_bf16": dict(sampler="UniPC", kw={"order": 3, "stochasticity": 1}, schedule="flow", model="FlowModel", shape=(64, 16, 128, 128), dtype="bf16"),
}
DEFAULT_WORKLOAD = "unipc3_sde_sdxl_bf16"
SWEEP = ["euler_sde_flow_f32_16", "euler_sde_flow_f32_64", "euler_sde_flow_f32_256", "euler_sde_flow_bf16_256", "adams9_sde_video_bf16", "adams9_sde_video_f32", "unipc3_sde_flux_bf16"]
TORCH_DTYPE = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16, "f64": torch.float64}


def numel_of(shape: tuple[int, ...]) -> int:
    n = 1
    for s in shape:
        n *= s
    return n


def measured_peak() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
# product arm


class Trajectory:
    """One latent batch walking a 25-step schedule with the analytic Gaussian denoiser's predictions pre-recorded."""

    def __init__(self, spec: dict, device: torch.device, seed: int, predictions: int | None = None, supplied_noise: bool = False) -> None:
        import cases
        from skrample_b200 import scheduling
        from skrample_b200.sampling import models, structured

        self.spec = spec
        self.sampler = cases.make_sampler(structured, models, {"sampler": spec["sampler"], "kw": spec["kw"]})
        self.schedule = cases.make_schedule(scheduling, spec["schedule"])
        self.model = cases.make_model(models, spec["model"])
        self.dtype = TORCH_DTYPE[spec["dtype"]]
        self.device = device
        g = torch.Generator(device=device).manual_seed(seed)
        shape = spec["shape"]
        self.points = self.schedule.schedule(STEPS_PER_TRAJECTORY)
        sigma_max = self.points[0].sigma
        self.x0 = (torch.randn(shape, device=device, generator=g) * sigma_max).to(self.dtype)
        count = predictions or STEPS_PER_TRAJECTORY
        self.count = count
        self.predictions: list[torch.Tensor] = []
        # "Random noise" of the workload: one CUDA generator per batch item (what the diffusers wrapper builds).
        # The draws are Philox keys; the step kernel generates the normals itself, nothing is written to memory.
        from skrample_b200.pytorch import noise as sk_noise

        self.noise_source = sk_noise.BatchTensorNoise.from_batch_inputs(
            sk_noise.Random,
            tuple(shape[1:]),
            [torch.Generator(device=device).manual_seed(seed * 1000 + i) for i in range(shape[0])],
            dtype=torch.float32,
        )
        self.noises = [self.noise_source.lazy(None) for _ in range(count)] if self.sampler.require_noise else [None] * count
        if supplied_noise and self.sampler.require_noise:
            # batches beyond the kernel's Philox table come back as tensors already; either way the workload's noise
            # is stored in the latent dtype, as a pipeline would hand it over
            self.noises = [(z.materialize() if hasattr(z, "materialize") else z).to(self.dtype) for z in self.noises]
        self.reset()

    def reset(self) -> None:
        self.x = self.x0
        self.n = 0
        self.previous: list = []

    def denoise(self, x: torch.Tensor, n: int) -> torch.Tensor:
        "Closed-form posterior mean for data ~ N(0, 1), expressed in the model's output space (excluded from timing)."
        from skrample_b200.sampling import models

        p = self.points[n]
        x32 = x.float()
        xhat = x32 * (p.alpha / (p.alpha * p.alpha + p.sigma * p.sigma))
        out = self.model.from_x(x32, xhat, p) if not isinstance(self.model, models.DataModel) else xhat
        return out.to(self.dtype)

    def record(self) -> None:
        "Walk the trajectory once, recording the network outputs the timed loop will replay."
        self.reset()
        self.predictions = []
        for n in range(STEPS_PER_TRAJECTORY):
            if n < self.count:
                self.predictions.append(self.denoise(self.x, n))
            self.step()
        self.reset()

    def step(self, prediction: torch.Tensor | None = None, noise: torch.Tensor | None = None) -> torch.Tensor:
        from skrample_b200.common import Step

        n = self.n
        res = self.sampler.sample(
            self.x,
            self.predictions[n % self.count] if prediction is None else prediction,
            Step.from_int(n, STEPS_PER_TRAJECTORY),
            self.model,
            self.schedule,
            (self.noises[n % self.count] if noise is None else noise) if self.sampler.require_noise else None,
            self.previous,
        )
        if self.sampler.require_previous:
            self.previous = (self.previous + [res])[-self.sampler.require_previous :]
        self.x = res.final
        self.n += 1
        if self.n == STEPS_PER_TRAJECTORY:
            self.reset()
        return res.final


SUPPLIED_NOISE = True


def step_bytes(traj_spec: dict, device: torch.device) -> list[int]:
    "Algorithmic bytes (sum of distinct tensor reads + writes) of each of the 25 steps, counted from the launches."
    from skrample_b200 import native

    t = Trajectory(traj_spec, device, seed=99, predictions=1, supplied_noise=SUPPLIED_NOISE)
    t.record()
    out: list[int] = []
    for _ in range(STEPS_PER_TRAJECTORY):
        native.ACCOUNT["bytes"] = 0
        native.ACCOUNT["on"] = True
        t.step()
        native.ACCOUNT["on"] = False
        out.append(native.ACCOUNT["bytes"])
    return out


class ClockSampler:
    "nvidia-smi clocks / throttle reasons while the timed region runs."

    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int) -> None:
        self.rows: list[list[str]] = []
        self.proc: subprocess.Popen | None = None
        self.index = index

    def __enter__(self) -> "ClockSampler":
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc: object) -> None:
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        clocks = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        maxes = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {
            "sm_mhz": clocks[len(clocks) // 2] if clocks else None,
            "sm_max_mhz": max(maxes) if maxes else None,
            "reasons": reasons,
            "samples": len(clocks),
        }


def graph_throughput(spec: dict, device: torch.device, steps: int, warmup: int, min_ring_bytes: int, streams: int = 1) -> dict:
    """Replay `steps` sampler launches from a CUDA graph; returns timing + byte accounting.  With `streams` > 1 the
    graph has that many parallel branches (latent batch r runs on branch r % streams): independent latent batches
    overlap the way concurrent requests would, which hides the launch / load / store phases of small steps."""
    from skrample_b200 import native

    per_step = step_bytes(spec, device)
    traj_bytes = sum(per_step)
    n = numel_of(spec["shape"])
    esize = TORCH_DTYPE[spec["dtype"]].itemsize
    # enough interleaved replicas that one round over them touches more than 2x L2
    replica_touch = max(per_step)
    replicas = max(2, min(64, -(-min_ring_bytes // replica_touch)))
    big = n * esize > 64 * 1024 * 1024
    keep = 2 if big else STEPS_PER_TRAJECTORY  # recorded predictions/noises per replica (memory bound for huge latents)
    trajs = [Trajectory(spec, device, seed=1234 + i, predictions=keep, supplied_noise=SUPPLIED_NOISE) for i in range(replicas)]
    for t in trajs:
        t.record()

    streams = max(1, min(streams, replicas))
    rounds = replicas * STEPS_PER_TRAJECTORY  # every replica walks exactly one trajectory per graph replay
    stream = torch.cuda.Stream(device=device)
    branches = [torch.cuda.Stream(device=device) for _ in range(streams)] if streams > 1 else []

    def run(count: int) -> None:
        if not branches:
            for k in range(count):
                trajs[k % replicas].step()
            return
        for branch in branches:  # fork
            branch.wait_stream(stream)
        for b, branch in enumerate(branches):
            with torch.cuda.stream(branch):
                mine = list(range(b, replicas, streams))
                for k in range(count // replicas):
                    for r in mine:
                        trajs[r].step()
        for branch in branches:  # join
            stream.wait_stream(branch)

    with torch.cuda.stream(stream):
        run(rounds)  # eager warm-up: caches, allocator
        torch.cuda.synchronize(device)
        before = native.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            run(rounds)
        launches_per_replay = native.launch_count() - before
    torch.cuda.synchronize(device)

    replays = max(1, -(-steps // rounds))
    warm = max(1, -(-warmup // rounds))
    for _ in range(warm):
        graph.replay()
    torch.cuda.synchronize(device)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index or 0) as clocks:
        start.record()
        for _ in range(replays):
            graph.replay()
        stop.record()
        torch.cuda.synchronize(device)
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    total_steps = replays * rounds
    total_bytes = replays * replicas * traj_bytes
    return {
        "elapsed_ms": elapsed_ms,
        "steps": total_steps,
        "bytes": total_bytes,
        "launches": replays * launches_per_replay,
        "replicas": replicas,
        "bytes_per_step_avg": traj_bytes / STEPS_PER_TRAJECTORY,
        "bytes_per_step_max": max(per_step),
        "clocks": clocks.summary(),
        "batch": spec["shape"][0],
    }


# BASELINE.json configs[2]: an explicit Runge-Kutta step (one launch per stage + the final update)
RK_SWEEP = {
    "rkultra4_flux_bf16": dict(order=4, shape=(16, 16, 128, 128), dtype="bf16"),
    "rkultra4_flux_f32": dict(order=4, shape=(16, 16, 128, 128), dtype="f32"),
}


def rk_step_throughput(spec: dict, device: torch.device, replicas: int = 16, reps: int = 20) -> dict:
    """RKUltra(order) / FlowShift(Linear(), 3) / FlowModel over `replicas` interleaved latents (inputs come from HBM).
    The network is a table of pre-recorded outputs, so the captured graph holds only this library's launches."""
    from skrample_b200 import native, scheduling
    from skrample_b200.common import Step
    from skrample_b200.sampling import functional, models

    dtype = TORCH_DTYPE[spec["dtype"]]
    sampler = functional.RKUltra(order=spec["order"])
    schedule, model_transform = scheduling.FlowShift(scheduling.Linear(), shift=3.0), models.FlowModel()
    g = torch.Generator(device=device).manual_seed(0)
    latents = [torch.randn(spec["shape"], device=device, generator=g).to(dtype) for _ in range(replicas)]
    outputs = [[torch.randn(spec["shape"], device=device, generator=g).to(dtype) * 0.3 for _ in range(8)] for _ in range(replicas)]
    interior = (3, 12)  # two interior steps of a 25-step schedule per replica

    def run() -> None:
        for r in range(replicas):
            calls, table, x = iter(range(1 << 30)), outputs[r], latents[r]
            for n in interior:
                x = sampler.step(x, lambda s, t, sigma, alpha: table[next(calls) % 8], model_transform, schedule, Step.from_int(n, STEPS_PER_TRAJECTORY))

    stream = torch.cuda.Stream(device=device)
    with torch.cuda.stream(stream):
        run()
        torch.cuda.synchronize(device)
        before = native.launch_count()
        native.ACCOUNT["bytes"] = 0
        native.ACCOUNT["on"] = True
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            run()
        native.ACCOUNT["on"] = False
        launches = native.launch_count() - before
    torch.cuda.synchronize(device)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize(device)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        graph.replay()
    stop.record()
    torch.cuda.synchronize(device)
    ms = start.elapsed_time(stop) / reps
    rk_steps = replicas * len(interior)
    return {"us_per_rk_step": ms * 1e3 / rk_steps, "launches_per_step": launches / rk_steps, "GBps": native.ACCOUNT["bytes"] / (ms * 1e-3) / 1e9, "bytes_per_step": native.ACCOUNT["bytes"] / rk_steps}


def noise_generator_times(device: torch.device, unit: tuple[int, ...] = (16, 21, 90, 160), reps: int = 30) -> list[dict]:
    """Device time per draw of each noise generator on one BASELINE.json configs[3] unit (a 16x21x90x160 video latent,
    fp32 generator dtype): CUDA events around `generate`, after warm-up (cuFFT plans, allocator)."""
    from skrample_b200.common import Step
    from skrample_b200.pytorch import noise

    rows = []
    step = Step.from_int(5, STEPS_PER_TRAJECTORY)
    for name, props in (("Random", None), ("Offset", noise.OffsetProps()), ("Pyramid", noise.PyramidProps()), ("Colored", noise.ColoredProps()), ("Brownian", noise.BrownianProps())):
        cls = getattr(noise, name)
        generator = torch.Generator(device=device).manual_seed(1)
        source = cls.from_inputs(unit, generator, dtype=torch.float32) if props is None else cls.from_inputs(unit, generator, props, dtype=torch.float32)
        for _ in range(20):
            out = source.generate(step)
        torch.cuda.synchronize(device)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(reps):
            out = source.generate(step)
        stop.record()
        torch.cuda.synchronize(device)
        us = start.elapsed_time(stop) / reps * 1e3
        rows.append({"generator": name, "unit_shape": list(unit), "us_per_draw": us, "GBps_written": out.numel() * out.element_size() / us / 1e3})
    return rows


def e2e_throughput(spec: dict, device: torch.device, steps: int, warmup: int, inflight: int = 1) -> dict:
    """Public API with host buffers: H2D of the step's prediction, noise drawn on the device, sampler.sample, D2H of the
    result - every step.  ``inflight`` independent latent batches (requests) are advanced round robin: a request's next
    step starts only after its previous result has arrived in host memory, but while that copy is in flight the host
    prepares and launches the other requests' steps.  ``inflight=1`` is the plain synchronous loop."""
    trajs = [Trajectory(spec, device, seed=4321 + 17 * i) for i in range(inflight)]
    for traj in trajs:
        traj.record()
    per_step = step_bytes(spec, device)
    host_pred = [[p.cpu().pin_memory() for p in traj.predictions] for traj in trajs]
    result_host = [torch.empty(spec["shape"], dtype=trajs[0].dtype).pin_memory() for _ in trajs]
    arrived = [torch.cuda.Event() for _ in trajs]
    pending = [False] * inflight

    def one(k: int) -> None:
        slot = k % inflight
        traj = trajs[slot]
        if pending[slot]:
            arrived[slot].synchronize()  # the caller consumes this request's previous result before its next step
        pred = host_pred[slot][traj.n].to(device, non_blocking=True)
        noise = None
        if traj.sampler.require_noise:  # fresh noise every step, generated on the device (fill kernel or in-step draw)
            noise = traj.noise_source.auto(None) if SUPPLIED_NOISE else traj.noise_source.lazy(None)
        final = traj.step(pred, noise)
        result_host[slot].copy_(final, non_blocking=True)
        if inflight == 1:
            torch.cuda.current_stream().synchronize()
        else:
            arrived[slot].record()
            pending[slot] = True

    def drain() -> None:
        torch.cuda.synchronize(device)
        for slot in range(inflight):
            pending[slot] = False

    for k in range(warmup):
        one(k)
    drain()
    for traj in trajs:
        traj.reset()
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        one(k)
    torch.cuda.synchronize(device)
    elapsed = time.perf_counter() - t0
    barrier()
    n = numel_of(spec["shape"])
    esize = trajs[0].dtype.itemsize
    total_bytes = sum(per_step[(k // inflight) % STEPS_PER_TRAJECTORY] for k in range(steps))
    return {
        "elapsed_s": elapsed,
        "bytes": total_bytes,
        "h2d": n * esize,
        "d2h": n * esize,
    }


def e2e_graphed_throughput(spec: dict, device: torch.device, steps: int, warmup: int) -> dict:
    """The e2e step through skrample_b200.graphs.GraphedTrajectory: the same host buffers and copies, the sampler
    launches replayed from CUDA graphs captured once."""
    from skrample_b200.graphs import GraphedTrajectory

    traj = Trajectory(spec, device, seed=4321)
    traj.record()
    host_pred = [p.cpu().pin_memory() for p in traj.predictions]
    result_host = torch.empty(spec["shape"], dtype=traj.dtype).pin_memory()
    graphed = GraphedTrajectory(traj.sampler, traj.model, traj.schedule, STEPS_PER_TRAJECTORY, like=traj.x0)

    def one() -> None:
        if graphed.position == len(graphed):
            graphed.start(traj.x0)
        graphed.prediction().copy_(host_pred[graphed.position], non_blocking=True)
        if traj.sampler.require_noise:
            traj.noise_source.generate_into(graphed.noise(), None)  # fresh noise every step, written in place
        final = graphed.step()
        result_host.copy_(final, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    graphed.start(traj.x0)
    for _ in range(warmup):
        one()
    graphed.start(traj.x0)
    torch.cuda.synchronize(device)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize(device)
    elapsed = time.perf_counter() - t0
    barrier()
    return {"elapsed_s": elapsed}


def measured_traffic() -> float | None:
    """DRAM bytes per launch of the dominant step kernel on the default workload, from the latest ncu launch list
    summarised under profiles/ (tools/profile_round.sh + tools/summarize_profiles.py); None when not captured."""
    if "--workload" in sys.argv:
        return None
    found = sorted((Path(__file__).resolve().parent / "profiles").glob("r*_traffic.json"))
    if not found:
        return None
    try:
        return float(json.loads(found[-1].read_text())["dram_bytes_per_launch"])
    except (OSError, ValueError, KeyError):
        return None


def barrier() -> None:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()


def max_over_ranks(value: float, device: torch.device) -> float:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        t = torch.tensor([value], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())
    return value


# --------------------------------------------------------------------------------------------------------------
# CPU oracle arm (reference algorithm restated, torch-CPU tensors so every host thread is used)


def cpu_oracle_steps(spec: dict, steps: int, warmup: int, budget_s: float) -> dict:
    import cases
    import oracle_run
    from oracle import skrample_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    case = {"sampler": spec["sampler"], "kw": spec["kw"], "schedule": spec["schedule"], "model": spec["model"]}
    model = oracle_run.MODELS[spec["model"]]
    sch = oracle_run.schedule(spec["schedule"])
    need_noise, need_prev = oracle_run.require(case)
    shape = spec["shape"]
    g = torch.Generator().manual_seed(7)
    pts = sch.schedule(STEPS_PER_TRAJECTORY)
    x0 = torch.randn(shape, generator=g) * pts[0].sigma
    noises = [torch.randn(shape, generator=g) for _ in range(STEPS_PER_TRAJECTORY)]
    per_step_bytes = reference_step_bytes(spec)

    state = {"x": x0, "n": 0, "prev": []}

    def one() -> float:
        n = state["n"]
        p = pts[n]
        x = state["x"]
        xhat = x * (p.alpha / (p.alpha * p.alpha + p.sigma * p.sigma))
        out = model.from_x(x, xhat, p) if model.kind != "data" else xhat  # analytic denoiser, excluded from timing
        cur = O.Rec(x, out, O.St.from_int(n, STEPS_PER_TRAJECTORY), noises[n] if need_noise else None)
        t0 = time.perf_counter()
        rec = oracle_run.one_step(case, cur, state["prev"], model, sch)
        dt = time.perf_counter() - t0
        state["prev"] = (state["prev"] + [rec])[-need_prev:] if need_prev else []
        state["x"] = rec.final
        state["n"] = (n + 1) % STEPS_PER_TRAJECTORY
        if state["n"] == 0:
            state["x"], state["prev"] = x0, []
        return dt

    for _ in range(warmup):
        one()
    state.update(x=x0, n=0, prev=[])
    spent = 0.0
    done = 0
    total_bytes = 0
    wall0 = time.perf_counter()
    while done < steps and (time.perf_counter() - wall0) < budget_s:
        total_bytes += per_step_bytes[state["n"]]
        spent += one()
        done += 1
    return {"seconds": spent, "steps": done, "bytes": total_bytes, "threads": threads}


def reference_step_bytes(spec: dict) -> list[int]:
    """Algorithmic bytes per step of the workload, by the same rule as the GPU arm (distinct N-sized tensors read +
    written by a minimal-traffic step: history order + sample + prediction [+ noise] reads, results written)."""
    n = numel_of(spec["shape"])
    esize = 4  # the CPU arm computes and stores fp32
    name, kw = spec["sampler"], spec["kw"]
    order = kw.get("order", 2 if name != "Euler" else 1)
    cap = {"Euler": 1, "DPM": 3, "Adams": 9, "UniP": 9, "UniPC": 9}.get(name, 1)
    order = max(1, min(order, cap))
    noise = 1 if abs(kw.get("stochasticity", 0)) > 1e-8 else 0
    convert = 0 if spec["model"] == "DataModel" else 1
    out: list[int] = []
    for i in range(STEPS_PER_TRAJECTORY):
        k = max(1, min(order, i + 1, STEPS_PER_TRAJECTORY - i))
        if name == "UniPC":
            kc = max(1, min(order, i, STEPS_PER_TRAJECTORY - (i - 1))) if i > 0 else 0
            reads = 2 + noise + (k - 1 if i == 0 else max(k - 1, kc) + 1 + noise)
            writes = 2 + (1 if i > 0 else 0)
        else:
            reads = 2 + noise + (k - 1)
            writes = 1 + (1 if convert and cap > 1 and order > 1 else 0)
        out.append((reads + writes) * n * esize)
    return out


# --------------------------------------------------------------------------------------------------------------


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--sweep", action="store_true", help="also time the larger BASELINE shapes (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4, help="parallel graph branches the latent batches of `value` run on (1: a single chain)")
    ap.add_argument("--inflight", type=int, default=2, help="independent requests advanced round robin by the e2e leg (1: the synchronous loop only)")
    ap.add_argument("--fused-noise", action="store_true", help="draw the noise inside the step kernel (PhiloxDraw) instead of reading the tensor skr_noise_fill wrote")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: anything a library prints there while the bench runs (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for one) goes to stderr instead
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict) -> None:
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    global SUPPLIED_NOISE
    SUPPLIED_NOISE = not args.fused_noise

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    spec = WORKLOADS[args.workload]
    n = numel_of(spec["shape"])
    config = {
        "workload": f"{spec['sampler']}({', '.join(f'{k}={v}' for k, v in spec['kw'].items())}) {spec['schedule']} {spec['model']} latent {'x'.join(map(str, spec['shape']))} {spec['dtype']} storage / fp32 compute, {STEPS_PER_TRAJECTORY}-step trajectories, Random noise (Philox, one generator per item; {'drawn inside the step kernel' if args.fused_noise else 'written by skr_noise_fill before the timed region, like the CPU arm'}), analytic Gaussian denoiser (pre-recorded)",
        "name": args.workload,
        "per_gpu_batch": spec["shape"][0],
        "global_batch": spec["shape"][0] * world,
        "parallelism": f"batch-sharded x{world}, no collective on the step path",
    }

    if args.impl == "reference":
        if rank != 0:
            return
        budget = 120.0
        res = cpu_oracle_steps(spec, args.steps, min(args.warmup, 25), budget)
        gbs = res["bytes"] / res["seconds"] / 1e9
        lsps = res["steps"] * spec["shape"][0] / res["seconds"]
        line = {
            "impl": "reference",
            "metric": "sampler latent-steps/s (one batch item advanced one solver step)",
            "value": lsps,
            "unit": "latent-steps/s",
            "n_gpus": args.gpus,
            "steps": res["steps"],
            "warmup": min(args.warmup, 25),
            "ms_per_step": res["seconds"] / res["steps"] * 1e3,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": config,
            "sampler_step_GBps": gbs,
            "cpu_baseline": {
                "value": lsps,
                "unit": "latent-steps/s",
                "cores": res["threads"],
                "kind": "port",
                "sample": f"{res['steps']} sampler steps of the workload in fp32 on torch-CPU tensors (oracle/skrample_oracle.py), sampler time only",
            },
            "e2e": {"value": lsps, "unit": "latent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for --impl ours (there is no CPU fallback)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)

    from skrample_b200 import native

    native.load()
    peak, peak_src = measured_peak()

    launches_before = native.launch_count()
    dev = graph_throughput(spec, device, args.steps, args.warmup, 2 * L2_BYTES)
    timed_launches = dev["launches"]
    elapsed_ms = max_over_ranks(dev["elapsed_ms"], device)
    gbs = dev["bytes"] * world / (elapsed_ms / 1e3) / 1e9
    latent_steps = dev["steps"] * spec["shape"][0] * world / (elapsed_ms / 1e3)
    ms_per_step = elapsed_ms / dev["steps"]
    per_launch_us = elapsed_ms * 1e3 / dev["launches"]
    bytes_per_launch = dev["bytes"] / dev["launches"]
    achieved = bytes_per_launch / (per_launch_us * 1e-6) / 1e9

    # the same launches as one chain on one stream: a launch's duration is well defined there, so the roofline of the
    # step kernel is taken from this run
    one_stream = {
        "value": latent_steps,
        "unit": "latent-steps/s",
        "ms_per_step": ms_per_step,
        "sampler_step_GBps": gbs,
        "frac_of_measured_peak": gbs / world / peak,
        "gpu_launches": timed_launches,
    }
    branches = 1
    if args.streams > 1:
        # whole-job throughput: the interleaved latent batches are independent requests, so the graph runs them on
        # parallel branches and the launch / load / store phases of different batches overlap
        torch.cuda.empty_cache()
        many = graph_throughput(spec, device, args.steps, args.warmup, 2 * L2_BYTES, streams=args.streams)
        many_ms = max_over_ranks(many["elapsed_ms"], device)
        branches = min(args.streams, many["replicas"])
        dev = many
        timed_launches = many["launches"]
        gbs = many["bytes"] * world / (many_ms / 1e3) / 1e9
        latent_steps = many["steps"] * spec["shape"][0] * world / (many_ms / 1e3)
        ms_per_step = many_ms / many["steps"]
        torch.cuda.empty_cache()

    e2e_steps = min(args.steps, 500)
    single = e2e_throughput(spec, device, e2e_steps, min(args.warmup, 50))
    single_elapsed = max_over_ranks(single["elapsed_s"], device)
    e2e, e2e_elapsed = single, single_elapsed
    if args.inflight > 1:
        e2e = e2e_throughput(spec, device, e2e_steps, min(args.warmup, 50), inflight=args.inflight)
        e2e_elapsed = max_over_ranks(e2e["elapsed_s"], device)
    e2e_gbs = e2e["bytes"] * world / e2e_elapsed / 1e9
    graphed_elapsed = max_over_ranks(e2e_graphed_throughput(spec, device, e2e_steps, min(args.warmup, 50))["elapsed_s"], device)

    line = {
        "metric": "sampler latent-steps/s (one batch item advanced one solver step)",
        "value": latent_steps,
        "unit": "latent-steps/s",
        "n_gpus": world,
        "steps": dev["steps"],
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": config
        | {
            "launch": f"CUDA graph replay of the sampler launches, latent batches on {branches} parallel graph branch(es)",
            "graph_branches": branches,
            "l2": f"{dev['replicas']} interleaved latent batches, working set per round > 2x L2 (inputs come from HBM)",
            "storage_dtype": spec["dtype"],
        },
        "sampler_step_GBps": gbs,
        "pct_of_hbm_peak": {"measured": gbs / world / peak, "nominal_8TBs": gbs / world / 8000.0},
        "gpu_launches": timed_launches,
        "clocks": dev["clocks"],
        "one_stream": one_stream,
        "e2e": {
            "value": e2e_steps * spec["shape"][0] * world / e2e_elapsed,
            "unit": "latent-steps/s",
            "h2d_bytes_per_step": e2e["h2d"],
            "d2h_bytes_per_step": e2e["d2h"],
            "sampler_step_GBps": e2e_gbs,
            "ms_per_step": e2e_elapsed / e2e_steps * 1e3,
            "steps": e2e_steps,
            "api": "structured sampler .sample() per step (the reference's call), pinned-host prediction in, result out; noise from BatchTensorNoise.auto (in-kernel Philox draw at this size)",
            "requests_in_flight": max(1, args.inflight),
            "one_request": {
                "value": e2e_steps * spec["shape"][0] * world / single_elapsed,
                "ms_per_step": single_elapsed / e2e_steps * 1e3,
                "note": "the plain synchronous loop: every step waits for its result on the host before the next begins",
            },
        },
        "e2e_graphed": {
            "value": e2e_steps * spec["shape"][0] * world / graphed_elapsed,
            "unit": "latent-steps/s",
            "ms_per_step": graphed_elapsed / e2e_steps * 1e3,
            "steps": e2e_steps,
            "api": "skrample_b200.graphs.GraphedTrajectory.step(): same host buffers and copies, launches replayed from CUDA graphs",
        },
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": measured_traffic(),
            "kernel": "skr::block_kernel",
            "bytes_per_launch": bytes_per_launch,
            "us_per_launch": per_launch_us,
            "measured_on": "the one-stream chain of the same launches (`one_stream`), CUDA events around the timed replays",
            "peak_source": peak_src,
        },
    }
    assert native.launch_count() > launches_before

    if rank == 0 and world == 1 and args.sweep:
        sweep = []
        for name in SWEEP:
            s = WORKLOADS[name]
            torch.cuda.empty_cache()
            r = graph_throughput(s, device, STEPS_PER_TRAJECTORY * 8, STEPS_PER_TRAJECTORY * 2, 2 * L2_BYTES)
            us = r["elapsed_ms"] * 1e3 / r["launches"]
            ach = r["bytes"] / r["launches"] / (us * 1e-6) / 1e9
            sweep.append({"workload": name, "shape": list(s["shape"]), "dtype": s["dtype"], "us_per_step": us, "GBps": ach, "frac_of_measured_peak": ach / peak, "latent_steps_per_s": s["shape"][0] / (us * 1e-6), "bytes_per_step_avg": r["bytes_per_step_avg"]})
        for name, s in RK_SWEEP.items():
            torch.cuda.empty_cache()
            r = rk_step_throughput(s, device)
            sweep.append({"workload": name, "shape": list(s["shape"]), "dtype": s["dtype"], "us_per_step": r["us_per_rk_step"], "launches_per_step": r["launches_per_step"], "GBps": r["GBps"], "frac_of_measured_peak": r["GBps"] / peak, "latent_steps_per_s": s["shape"][0] / (r["us_per_rk_step"] * 1e-6), "bytes_per_step_avg": r["bytes_per_step"]})
        line["sweep"] = sweep
        line["noise_generators"] = noise_generator_times(device)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_oracle_steps(spec, 10_000, 25, 15.0)
        line["cpu_baseline"] = {
            "value": res["steps"] * spec["shape"][0] / res["seconds"],
            "unit": "latent-steps/s",
            "cores": res["threads"],
            "kind": "port",
            "sample": f"{res['steps']} sampler steps of the workload in fp32 on torch-CPU tensors (oracle/skrample_oracle.py), sampler time only",
            "sampler_step_GBps": res["bytes"] / res["seconds"] / 1e9,
            "ms_per_step": res["seconds"] / res["steps"] * 1e3,
        }

    if rank == 0:
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
