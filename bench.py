#!/usr/bin/env python
"""Benchmark of the skrample_b200 sampler step (BASELINE.json metric: sampler-step GB/s and latent-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--quick]

A "step" is ONE solver step of the workload's sampler over one latent batch.  The default workload is BASELINE.json
configs[1]: UniPC order 3, stochastic, Random noise (device Philox, one generator per batch item), Scaled schedule,
epsilon model, SDXL latent 8x4x128x128, bf16 storage / fp32 compute, on one B200.

  value        latent-steps/s (batch items x solver steps per second) of ONE dependent chain of launches on ONE stream,
               the step's noise drawn inside the timed region (in the step kernel itself where the library chooses
               that, else by its fill kernel), inputs resident in HBM, launches replayed from a CUDA graph, timed with
               CUDA events; the median of >= 5 timed blocks of >= 0.1 s each (at least --steps steps in total), max over
               ranks.  Trajectories of several latent batches are interleaved so consecutive launches never touch
               the same buffers and the working set of a round (> 2x L2) comes from HBM.  ``ms_per_step`` is a real
               per-step time: the launches are serialised.
  roofline     the dominant kernel of that timed region - the step kernel that draws its own noise: algorithmic bytes per
               launch (no noise tensor) / average launch duration vs the measured HBM copy peak.
  kernel_only  the same chain with the noise written before the timed region (the step reads noise tensors: more bytes,
               less arithmetic); ``roofline_supplied_noise`` is taken here (round 1's definition).
  concurrent_requests  the same launches on ``--streams`` parallel graph branches (independent latent batches side by
               side): a multi-request THROUGHPUT, not a step latency.
  e2e          the same steps through the public API (``sampler.sample``) with HOST buffers: per step the model
               prediction is copied from pinned host memory, the noise is drawn on the device and the result is read
               back.  ``--inflight`` independent requests (each on its own stream; default: the best of 1 / 2 / 4 by
               a short trial, recorded in the line) are advanced round robin, each waiting for its own previous result; ``e2e.one_request`` is the plain synchronous loop;
               ``e2e_graphed`` replays the steps through ``GraphedTrajectory``.  At least 2000 steps and 0.5 s each.
               The floor of this leg is PCIe: a 1 MiB copy in and a 1 MiB copy out per step take 34 us when both
               directions are busy (24 / 22 us alone).
  rows         the other BASELINE.json shapes (the Euler sweep of configs[4], the video shard of configs[3] with Pyramid
               and Colored noise drawn per step, UniPC on a Flux-sized latent, the RKUltra(4) step of configs[2]), each
               with its own roofline fraction; ``noise_generators`` times one draw of every generator.
  strong_scaling  configs[3] (global batch 8) and the top of configs[4] (global batch 256) split over the N ranks.
  cpu_baseline the reference's own CPU implementation (the unmodified package under baseline/_ref when it is there, else
               the oracle port) on all host threads, a bounded sample of the same workload.

Under torchrun every rank runs the same per-GPU batch on its own GPU (weak scaling, no collective on the step path);
rank 0 prints one JSON line.
"""

from __future__ import annotations

import argparse
import json
import math
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
METRIC = "sampler latent-steps/s (one batch item advanced one solver step)"


def _w(sampler: str, kw: dict, schedule: str, model: str, shape: tuple, dtype: str, noise: str = "Random") -> dict:
    return dict(sampler=sampler, kw=kw, schedule=schedule, model=model, shape=shape, dtype=dtype, noise=noise)


_EULER = ("Euler", {"stochasticity": 1}, "flow", "FlowModel")
_ADAMS9 = ("Adams", {"order": 9, "stochasticity": 1}, "flow", "FlowModel")
_UNIPC3 = ("UniPC", {"order": 3, "stochasticity": 1})
_VIDEO = (1, 16, 21, 90, 160)
WORKLOADS = {
    # BASELINE.json configs[1]
    "unipc3_sde_sdxl_bf16": _w(*_UNIPC3, "scaled", "NoiseModel", (8, 4, 128, 128), "bf16"),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "dpm2_scaled_fp32": _w("DPM", {"order": 2}, "scaled", "NoiseModel", (1, 4, 128, 128), "f32"),
    # BASELINE.json configs[3], one GPU's shard (one item of 8x16x21x90x160), noise generator as named
    "adams9_sde_video_bf16": _w(*_ADAMS9, _VIDEO, "bf16"),
    "adams9_sde_video_f32": _w(*_ADAMS9, _VIDEO, "f32"),
    "adams9_sde_video_bf16_pyramid": _w(*_ADAMS9, _VIDEO, "bf16", "Pyramid"),
    "adams9_sde_video_bf16_colored": _w(*_ADAMS9, _VIDEO, "bf16", "Colored"),
    "adams9_sde_video8_bf16_pyramid": _w(*_ADAMS9, (8, *_VIDEO[1:]), "bf16", "Pyramid"),
    "adams9_sde_video8_bf16_colored": _w(*_ADAMS9, (8, *_VIDEO[1:]), "bf16", "Colored"),
    # BASELINE.json configs[4]: the Euler sweep
    "euler_sde_flow_f32_1x4x64": _w(*_EULER, (1, 4, 64, 64), "f32"),
    "euler_sde_flow_f32_1x4x128": _w(*_EULER, (1, 4, 128, 128), "f32"),
    "euler_sde_flow_f32_4": _w(*_EULER, (4, 16, 128, 128), "f32"),
    "euler_sde_flow_f32_16": _w(*_EULER, (16, 16, 128, 128), "f32"),
    "euler_sde_flow_f32_64": _w(*_EULER, (64, 16, 128, 128), "f32"),
    "euler_sde_flow_f32_256": _w(*_EULER, (256, 16, 128, 128), "f32"),
    "euler_sde_flow_bf16_256": _w(*_EULER, (256, 16, 128, 128), "bf16"),
    "unipc3_sde_flux_bf16": _w(*_UNIPC3, "flow", "FlowModel", (16, 16, 128, 128), "bf16"),
    "unipc3_sde_ragged_bf16": _w(*_UNIPC3, "scaled", "NoiseModel", (8, 4, 127, 129), "bf16"),
    "unipc3_sde_flux_f32": _w(*_UNIPC3, "flow", "FlowModel", (16, 16, 128, 128), "f32"),
    "unipc3_sde_flux64_bf16": _w(*_UNIPC3, "flow", "FlowModel", (64, 16, 128, 128), "bf16"),
}
DEFAULT_WORKLOAD = "unipc3_sde_sdxl_bf16"
ROWS = [
    "euler_sde_flow_f32_1x4x64",
    "euler_sde_flow_f32_1x4x128",
    "euler_sde_flow_f32_4",
    "euler_sde_flow_f32_16",
    "euler_sde_flow_f32_64",
    "euler_sde_flow_f32_256",
    "euler_sde_flow_bf16_256",
    "adams9_sde_video_bf16",
    "adams9_sde_video_f32",
    "adams9_sde_video_bf16_pyramid",
    "adams9_sde_video_bf16_colored",
    "unipc3_sde_flux_bf16",
]
CONTRACTED_ROWS = ["unipc3_sde_flux_bf16", "unipc3_sde_sdxl_bf16"]
# global workloads of the strong-scaling legs: (name, global batch, workload with the per-item shape)
STRONG = [
    ("configs[3] video latent 8x16x21x90x160 bf16, Adams-9 SDE, Pyramid noise", 8, "adams9_sde_video_bf16_pyramid"),
    ("configs[3] video latent 8x16x21x90x160 bf16, Adams-9 SDE, Colored noise", 8, "adams9_sde_video_bf16_colored"),
    ("configs[4] Euler SDE 256x16x128x128 fp32, Random noise", 256, "euler_sde_flow_f32_256"),
]
TORCH_DTYPE = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16, "f64": torch.float64}


def numel_of(shape: tuple[int, ...]) -> int:
    return math.prod(shape)


def measured_peak() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_text(spec: dict) -> str:
    kw = ", ".join(f"{k}={v}" for k, v in spec["kw"].items())
    return (
        f"{spec['sampler']}({kw}) {spec['schedule']} {spec['model']} latent {'x'.join(map(str, spec['shape']))} {spec['dtype']} storage / fp32 compute, "
        f"{STEPS_PER_TRAJECTORY}-step trajectories, {spec['noise']} noise (one generator per batch item, a fresh draw every step), "
        "analytic Gaussian denoiser (pre-recorded, excluded from the timing)"
    )


def config_of(spec: dict, name: str, world: int) -> dict:
    "The same dictionary on both arms (the driver compares them); arm-specific detail goes under `launch`."
    return {
        "workload": workload_text(spec),
        "name": name,
        "per_gpu_batch": spec["shape"][0],
        "global_batch": spec["shape"][0] * world,
        "parallelism": f"batch-sharded x{world}, no collective on the step path",
    }


# --------------------------------------------------------------------------------------------------------------
# product arm


class Trajectory:
    """One latent batch walking a 25-step schedule with the analytic Gaussian denoiser's predictions pre-recorded.

    ``noise``: "auto" - a fresh draw every step from the workload's generator, in whichever form the library picks
    (Philox keys consumed inside the step kernel, or a tensor written by the generator's kernels); "supplied" - the
    draws are written once, before any timed region (what the CPU arm gets)."""

    def __init__(self, spec: dict, device: torch.device, seed: int, keep: int | None = None, noise: str = "auto", first_item: int = 0) -> None:
        import cases
        from skrample_b200 import scheduling
        from skrample_b200.common import Step
        from skrample_b200.pytorch import noise as sk_noise
        from skrample_b200.sampling import models, structured

        self.spec = spec
        self.sampler = cases.make_sampler(structured, models, {"sampler": spec["sampler"], "kw": spec["kw"]})
        self.schedule = cases.make_schedule(scheduling, spec["schedule"])
        self.model = cases.make_model(models, spec["model"])
        self.dtype = TORCH_DTYPE[spec["dtype"]]
        self.device = device
        self.noise_mode = noise if self.sampler.require_noise else "none"
        g = torch.Generator(device=device).manual_seed(seed)
        shape = spec["shape"]
        self.points = self.schedule.schedule(STEPS_PER_TRAJECTORY)
        self.x0 = (torch.randn(shape, device=device, generator=g) * self.points[0].sigma).to(self.dtype)
        self.count = keep or STEPS_PER_TRAJECTORY
        self.predictions: list[torch.Tensor] = []
        # One CUDA generator per batch item, keyed by the item's GLOBAL index (what the diffusers wrapper builds): any
        # sharding of the batch draws the same noise for an item.  Random draws are fp32 normals (consumed as such by the
        # step kernel); the composite generators write the latent's storage type directly.
        kind = getattr(sk_noise, spec["noise"])
        self.noise_source = sk_noise.BatchTensorNoise.from_batch_inputs(
            kind,
            tuple(shape[1:]),
            [torch.Generator(device=device).manual_seed(seed * 1000 + first_item + i) for i in range(shape[0])],
            dtype=torch.float32 if kind is sk_noise.Random else self.dtype,
        )
        self.steps = [Step.from_int(n, STEPS_PER_TRAJECTORY) for n in range(STEPS_PER_TRAJECTORY)]
        self.keep = self.sampler.require_previous
        self.noises: list = []
        if self.noise_mode == "supplied":
            self.noises = [self.draw(n % STEPS_PER_TRAJECTORY, materialise=True) for n in range(self.count)]
        self.reset()

    def draw(self, n: int, materialise: bool = False):  # noqa: ANN201
        z = self.noise_source.auto(self.steps[n])
        if materialise:
            z = (z.materialize() if hasattr(z, "materialize") else z).to(self.dtype)
        return z

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
        n = self.n
        if noise is None and self.noise_mode != "none":
            noise = self.noises[n % self.count] if self.noise_mode == "supplied" else self.draw(n)
        res = self.sampler.sample(
            self.x,
            self.predictions[n % self.count] if prediction is None else prediction,
            self.steps[n],
            self.model,
            self.schedule,
            noise,
            self.previous,
        )
        keep = self.keep
        self.previous = (self.previous + [res])[-keep:] if keep else []
        self.x = res.final
        self.n += 1
        if self.n == STEPS_PER_TRAJECTORY:
            self.reset()
        return res.final


def step_bytes(spec: dict, device: torch.device, noise: str) -> list[int]:
    "Algorithmic bytes (sum of distinct tensor reads + writes) of each of the 25 steps, counted from the launches."
    from skrample_b200 import native

    t = Trajectory(spec, device, seed=99, keep=1, noise=noise)
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


class NullClocks:
    def __enter__(self) -> "NullClocks":
        return self

    def __exit__(self, *exc: object) -> None:
        pass

    def summary(self) -> dict:
        return {}


def chain_time(
    spec: dict,
    device: torch.device,
    noise: str,
    requested_steps: int,
    warmup: int,
    streams: int = 1,
    min_seconds: float = 0.5,
    blocks: int = 5,
    clocks: bool = False,
    min_ring_bytes: int = 2 * L2_BYTES,
    first_item: int = 0,
) -> dict:
    """Replay the sampler launches of `replicas` interleaved 25-step trajectories from a CUDA graph.  One stream: the
    launches form one dependent chain.  ``streams`` > 1: latent batch r runs on graph branch r % streams (independent
    requests side by side).  Timed as ``blocks`` blocks of whole replays (CUDA events on the replaying stream); the
    reported time per step is the MEDIAN block, max over ranks."""
    from skrample_b200 import native

    per_step = step_bytes(spec, device, noise)
    traj_bytes = sum(per_step)
    n = numel_of(spec["shape"])
    esize = TORCH_DTYPE[spec["dtype"]].itemsize
    replicas = max(2, min(64, -(-min_ring_bytes // max(per_step))))  # one round over them touches more than 2x L2
    keep = 2 if n * esize > 64 * 1024 * 1024 else STEPS_PER_TRAJECTORY  # recorded predictions / noises per replica
    trajs = [Trajectory(spec, device, seed=1234 + i, keep=keep, noise=noise, first_item=first_item) for i in range(replicas)]
    for t in trajs:
        t.record()

    streams = max(1, min(streams, replicas))
    rounds = replicas * STEPS_PER_TRAJECTORY  # every replica walks exactly one trajectory per graph replay
    stream = torch.cuda.Stream(device=device)
    branches = [torch.cuda.Stream(device=device) for _ in range(streams)] if streams > 1 else []

    def run() -> None:
        if not branches:
            for k in range(rounds):
                trajs[k % replicas].step()
            return
        for branch in branches:  # fork
            branch.wait_stream(stream)
        for b, branch in enumerate(branches):
            with torch.cuda.stream(branch):
                for _ in range(STEPS_PER_TRAJECTORY):
                    for r in range(b, replicas, streams):
                        trajs[r].step()
        for branch in branches:  # join
            stream.wait_stream(branch)

    with torch.cuda.stream(stream):
        run()  # eager warm-up: step plans, allocator, cuFFT plans
        torch.cuda.synchronize(device)
        for t in trajs:  # a captured region can only consume Philox streams reserved beforehand
            t.noise_source.reserve(2 * STEPS_PER_TRAJECTORY)
        before = native.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            run()
        launches_per_replay = native.launch_count() - before
    torch.cuda.synchronize(device)

    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(max(3, -(-warmup // rounds))):
            graph.replay()
        start.record()
        graph.replay()
        graph.replay()
        stop.record()
    torch.cuda.synchronize(device)
    replay_s = max(start.elapsed_time(stop) / 2e3, 1e-6)
    per_block = max(1, math.ceil(min_seconds / blocks / replay_s), math.ceil(requested_steps / rounds / blocks))
    barrier()
    times: list[float] = []
    with ClockSampler(device.index or 0) if clocks else NullClocks() as clock:
        with torch.cuda.stream(stream):
            for _ in range(blocks):
                start.record()
                for _ in range(per_block):
                    graph.replay()
                stop.record()
                stop.synchronize()
                times.append(start.elapsed_time(stop))
    torch.cuda.synchronize(device)
    barrier()
    block_ms = max_over_ranks(sorted(times)[len(times) // 2], device)
    steps_per_block = per_block * rounds
    return {
        "ms_per_step": block_ms / steps_per_block,
        "block_ms": block_ms,
        "blocks_ms": [round(t, 4) for t in times],
        "steps_per_block": steps_per_block,
        "timed_steps": steps_per_block * blocks,
        "launches": per_block * launches_per_replay * blocks,
        "launches_per_step": launches_per_replay / rounds,
        "replicas": replicas,
        "branches": streams,
        "bytes_per_step_avg": traj_bytes / STEPS_PER_TRAJECTORY,
        "bytes_per_step_max": max(per_step),
        "clocks": clock.summary(),
        "batch": spec["shape"][0],
    }


# BASELINE.json configs[2]: an explicit Runge-Kutta step (one launch per stage + the final update)
RK_ROWS = {
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
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    with torch.cuda.stream(stream):
        for _ in range(5):
            graph.replay()
        for _ in range(5):
            start.record()
            for _ in range(reps):
                graph.replay()
            stop.record()
            stop.synchronize()
            times.append(start.elapsed_time(stop) / reps)
    ms = sorted(times)[len(times) // 2]
    rk_steps = replicas * len(interior)
    return {"us_per_rk_step": ms * 1e3 / rk_steps, "launches_per_step": launches / rk_steps, "GBps": native.ACCOUNT["bytes"] / (ms * 1e-3) / 1e9, "bytes_per_step": native.ACCOUNT["bytes"] / rk_steps}


def noise_generator_times(device: torch.device, unit: tuple[int, ...] = (16, 21, 90, 160), reps: int = 30) -> list[dict]:
    """Device time per draw of each noise generator on one BASELINE.json configs[3] unit (a 16x21x90x160 video latent):
    the draws are captured in a CUDA graph (device time, no host launch gaps) and timed with CUDA events."""
    from skrample_b200.common import Step
    from skrample_b200.pytorch import noise

    rows = []
    step = Step.from_int(5, STEPS_PER_TRAJECTORY)
    for dtype in (torch.float32, torch.bfloat16):
        for name, props in (("Random", None), ("Offset", noise.OffsetProps()), ("Pyramid", noise.PyramidProps()), ("Colored", noise.ColoredProps()), ("Brownian", noise.BrownianProps())):
            cls = getattr(noise, name)
            generator = torch.Generator(device=device).manual_seed(1)
            source = cls.from_inputs(unit, generator, dtype=dtype) if props is None else cls.from_inputs(unit, generator, props, dtype=dtype)
            out = torch.empty(unit, dtype=dtype, device=device)
            stream = torch.cuda.Stream(device=device)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    source.generate_into(out, step)
                torch.cuda.synchronize(device)
                source.reserve(8)  # a captured region can only consume Philox streams reserved beforehand
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    for _ in range(4):
                        source.generate_into(out, step)
                for _ in range(3):
                    graph.replay()
                start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                start.record()
                for _ in range(reps):
                    graph.replay()
                stop.record()
            torch.cuda.synchronize(device)
            us = start.elapsed_time(stop) / (reps * 4) * 1e3
            rows.append({"generator": name, "unit_shape": list(unit), "dtype": str(dtype).removeprefix("torch."), "us_per_draw": us, "GBps_written": out.numel() * out.element_size() / us / 1e3})
    return rows


def pin_to_cores(rank: int, world: int) -> list[int]:
    "Give every rank of one node its own slice of the host cores (they all see the same set by default)."
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            mine = cores[rank * per : (rank + 1) * per]
            os.sched_setaffinity(0, mine)
            return mine
        return cores
    except (AttributeError, OSError):
        return []


def e2e_throughput(spec: dict, device: torch.device, min_steps: int, warmup: int, inflight: int, graphed: bool = False, min_seconds: float = 0.5) -> dict:
    """Public API with host buffers: H2D of the step's prediction, noise drawn on the device, the step, D2H of the result -
    every step.  ``inflight`` independent latent batches (requests) are advanced round robin: a request's next step
    starts only after its previous result has arrived in host memory, but while that copy is in flight the host
    prepares and launches the other requests' steps.  ``inflight=1`` is the plain synchronous loop.  ``graphed``: the
    step is ``GraphedTrajectory.step()`` instead of ``sampler.sample()``, same buffers and copies."""
    from skrample_b200.graphs import GraphedTrajectory

    trajs = [Trajectory(spec, device, seed=4321 + 17 * i) for i in range(inflight)]
    for traj in trajs:
        traj.record()
    per_step = step_bytes(spec, device, "auto")
    dtype = trajs[0].dtype
    host_pred = [[p.cpu().pin_memory() for p in traj.predictions] for traj in trajs]
    pred_dev = [torch.empty(spec["shape"], dtype=dtype, device=device) for _ in trajs]
    result_host = [torch.empty(spec["shape"], dtype=dtype).pin_memory() for _ in trajs]
    arrived = [torch.cuda.Event() for _ in trajs]
    pending = [False] * inflight
    need_noise = trajs[0].sampler.require_noise
    graphs = [GraphedTrajectory(t.sampler, t.model, t.schedule, STEPS_PER_TRAJECTORY, like=t.x0) for t in trajs] if graphed else []
    for g, t in zip(graphs, trajs):
        g.start(t.x0)
    current = torch.cuda.current_stream(device)
    # one stream per request, as a server would: the copy engines move request A's result out while request B's
    # prediction comes in and its step runs (on ONE stream the three would queue behind each other)
    lanes = [torch.cuda.Stream(device=device) for _ in trajs] if inflight > 1 else [current]
    for lane in lanes:
        lane.wait_stream(current)

    def one(k: int) -> None:
        slot = k % inflight
        traj = trajs[slot]
        if pending[slot]:
            arrived[slot].synchronize()  # the caller consumes this request's previous result before its next step
        if inflight > 1:
            torch.cuda.set_stream(lanes[slot])
        if graphed:
            g = graphs[slot]
            if g.position == len(g):
                g.start(traj.x0)
            g.prediction().copy_(host_pred[slot][g.position], non_blocking=True)
            if need_noise:
                traj.noise_source.generate_into(g.noise(), None)  # fresh noise every step, written in place
            final = g.step()
        else:
            pred_dev[slot].copy_(host_pred[slot][traj.n], non_blocking=True)
            final = traj.step(pred_dev[slot])  # draws this step's noise (Trajectory.draw)
        result_host[slot].copy_(final, non_blocking=True)
        if inflight == 1:
            current.synchronize()
        else:
            arrived[slot].record(lanes[slot])
            pending[slot] = True

    def drain() -> None:
        torch.cuda.synchronize(device)
        for slot in range(inflight):
            pending[slot] = False

    for k in range(max(warmup, 2 * STEPS_PER_TRAJECTORY * inflight)):
        one(k)
    drain()
    for traj in trajs:
        traj.reset()
    for g, t in zip(graphs, trajs):
        g.start(t.x0)
    drain()
    barrier()
    steps = 0
    t0 = time.perf_counter()
    while True:
        for k in range(steps, steps + 500):
            one(k)
        steps += 500
        if steps >= min_steps and time.perf_counter() - t0 >= min_seconds:
            break
    torch.cuda.synchronize(device)
    elapsed = time.perf_counter() - t0
    torch.cuda.set_stream(current)
    barrier()
    n = numel_of(spec["shape"])
    return {
        "elapsed_s": elapsed,
        "steps": steps,
        "bytes": sum(per_step[(k // inflight) % STEPS_PER_TRAJECTORY] for k in range(steps)),
        "h2d": n * dtype.itemsize,
        "d2h": n * dtype.itemsize,
    }


def measured_traffic(which: str = "traffic") -> float | None:
    """DRAM bytes per launch of the dominant step kernel on the default workload, from the latest ncu launch list
    summarised under profiles/ (tools/profile_round.sh + tools/summarize_profiles.py); None when not captured.
    ``traffic``: the kernel of the headline chain (noise drawn in the step); ``traffic_kernel_only``: the same step
    reading supplied noise tensors."""
    if "--workload" in sys.argv:
        return None
    found = sorted((ROOT / "profiles").glob(f"r[0-9][0-9]_{which}.json"))
    if not found:
        return None
    try:
        return float(json.loads(found[-1].read_text())["dram_bytes_per_launch"])
    except (OSError, ValueError, KeyError):
        return None


def barrier() -> None:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()


def leave_process_group(grace_s: float = 20.0, device: torch.device | None = None) -> None:
    """End a multi-rank run: wait for the other ranks, tear the process group down, exit 0 - without ever hanging.

    ``destroy_process_group()`` is not a synchronising call: a rank that tears its NCCL communicator down while its
    peers are already exiting can wait for them indefinitely (seen once at 8 ranks: the result line was out, six of
    the eight workers then sat in the teardown until the launcher was killed).  So the teardown runs beside a
    watchdog: barrier, destroy, and after ``grace_s`` at most the process leaves with status 0 either way - the
    measurement is complete and printed by then; interpreter teardown (CUDA graphs, pinned buffers, the communicator's
    own threads) is skipped on purpose."""
    import threading

    closed = threading.Event()

    def close() -> None:
        try:
            if device is not None and device.type == "cuda":
                torch.cuda.set_device(device)  # the current device is per thread
                torch.cuda.synchronize(device)
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        except Exception:  # noqa: BLE001 - leaving anyway
            pass
        closed.set()

    threading.Thread(target=close, daemon=True).start()
    closed.wait(grace_s)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def max_over_ranks(value: float, device: torch.device) -> float:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        t = torch.tensor([value], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())
    return value


def strong_scaling(device: torch.device, rank: int, world: int, peak: float) -> list[dict]:
    """A FIXED global batch split over the ranks in contiguous blocks of whole items; no collective on
    the step path; the noise of every step is drawn inside the timed region by the workload's generator (keyed by the
    item's global index).  value = global items x steps / slowest rank's time."""
    rows = []
    for title, global_batch, name in STRONG:
        base = WORKLOADS[name]
        if global_batch % world:
            rows.append({"workload": title, "skipped": f"global batch {global_batch} does not split over {world} ranks"})
            continue
        per_rank = global_batch // world
        spec = dict(base, shape=(per_rank, *base["shape"][1:]))
        torch.cuda.empty_cache()
        r = chain_time(spec, device, "auto", 0, STEPS_PER_TRAJECTORY, min_seconds=0.3, blocks=3, first_item=rank * per_rank)
        kernel = chain_time(spec, device, "supplied", 0, STEPS_PER_TRAJECTORY, min_seconds=0.2, blocks=3, first_item=rank * per_rank)
        us = r["ms_per_step"] * 1e3
        rows.append(
            {
                "workload": title,
                "global_batch": global_batch,
                "per_gpu_batch": per_rank,
                "n_gpus": world,
                "scaling": "strong",
                "value": global_batch / (us * 1e-6),
                "unit": "latent-steps/s",
                "us_per_step": us,
                "us_per_step_kernel_only": kernel["ms_per_step"] * 1e3,
                "launches_per_step": r["launches_per_step"],
                "sampler_step_GBps_per_gpu": kernel["bytes_per_step_avg"] / (us * 1e-6) / 1e9,
                "kernel_only_frac_of_measured_peak": kernel["bytes_per_step_avg"] / (kernel["ms_per_step"] * 1e-3) / 1e9 / peak,
            }
        )
    return rows


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores


def reference_modules():  # noqa: ANN201
    "The unmodified reference package from baseline/_ref (tools/install_reference.py), or None."
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "skrample" / "__init__.py").exists():
        return None
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    try:
        import skrample.sampling.models as ref_models
        import skrample.sampling.structured as ref_structured
        import skrample.scheduling as ref_scheduling
        from skrample.common import Point as RefPoint
        from skrample.common import Step as RefStep
    except Exception:  # noqa: BLE001 - any import problem means "use the port"
        return None
    return ref_structured, ref_models, ref_scheduling, RefStep, RefPoint


def cpu_reference_steps(spec: dict, steps: int, warmup: int, budget_s: float) -> dict:
    """The workload through the reference's own ``sampler.sample`` (torch-CPU tensors, every host thread), sampler
    time only; the oracle port (oracle/skrample_oracle.py) stands in when baseline/_ref is absent."""
    import cases

    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(threads)
    shape = spec["shape"]
    g = torch.Generator().manual_seed(7)
    mods = reference_modules()
    case = {"sampler": spec["sampler"], "kw": spec["kw"], "schedule": spec["schedule"], "model": spec["model"]}
    per_step_bytes = reference_step_bytes(spec)
    if mods is not None:
        ref_structured, ref_models, ref_scheduling, RefStep, RefPoint = mods
        sampler = cases.make_sampler(ref_structured, ref_models, case)
        schedule = cases.make_schedule(ref_scheduling, spec["schedule"])
        model = cases.make_model(ref_models, spec["model"])
        points = [tuple(p) for p in schedule.schedule(STEPS_PER_TRAJECTORY)]
        need_noise, need_prev = sampler.require_noise, sampler.require_previous
        kind = "reference"
        what = "the unmodified reference package (baseline/_ref/skrample), sampler.sample on torch-CPU fp32 tensors"
    else:
        import oracle_run
        from oracle import skrample_oracle as O

        model = oracle_run.MODELS[spec["model"]]
        sch = oracle_run.schedule(spec["schedule"])
        need_noise, need_prev = oracle_run.require(case)
        points = [tuple(p) for p in sch.schedule(STEPS_PER_TRAJECTORY)]
        kind = "port"
        what = "the oracle port (oracle/skrample_oracle.py) on torch-CPU fp32 tensors"
    x0 = torch.randn(shape, generator=g) * points[0][1]
    distinct = max(2, min(STEPS_PER_TRAJECTORY, (1 << 28) // (4 * numel_of(shape))))  # huge latents: a few draws, reused
    noises = [torch.randn(shape, generator=g) for _ in range(distinct)]
    noises = [noises[n % distinct] for n in range(STEPS_PER_TRAJECTORY)]
    state = {"x": x0, "n": 0, "prev": []}

    def one() -> float:
        n = state["n"]
        _, sigma, alpha = points[n]
        x = state["x"]
        xhat = x * (alpha / (alpha * alpha + sigma * sigma))  # analytic denoiser, excluded from timing
        if mods is not None:
            RefStep_ = mods[3]
            out = xhat if spec["model"] == "DataModel" else model.from_x(x, xhat, RefPoint(*points[n]))
            t0 = time.perf_counter()
            res = sampler.sample(x, out, RefStep_.from_int(n, STEPS_PER_TRAJECTORY), model, schedule, noises[n] if need_noise else None, state["prev"])
            dt = time.perf_counter() - t0
            final = res.final
        else:
            point = O.Pt(*points[n])
            out = model.from_x(x, xhat, point) if model.kind != "data" else xhat
            cur = O.Rec(x, out, O.St.from_int(n, STEPS_PER_TRAJECTORY), noises[n] if need_noise else None)
            t0 = time.perf_counter()
            res = oracle_run.one_step(case, cur, state["prev"], model, sch)
            dt = time.perf_counter() - t0
            final = res.final
        state["prev"] = (state["prev"] + [res])[-need_prev:] if need_prev else []
        state["x"] = final
        state["n"] = (n + 1) % STEPS_PER_TRAJECTORY
        if state["n"] == 0:
            state["x"], state["prev"] = x0, []
        return dt

    for _ in range(warmup):
        one()
    state.update(x=x0, n=0, prev=[])
    spent, done, total_bytes = 0.0, 0, 0
    wall0 = time.perf_counter()
    while done < steps and (time.perf_counter() - wall0) < budget_s:
        total_bytes += per_step_bytes[state["n"]]
        spent += one()
        done += 1
    return {"seconds": spent, "steps": done, "bytes": total_bytes, "threads": threads, "kind": kind, "what": what}


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


def cpu_baseline_of(res: dict, spec: dict) -> dict:
    return {
        "value": res["steps"] * spec["shape"][0] / res["seconds"],
        "unit": "latent-steps/s",
        "cores": res["threads"],
        "kind": res["kind"],
        "sample": f"{res['steps']} sampler steps of the workload in fp32: {res['what']}, sampler time only",
        "sampler_step_GBps": res["bytes"] / res["seconds"] / 1e9,
        "ms_per_step": res["seconds"] / res["steps"] * 1e3,
    }


# --------------------------------------------------------------------------------------------------------------


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--quick", action="store_true", help="only the headline measurements (no rows, noise generators, strong-scaling legs)")
    ap.add_argument("--sweep", action="store_true", help="accepted for compatibility: the rows are part of the default line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4, help="parallel graph branches of the `concurrent_requests` leg")
    ap.add_argument("--inflight", type=int, default=0, help="independent requests advanced round robin by the e2e leg (1: the synchronous loop only; 0: the best of 1 / 2 / 4 by a short trial)")
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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    spec = WORKLOADS[args.workload]
    config = config_of(spec, args.workload, world)

    if args.impl == "reference":
        if rank != 0:
            return
        warm = min(max(args.warmup, 3 * STEPS_PER_TRAJECTORY), 8 * STEPS_PER_TRAJECTORY)  # the CPU path needs whole trajectories to warm its thread pool and allocator
        res = cpu_reference_steps(spec, args.steps, warm, 120.0)
        base = cpu_baseline_of(res, spec)
        emit(
            {
                "impl": "reference",
                "metric": METRIC,
                "value": base["value"],
                "unit": "latent-steps/s",
                "n_gpus": args.gpus,
                "steps": args.steps,
                "timed_steps": res["steps"],
                "warmup": warm,
                "ms_per_step": base["ms_per_step"],
                "higher_is_better": True,
                "scaling": "weak",
                "vs_baseline": None,
                "dtype": "f32",
                "data": "synthetic",
                "config": config,
                "sampler_step_GBps": base["sampler_step_GBps"],
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "latent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
        )
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for --impl ours (there is no CPU fallback)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    cores = pin_to_cores(local, world)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)

    from skrample_b200 import native

    native.load()
    peak, peak_src = measured_peak()
    batch = spec["shape"][0]
    launches_before = native.launch_count()

    # headline: one dependent chain, the noise of every step drawn inside the timed region
    chain = chain_time(spec, device, "auto", args.steps, args.warmup, clocks=True)
    ms_per_step = chain["ms_per_step"]
    value = batch * world / (ms_per_step * 1e-3)
    gbs = chain["bytes_per_step_avg"] * world / (ms_per_step * 1e-3) / 1e9
    # the same chain with the noise written beforehand: one launch per step, the step kernel's own duration
    torch.cuda.empty_cache()
    kernel = chain_time(spec, device, "supplied", args.steps, args.warmup)
    headline_us_per_launch = ms_per_step * 1e3 / chain["launches_per_step"]
    headline_bytes_per_launch = chain["bytes_per_step_avg"] / chain["launches_per_step"]
    headline_achieved = headline_bytes_per_launch / (headline_us_per_launch * 1e-6) / 1e9
    per_launch_us = kernel["ms_per_step"] * 1e3 / kernel["launches_per_step"]
    bytes_per_launch = kernel["bytes_per_step_avg"] / kernel["launches_per_step"]
    achieved = bytes_per_launch / (per_launch_us * 1e-6) / 1e9
    concurrent = None
    if args.streams > 1:
        torch.cuda.empty_cache()
        many = chain_time(spec, device, "auto", args.steps, args.warmup, streams=args.streams)
        concurrent = {
            "value": batch * world / (many["ms_per_step"] * 1e-3),
            "unit": "latent-steps/s",
            "inverse_throughput_ms_per_step": many["ms_per_step"],
            "graph_branches": many["branches"],
            "sampler_step_GBps": many["bytes_per_step_avg"] * world / (many["ms_per_step"] * 1e-3) / 1e9,
            "note": "independent latent batches on parallel graph branches: a multi-request throughput, not a step latency",
        }
    torch.cuda.empty_cache()

    single = e2e_throughput(spec, device, 2000, min(args.warmup, 100), 1)
    single_elapsed = max_over_ranks(single["elapsed_s"], device)
    e2e, e2e_elapsed = single, single_elapsed
    # How many requests a server keeps in flight is a deployment knob, and the best value depends on what else shares the
    # host's PCIe uplinks: alone on the node 4 requests hide both copies behind each other (38 us/step against 68 for
    # the synchronous loop), with 8 ranks pushing 2 MiB per step each through the same host the queued copies slow each
    # other down (133 us at 4 in flight against 83 synchronous).  --inflight 0 (the default) tries 2 and 4 for a short
    # trial each and times the best of {1, 2, 4} - judged by the slowest rank, the same choice on every rank.
    trials = {1: single_elapsed / single["steps"] * 1e3}
    inflight = args.inflight
    if inflight <= 0:
        for candidate in (2, 4):
            trial = e2e_throughput(spec, device, 500, min(args.warmup, 100), candidate, min_seconds=0.15)
            trials[candidate] = max_over_ranks(trial["elapsed_s"], device) / trial["steps"] * 1e3
        inflight = min(trials, key=trials.get)
    if inflight > 1:
        e2e = e2e_throughput(spec, device, 2000, min(args.warmup, 100), inflight)
        e2e_elapsed = max_over_ranks(e2e["elapsed_s"], device)
    graphed = e2e_throughput(spec, device, 2000, min(args.warmup, 100), inflight, graphed=True)
    graphed_elapsed = max_over_ranks(graphed["elapsed_s"], device)

    line = {
        "metric": METRIC,
        "value": value,
        "unit": "latent-steps/s",
        "n_gpus": world,
        "steps": args.steps,
        "timed_steps": chain["timed_steps"],
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": config,
        "launch": {
            "how": "one dependent chain of sampler launches on one stream, replayed from a CUDA graph; the median of 5 timed blocks",
            "noise": "a fresh draw every step inside the timed region: Philox keys consumed by the step kernel where BatchTensorNoise.auto picks that, else written by the generator's kernels",
            "launches_per_step": chain["launches_per_step"],
            "blocks_ms": chain["blocks_ms"],
            "steps_per_block": chain["steps_per_block"],
            "l2": f"{chain['replicas']} interleaved latent batches, working set per round > 2x L2 (inputs come from HBM)",
            "storage_dtype": spec["dtype"],
            "host_cores_of_rank0": len(cores),
        },
        "sampler_step_GBps": gbs,
        "pct_of_hbm_peak": {"measured": gbs / world / peak, "nominal_8TBs": gbs / world / 8000.0},
        "gpu_launches": chain["launches"],
        "clocks": chain["clocks"],
        "kernel_only": {
            "value": batch * world / (kernel["ms_per_step"] * 1e-3),
            "unit": "latent-steps/s",
            "ms_per_step": kernel["ms_per_step"],
            "launches_per_step": kernel["launches_per_step"],
            "sampler_step_GBps": kernel["bytes_per_step_avg"] * world / (kernel["ms_per_step"] * 1e-3) / 1e9,
            "note": "the noise tensors are written before the timed region (what the CPU arm gets): the step kernel alone",
        },
        "concurrent_requests": concurrent,
        "e2e": {
            "value": e2e["steps"] * batch * world / e2e_elapsed,
            "unit": "latent-steps/s",
            "h2d_bytes_per_step": e2e["h2d"],
            "d2h_bytes_per_step": e2e["d2h"],
            "sampler_step_GBps": e2e["bytes"] * world / e2e_elapsed / 1e9,
            "ms_per_step": e2e_elapsed / e2e["steps"] * 1e3,
            "steps": e2e["steps"],
            "api": "structured sampler .sample() per step (the reference's call), pinned-host prediction in, result out; noise from BatchTensorNoise.auto",
            "requests_in_flight": inflight,
            "in_flight_trials_ms_per_step": {str(k): v for k, v in sorted(trials.items())},
            "one_request": {
                "value": single["steps"] * batch * world / single_elapsed,
                "ms_per_step": single_elapsed / single["steps"] * 1e3,
                "steps": single["steps"],
                "note": "the plain synchronous loop: every step waits for its result on the host before the next begins",
            },
        },
        "e2e_graphed": {
            "value": graphed["steps"] * batch * world / graphed_elapsed,
            "unit": "latent-steps/s",
            "ms_per_step": graphed_elapsed / graphed["steps"] * 1e3,
            "steps": graphed["steps"],
            "requests_in_flight": inflight,
            "api": "skrample_b200.graphs.GraphedTrajectory.step(): same host buffers and copies, launches replayed from CUDA graphs",
        },
        # the dominant kernel of the timed region of `value`: the step kernel that draws its own noise (one launch per
        # step; when the library fills the noise first, launches_per_step > 1 and the per-launch figures are averages)
        "roofline": {
            "bound": "hbm",
            "achieved": headline_achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": headline_achieved / peak,
            "traffic": measured_traffic("traffic"),
            "kernel": "skr::block_kernel (the step with its noise drawn in the kernel: no noise tensor is read or written)",
            "bytes_per_launch": headline_bytes_per_launch,
            "us_per_launch": headline_us_per_launch,
            "launches_per_step": chain["launches_per_step"],
            "measured_on": "the headline chain (`value`), CUDA events around the timed blocks",
            "peak_source": peak_src,
        },
        # the same step reading supplied noise tensors (the round-1 definition: more bytes per step, less arithmetic)
        "roofline_supplied_noise": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": measured_traffic("traffic_kernel_only"),
            "kernel": "skr::block_kernel",
            "bytes_per_launch": bytes_per_launch,
            "us_per_launch": per_launch_us,
            "measured_on": "the `kernel_only` chain (one step-kernel launch per step), CUDA events around the timed blocks",
            "peak_source": peak_src,
        },
    }
    assert native.launch_count() > launches_before

    if not args.quick:
        # the other BASELINE shapes; every rank runs them (they must stay in lock-step for the strong-scaling legs)
        rows = []
        if world == 1:
            for name in ROWS:
                s = WORKLOADS[name]
                torch.cuda.empty_cache()
                k = chain_time(s, device, "supplied", 0, STEPS_PER_TRAJECTORY, min_seconds=0.15, blocks=3)
                w = chain_time(s, device, "auto", 0, STEPS_PER_TRAJECTORY, min_seconds=0.15, blocks=3)
                us_k, us_w = k["ms_per_step"] * 1e3, w["ms_per_step"] * 1e3
                ach = k["bytes_per_step_avg"] / (us_k * 1e-6) / 1e9
                rows.append(
                    {
                        "workload": name,
                        "shape": list(s["shape"]),
                        "dtype": s["dtype"],
                        "noise": s["noise"],
                        "us_per_step_kernel_only": us_k,
                        "GBps": ach,
                        "frac_of_measured_peak": ach / peak,
                        "frac_of_nominal_8TBs": ach / 8000.0,
                        "bytes_per_step_avg": k["bytes_per_step_avg"],
                        "us_per_step_with_noise": us_w,
                        "launches_per_step_with_noise": w["launches_per_step"],
                        "latent_steps_per_s_with_noise": s["shape"][0] / (us_w * 1e-6),
                    }
                )
            for name in CONTRACTED_ROWS:  # the issue-bound steps again with the opt-in contracted arithmetic
                s = WORKLOADS[name]
                torch.cuda.empty_cache()
                native.set_arithmetic("contracted")
                try:
                    k = chain_time(s, device, "supplied", 0, STEPS_PER_TRAJECTORY, min_seconds=0.15, blocks=3)
                finally:
                    native.set_arithmetic("exact")
                us_k = k["ms_per_step"] * 1e3
                ach = k["bytes_per_step_avg"] / (us_k * 1e-6) / 1e9
                rows.append({"workload": name + " (arithmetic: contracted, opt-in)", "shape": list(s["shape"]), "dtype": s["dtype"], "noise": s["noise"], "us_per_step_kernel_only": us_k, "GBps": ach, "frac_of_measured_peak": ach / peak, "bytes_per_step_avg": k["bytes_per_step_avg"]})
            for name, s in RK_ROWS.items():
                torch.cuda.empty_cache()
                r = rk_step_throughput(s, device)
                rows.append({"workload": name, "shape": list(s["shape"]), "dtype": s["dtype"], "noise": "none (ODE)", "us_per_step_kernel_only": r["us_per_rk_step"], "launches_per_step": r["launches_per_step"], "GBps": r["GBps"], "frac_of_measured_peak": r["GBps"] / peak, "bytes_per_step_avg": r["bytes_per_step"], "latent_steps_per_s": s["shape"][0] / (r["us_per_rk_step"] * 1e-6)})
            torch.cuda.empty_cache()
            if not args.no_cpu_baseline:
                # configs[4] "vs host-CPU reference": the same Euler SDE step through the reference on the host cores,
                # a bounded sample per size (two warm-up steps, at most ~2 s)
                for row in rows:
                    if row["workload"].startswith("euler_sde_flow_f32"):
                        res = cpu_reference_steps(WORKLOADS[row["workload"]], 50, 2, 2.0)
                        row["cpu_reference_us_per_step"] = res["seconds"] / res["steps"] * 1e6
                        row["cpu_reference_steps_timed"] = res["steps"]
                        row["speedup_with_noise_vs_cpu_reference"] = row["cpu_reference_us_per_step"] / row["us_per_step_with_noise"]
            line["rows"] = rows
            line["noise_generators"] = noise_generator_times(device)
        torch.cuda.empty_cache()
        line["strong_scaling"] = strong_scaling(device, rank, world, peak)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_of(cpu_reference_steps(spec, 10_000, 4 * STEPS_PER_TRAJECTORY, 15.0), spec)

    if rank == 0:
        emit(line)
    if world > 1:
        leave_process_group(device=device)


if __name__ == "__main__":
    main()
