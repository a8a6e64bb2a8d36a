"""CUDA-graph capture of a whole sampling trajectory.

At SD/SDXL latent sizes one solver step is a 3-7 us kernel under 40-100 us of Python (dataclasses, plan lookup,
output allocation, ctypes).  The step programs of a trajectory depend only on (sampler, model, schedule, steps):
they can be recorded once and replayed.  :class:`GraphedTrajectory` runs the reference's structured-sampler loop
(reference: skrample/sampling/interface.py:23-59, the loop every structured sampler is driven by) once under stream
capture - one ``torch.cuda.CUDAGraph`` per step, all in one memory pool - and afterwards a step costs two optional
device copies and one graph launch.  Results are the bits of the eager loop: the same kernels run on the same data.

    traj = GraphedTrajectory(structured.UniPC(order=3, stochasticity=1), models.NoiseModel(), scheduling.Scaled(),
                             steps=25, like=latents)
    x = traj.start(latents)
    for n in range(25):
        traj.prediction(n).copy_(unet(x, *traj.point(n)))      # or: traj.step(prediction=..., noise=...)
        traj.noise(n).normal_()                                 # any device-side producer writing into the buffer
        x = traj.step()

Only CUDA tensors are supported (a missing native library raises, as everywhere else in the package); the buffers
returned by ``prediction(n)`` / ``noise(n)`` / ``step()`` belong to the trajectory and are overwritten by later steps
and by the next ``start``.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import torch

from skrample_b200.common import Point, Step
from skrample_b200.sampling import models
from skrample_b200.sampling.structured import SKSamples, StructuredSampler, step_options
from skrample_b200.scheduling import SkrampleSchedule


class GraphedTrajectory:
    "One structured sampler, one schedule, ``steps`` solver steps over a latent shaped like ``like``, as CUDA graphs."

    def __init__(
        self,
        sampler: StructuredSampler,
        model_transform: models.DiffusionModel,
        schedule: SkrampleSchedule,
        steps: int,
        like: torch.Tensor,
        include: slice = slice(None),
        final_dtype: torch.dtype | None = None,
    ) -> None:
        if not (isinstance(like, torch.Tensor) and like.is_cuda):
            raise ValueError("GraphedTrajectory captures device launches: `like` must be a CUDA tensor")
        self.sampler = sampler
        self.model_transform = model_transform
        self.schedule = schedule
        self.steps = steps
        self.indices: list[int] = list(range(steps))[include]
        if not self.indices:
            raise ValueError("empty trajectory")
        self.device = like.device
        self._final_dtype = final_dtype
        # Buffers a later step may still read: the previous step's noise (UniPC corrector) and, for models whose
        # conversion is the identity, the raw predictions of the whole history window.
        ring = sampler.require_previous + 1
        self._ring = ring
        self._sample = torch.empty_like(like, memory_format=torch.contiguous_format)
        self._predictions = [torch.empty_like(self._sample) for _ in range(ring)]
        self._noises = [torch.empty_like(self._sample) for _ in range(ring)] if sampler.require_noise else []
        self._graphs: list[torch.cuda.CUDAGraph] = []
        self._results: list[SKSamples] = []
        self._cursor = 0
        self._capture()

    # -- capture ------------------------------------------------------------------------------------------
    def _run_step(self, k: int, sample: torch.Tensor, previous: Sequence[SKSamples]) -> SKSamples:
        n = self.indices[k]
        noise = self._noises[k % self._ring] if self._noises else None
        with step_options(final_dtype=self._final_dtype):
            return self.sampler.sample(
                sample, self._predictions[k % self._ring], Step.from_int(n, self.steps), self.model_transform, self.schedule, noise, previous
            )

    def _walk(self, capture: bool, stream: torch.cuda.Stream, pool: Any) -> None:
        sample: torch.Tensor = self._sample
        previous: list[SKSamples] = []
        keep = self.sampler.require_previous
        for k in range(len(self.indices)):
            if capture:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, pool=pool, stream=stream):
                    result = self._run_step(k, sample, previous)
                self._graphs.append(graph)
                self._results.append(result)  # keeps every captured output alive: nothing in the pool is recycled
            else:
                result = self._run_step(k, sample, previous)
            if keep:
                previous = [*previous, result][-keep:]
            sample = result.final

    def _capture(self) -> None:
        with torch.cuda.device(self.device):
            stream = torch.cuda.Stream(device=self.device)
            stream.wait_stream(torch.cuda.current_stream(self.device))
            for buffer in (self._sample, *self._predictions, *self._noises):
                buffer.zero_()  # capture and warm-up run on whatever is in the buffers; keep it finite
            with torch.cuda.stream(stream):
                self._walk(False, stream, None)  # eager warm-up: plan cache, allocator, lazy module state
                stream.synchronize()
                self._walk(True, stream, torch.cuda.graph_pool_handle())
            torch.cuda.current_stream(self.device).wait_stream(stream)

    # -- replay -------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.indices)

    @property
    def position(self) -> int:
        "How many steps of the current trajectory have been taken."
        return self._cursor

    def point(self, k: int | None = None) -> Point:
        "Schedule point (timestep, sigma, alpha) the network is evaluated at for step ``k`` (default: the next one)."
        k = self._cursor if k is None else k
        return self.schedule.ipoint(Step.from_int(self.indices[k], self.steps).time_from)

    def prediction(self, k: int | None = None) -> torch.Tensor:
        "The buffer step ``k`` (default: the next one) reads the network output from."
        return self._predictions[(self._cursor if k is None else k) % self._ring]

    def noise(self, k: int | None = None) -> torch.Tensor:
        "The buffer step ``k`` (default: the next one) reads its noise from."
        if not self._noises:
            raise ValueError(f"{type(self.sampler).__name__} takes no noise")
        return self._noises[(self._cursor if k is None else k) % self._ring]

    def start(self, sample: torch.Tensor) -> torch.Tensor:
        "Begin a trajectory from ``sample`` (copied into the trajectory's own buffer, which is returned)."
        self._sample.copy_(sample, non_blocking=True)
        self._cursor = 0
        return self._sample

    def step(self, prediction: torch.Tensor | None = None, noise: torch.Tensor | None = None) -> torch.Tensor:
        "Advance one step on the current stream; ``prediction`` / ``noise`` are copied in unless already in place."
        k = self._cursor
        if k >= len(self._graphs):
            raise IndexError("trajectory finished: call start() to begin another")
        if prediction is not None and prediction is not self._predictions[k % self._ring]:
            self._predictions[k % self._ring].copy_(prediction, non_blocking=True)
        if noise is not None and self._noises and noise is not self._noises[k % self._ring]:
            self._noises[k % self._ring].copy_(noise, non_blocking=True)
        self._graphs[k].replay()
        self._cursor = k + 1
        return self._results[k].final

    def result(self, k: int) -> SKSamples:
        "Everything step ``k`` produced (final, corrected sample, x-hat); valid after ``step`` number ``k``."
        return self._results[k]
