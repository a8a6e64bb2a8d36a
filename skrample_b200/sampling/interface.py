"""Drive a structured (step-at-a-time) sampler through the functional (closure) interface.

reference: skrample/sampling/interface.py:13-59.  Every step of the loop is one fused launch of the wrapped
sampler; the ``previous`` window is trimmed to what the sampler declares it needs.
"""

from __future__ import annotations

import dataclasses

from skrample_b200 import scheduling
from skrample_b200.common import RNG, DeltaPoint, Point, Sample, Step

from . import functional, models, structured


@dataclasses.dataclass(frozen=True)
class StructuredFunctionalAdapter(functional.FunctionalSampler):
    sampler: structured.StructuredSampler

    def add_noise[T: Sample](self, sample: T, noise: T, point: Point) -> T:
        return self.sampler.add_noise(sample, noise, point)

    def remove_noise[T: Sample](self, sample: T, noise: T, point: Point) -> T:
        return self.sampler.remove_noise(sample, noise, point)

    def sample_model[T: Sample](
        self,
        sample: T,
        model: functional.SampleableModel[T],
        model_transform: models.DiffusionModel,
        schedule: scheduling.SkrampleSchedule,
        steps: int,
        include: slice = slice(None),
        rng: RNG[T] | None = None,
        callback: functional.SampleCallback | None = None,
    ) -> T:
        sampler = self.sampler
        points = schedule.schedule(steps)
        total = len(points)
        keep = sampler.require_previous
        wants_noise = rng is not None and sampler.require_noise
        clean = Point(0, 0, 1)
        history: list[structured.SKSamples[T]] = []

        for n in list(range(total))[include]:
            point = points[n]
            step = Step.from_int(n, total)
            done = sampler.sample_packed(
                structured.SampleInput(
                    sample=sample,
                    prediction=model(sampler.scale_input(sample, point), *point),
                    step=step,
                    noise=rng(step) if wants_noise else None,  # type: ignore[misc]
                ),
                model_transform,
                schedule,
                previous=history,
            )
            if keep > 0:
                history.append(done)
                del history[: max(len(history) - keep, 0)]
            sample = done.final
            if callback:
                callback(sample, n, DeltaPoint(point, points[n + 1] if n + 1 < total else clean))
        return sample
