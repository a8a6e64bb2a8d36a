"""Print which kernel shape each step of the common samplers would run with (host-only: no GPU needed).

    python tools/describe_steps.py [bf16|f16|f32] [--wrapper]

The step programs are emitted exactly as for CUDA tensors (the device check is patched), described by
skr_program_describe and then executed by the host executor so the trajectory advances."""
from __future__ import annotations

import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from skrample_b200 import native, scheduling
from skrample_b200.common import Step
from skrample_b200.sampling import models, structured
from skrample_b200.sampling import program as pg


def out_dtypes(program: pg.Program) -> list[torch.dtype]:
    "Output dtypes as the device path allocates them: solver state in fp32, the rest in the requested / sample dtype."
    first = program.inputs[0].dtype
    return [d if isinstance(d, torch.dtype) else (torch.float32 if d == structured.COMPUTE else first) for d in program.outputs]


def describe(program: pg.Program) -> str:
    tensors = list(program.inputs)
    outs = [torch.empty(0, dtype=d) for d in out_dtypes(program)]
    packed = native.pack_program(program, tensors, outs)
    text = ctypes.create_string_buffer(1024)
    native.load().skr_program_describe(ctypes.byref(packed), text, len(text))
    return text.value.decode()


def main() -> None:
    dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[next((a for a in sys.argv[1:] if not a.startswith("--")), "bf16")]
    wrapper = "--wrapper" in sys.argv
    seen: list[str] = []

    def spy(program: pg.Program) -> list:
        seen.append(describe(program))
        return [out.to(d) for out, d in zip(pg.execute_generic(program), out_dtypes(program))]

    pg.execute = spy  # type: ignore[assignment]
    pg.is_cuda_tensor = lambda v: isinstance(v, torch.Tensor)  # type: ignore[assignment]
    structured.pg = pg
    samplers = {
        "Euler": structured.Euler(),
        "Euler sde": structured.Euler(stochasticity=1),
        "DPM2": structured.DPM(order=2),
        "DPM3 sde": structured.DPM(order=3, stochasticity=1),
        "Adams4": structured.Adams(order=4),
        "UniP3": structured.UniP(order=3),
        "UniPC3 sde": structured.UniPC(order=3, stochasticity=1),
        "SPC": structured.SPC(),
    }
    for name, sampler in samplers.items():
        for model, schedule in ((models.NoiseModel(), scheduling.Scaled()), (models.FlowModel(), scheduling.FlowShift(scheduling.Linear()))):
            seen.clear()
            x = torch.randn(64).to(dtype)
            previous: list = []
            options = structured.step_options(final_dtype=dtype) if wrapper else structured.step_options()
            with options:
                for n in range(6):
                    res = sampler.sample(x, torch.randn(64).to(dtype), Step.from_int(n, 8), model, schedule, torch.randn(64).to(dtype), previous)
                    previous = (previous + [res])[-sampler.require_previous :] if sampler.require_previous else []
                    x = res.final.to(dtype)
            print(f"== {name} / {type(model).__name__}")
            for n, line in enumerate(seen):
                print(f"  step {n}: {line}")


if __name__ == "__main__":
    main()
