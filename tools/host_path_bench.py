"""Python-side cost of a plan-cache hit of ``sampler.sample`` WITHOUT a GPU (development aid).

The device launch is replaced by a stub that hands back pre-made tensors, so what is timed is exactly the host work of
the hit path: SampleInput, key, lookup, bind, result assembly.  Run on the build container to iterate on that path.

    python tools/host_path_bench.py [--profile]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch

from skrample_b200 import native, scheduling
from skrample_b200.common import Step
from skrample_b200.sampling import models, plan, structured
from skrample_b200.sampling import program as pg

pg.is_cuda_tensor = lambda v: isinstance(v, torch.Tensor)  # emit and cache what a device run would
pg._fusable = lambda values: True
structured.pg = pg
pool = {}


def stub(compiled, inputs, draws=None):  # noqa: ANN001, ANN201
    outs = pool.get(id(compiled))
    if outs is None:
        first = inputs[0]
        outs = pool[id(compiled)] = [torch.empty_like(first, dtype=torch.float32) for _ in compiled.out_specs]
    return list(outs)


class FakeCompiled:
    def __init__(self, program) -> None:  # noqa: ANN001
        self.n_inputs = len(program.inputs)
        self.out_specs = tuple(program.outputs)
        self.fast = {}  # no native plans: the Python hit path is what this tool times
        self.n_philox = 0


native.launch_compiled = stub
native.CompiledProgram = FakeCompiled
pg.execute = lambda program: [torch.empty_like(program.inputs[0], dtype=torch.float32) for _ in program.outputs]

sampler = structured.UniPC(order=3, stochasticity=1)
schedule, model = scheduling.Scaled(), models.NoiseModel()
shape = (8, 4, 128, 128)
x0 = torch.randn(shape).bfloat16()
preds = [torch.randn(shape).bfloat16() for _ in range(25)]
noises = [torch.randn(shape) for _ in range(25)]
steps = [Step.from_int(n, 25) for n in range(25)]


def trajectory() -> None:
    x, previous = x0, []
    keep = sampler.require_previous
    for n in range(25):
        res = sampler.sample(x, preds[n], steps[n], model, schedule, noises[n], previous)
        previous = (previous + [res])[-keep:]
        x = res.final


for _ in range(4):
    trajectory()
print("plan stats (hits, misses):", plan.stats())
reps = 400
t0 = time.perf_counter()
for _ in range(reps):
    trajectory()
dt = time.perf_counter() - t0
print(f"{dt / (reps * 25) * 1e6:.2f} us per sampler.sample (host work of a plan hit, launch stubbed)")
if "--profile" in sys.argv:
    import cProfile
    import pstats

    prof = cProfile.Profile()
    prof.enable()
    for _ in range(100):
        trajectory()
    prof.disable()
    pstats.Stats(prof).sort_stats("tottime").print_stats(18)
