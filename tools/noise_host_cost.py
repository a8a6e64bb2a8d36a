"""Host cost of one BatchTensorNoise draw, piece by piece (development aid)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from skrample_b200.pytorch import noise

dev = torch.device("cuda", 0)
gens = [torch.Generator(device=dev).manual_seed(i) for i in range(8)]
src = noise.BatchTensorNoise.from_batch_inputs(noise.Random, (4, 128, 128), gens, dtype=torch.float32)
out = torch.empty((8, 4, 128, 128), device=dev, dtype=torch.bfloat16)


def timeit(name, fn, n=2000):
    for _ in range(100):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    print(f"{name:40s} {dt:7.2f} us")


g = gens[0]
timeit("get_offset", lambda: g.get_offset())
timeit("set_offset", lambda: g.set_offset(1024))
timeit("initial_seed", lambda: g.initial_seed())
timeit("g.device", lambda: g.device)
r = src.generators[0]
timeit("Random._tick", lambda: r._tick())
timeit("Random._key", lambda: r._key())
timeit("_uniform_random", lambda: src._uniform_random())
timeit("lazy (keys + ticks of 8 generators)", lambda: src.lazy(None, _fallback=False))
timeit("torch.empty", lambda: torch.empty((8, 4, 128, 128), device=dev, dtype=torch.float32))
d = src.lazy(None, _fallback=False)
timeit("materialize_into", lambda: d.materialize_into(out))
timeit("generate", lambda: src.generate(None))
timeit("generate_into", lambda: src.generate_into(out, None))
