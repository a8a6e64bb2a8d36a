"""Build ``csrc/libskrample_b200.so`` in-tree with nvcc for sm_100a.

    python -m skrample_b200.build [--force] [-v] [-DNAME[=VALUE] ... --out=path/to/variant.so]

nvcc cross-compiles without a GPU.  The shared library stays in the source tree
(git-ignored) so it travels with the repository snapshot to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
ROOT = CSRC.parent.parent
LIB = CSRC / "libskrample_b200.so"
# (source, extra defines, object name): pinned_shapes.cu is compiled once per latent storage type
UNITS = [
    ("step_kernel.cu", (), "step_kernel.o"),
    ("noise_kernels.cu", (), "noise_kernels.o"),
    ("pinned_shapes.cu", ("SKR_LP=0",), "pinned_f32.o"),
    ("pinned_shapes.cu", ("SKR_LP=2",), "pinned_bf16.o"),
    ("pinned_shapes.cu", ("SKR_LP=3",), "pinned_f16.o"),
]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-fmad=false",  # every product and sum is individually rounded, like the reference's separate ATen ops
    "-Xcompiler",
    "-fPIC",
]


FAST = CSRC.parent / "_fast.so"
"CPython extension with the plan-cache hit path (csrc/fast_launch.cpp): torch tensors -> skr_plan_launch in one call."


def build_fast(force: bool = False) -> Path | None:
    """Compile ``_fast.so`` in-tree with g++ against the installed torch (ATen + pybind11).  Host glue only: when it
    cannot be built here the Python layer runs the same sequence itself, so a failure is reported, not raised."""
    source = CSRC / "fast_launch.cpp"
    header = ROOT / "include" / "skrample_b200.h"  # the module copies skr_philox tables: the struct layout is part of it
    if not force and FAST.exists() and FAST.stat().st_mtime >= max(source.stat().st_mtime, header.stat().st_mtime):
        return FAST
    try:
        import sysconfig

        import torch
        from torch.utils import cpp_extension

        includes = [*cpp_extension.include_paths(), sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
        libdir = str(Path(torch.__file__).resolve().parent / "lib")
        cmd = [
            shutil.which("g++") or "g++",
            "-O2",
            "-std=c++17",
            "-shared",
            "-fPIC",
            "-fvisibility=hidden",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
            "-DTORCH_EXTENSION_NAME=_fast",
            "-DTORCH_API_INCLUDE_EXTENSION_H",
            *(f"-I{path}" for path in includes),
            str(source),
            "-o",
            str(FAST),
            f"-L{libdir}",
            f"-Wl,-rpath,{libdir}",
            "-lc10",
            "-lc10_cuda",
            "-ltorch_cpu",
            "-ltorch_cuda",
            "-ltorch",
            "-ltorch_python",
        ]
        done = subprocess.run(cmd, capture_output=True, text=True)
        if done.returncode != 0:
            sys.stderr.write(done.stderr[-4000:])
            sys.stderr.write("\nskrample_b200: _fast.so was not built (the Python hit path is used instead)\n")
            return None
        return FAST
    except Exception as error:  # noqa: BLE001 - optional host-side accelerator
        sys.stderr.write(f"skrample_b200: _fast.so was not built: {error}\n")
        return None


def _stale() -> bool:
    if not LIB.exists():
        return True
    built = LIB.stat().st_mtime
    deps = [*CSRC.glob("*.cu"), *CSRC.glob("*.cuh"), *(ROOT / "include").glob("*.h"), Path(__file__)]
    return any(d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple[str, ...] = (), out: Path | None = None) -> Path:
    """Compile the library: the translation units in parallel, then one link.  ``defines`` / ``out`` build an
    experiment variant next to the product library (load it with ``SKRAMPLE_B200_LIB=<path>``); the default call
    builds the product."""
    target = out if out is not None else LIB
    if out is None:
        build_fast(force)
    if out is None and not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = CSRC / "build" / (target.stem if out is not None else "product")
    objdir.mkdir(parents=True, exist_ok=True)
    common = [*NVCC_FLAGS, *(f"-D{d}" for d in defines), *(["-Xptxas", "-v"] if verbose else [])]

    shared = [*CSRC.glob("*.cuh"), *(ROOT / "include").glob("*.h"), Path(__file__)]
    newest_shared = max(d.stat().st_mtime for d in shared)

    def compile_unit(unit: tuple[str, tuple[str, ...], str]) -> subprocess.CompletedProcess[str]:
        source, unit_defines, obj = unit
        made = objdir / obj
        if not force and not verbose and made.exists() and made.stat().st_mtime > max(newest_shared, (CSRC / source).stat().st_mtime):
            return subprocess.CompletedProcess([], 0, "", "")  # neither the unit nor a header changed since it was compiled
        flags = list(common)
        if source == "noise_kernels.cu":
            # noise values carry no bit-level contract beyond "every kernel draws the same normal for the same key"
            # (philox.cuh spells its roundings out), so interpolation / scaling arithmetic may contract to FMA here
            flags[flags.index("-fmad=false")] = "-fmad=true"
        cmd = [nvcc, *flags, *(f"-D{d}" for d in unit_defines), "-c", str(CSRC / source), "-o", str(objdir / obj)]
        return subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_unit, UNITS))
    log = "".join(r.stdout + r.stderr for r in results)
    if any(r.returncode != 0 for r in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed compiling libskrample_b200.so")
    link = subprocess.run(
        [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", str(target), *(str(objdir / u[2]) for u in UNITS)],
        capture_output=True,
        text=True,
    )
    if link.returncode != 0:
        sys.stderr.write(log + link.stdout + link.stderr)
        raise RuntimeError("nvcc failed linking libskrample_b200.so")
    if verbose:
        sys.stderr.write(log)
    return target


if __name__ == "__main__":
    _defines = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    _out = next((Path(a[6:]).resolve() for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=_defines, out=_out))
