"""Build ``csrc/libskrample_b200.so`` in-tree with nvcc for sm_100a.

    python -m skrample_b200.build [--force] [-v] [-DNAME[=VALUE] ... --out=path/to/variant.so]

nvcc cross-compiles without a GPU.  The shared library stays in the source tree
(git-ignored) so it travels with the repository snapshot to the GPU box.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
ROOT = CSRC.parent.parent
LIB = CSRC / "libskrample_b200.so"
SOURCES = ["step_kernel.cu", "noise_kernels.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-fmad=false",  # every product and sum is individually rounded, like the reference's separate ATen ops
    "--shared",
    "-Xcompiler",
    "-fPIC",
]


def _stale() -> bool:
    if not LIB.exists():
        return True
    built = LIB.stat().st_mtime
    deps = [*CSRC.glob("*.cu"), *CSRC.glob("*.cuh"), *(ROOT / "include").glob("*.h"), Path(__file__)]
    return any(d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple[str, ...] = (), out: Path | None = None) -> Path:
    """Compile the library.  ``defines`` / ``out`` build an experiment variant next to the product library
    (load it with ``SKRAMPLE_B200_LIB=<path>``); the default call builds the product."""
    target = out if out is not None else LIB
    if out is None and not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    sources = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    flags = [*NVCC_FLAGS, *(f"-D{d}" for d in defines), *(["-Xptxas", "-v"] if verbose else [])]
    cmd = [nvcc, *flags, "-o", str(target), *sources]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libskrample_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    return target


if __name__ == "__main__":
    _defines = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    _out = next((Path(a[6:]).resolve() for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=_defines, out=_out))
