#!/usr/bin/env python
"""Put the UNMODIFIED reference package where bench.py's reference arm looks for it: ``baseline/_ref/`` (git-ignored, but
shipped to the GPU box by gpurun, where /root/reference does not exist).

First the prescribed offline install (``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
baseline/_ref <reference>``).  The reference's build backend (``uv_build``, pyproject.toml:11-12) is not in the wheelhouse,
so that fails here; the package is pure Python (``module-root = ""``, a wheel would contain exactly the ``skrample/``
directory), so the fallback installs what the wheel would: a byte-for-byte copy of ``skrample/`` plus a marker file
recording how it got there.  Nothing under baseline/_ref is tracked or imported by the product.
"""

from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
TARGET = ROOT / "baseline" / "_ref"


def install(reference: Path | None = None, quiet: bool = True) -> str:
    reference = reference or Path(os.environ.get("SKRAMPLE_REF", "/root/reference"))
    if not (reference / "skrample" / "__init__.py").exists():
        return "reference tree not present"
    marker = TARGET / "INSTALLED.json"
    if marker.exists() and (TARGET / "skrample" / "__init__.py").exists():
        return json.loads(marker.read_text())["how"]
    TARGET.mkdir(parents=True, exist_ok=True)
    pip = subprocess.run(
        [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse", "--target", str(TARGET), str(reference)],
        capture_output=True,
        text=True,
    )
    if pip.returncode == 0 and (TARGET / "skrample" / "__init__.py").exists():
        how = "pip install --target baseline/_ref"
    else:
        reason = (pip.stderr.strip().splitlines() or ["pip failed"])[-1]
        if (TARGET / "skrample").exists():
            shutil.rmtree(TARGET / "skrample")
        shutil.copytree(reference / "skrample", TARGET / "skrample", ignore=shutil.ignore_patterns("__pycache__"))
        how = f"copy of the pure-Python package (pip could not build it: {reason})"
    marker.write_text(json.dumps({"how": how, "source": str(reference)}))
    if not quiet:
        print(how)
    return how


if __name__ == "__main__":
    print(install(quiet=True))
