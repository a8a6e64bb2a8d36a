#!/usr/bin/env python
"""Dump the Butcher coefficients of the reference's named tableaux to a data file.

Run in the build container only (needs /root/reference or $SKRAMPLE_REF):

    python tools/dump_tableaux.py

The coefficients are published constants (Butcher, Fehlberg, Cash-Karp, Dormand-Prince, Ruuth's SSP methods,
Feagin, Stepanov, ...); what is stored is their exact float64 value (``float.hex``) as the reference evaluates
them, so ``skrample_b200.sampling.tableaux`` reproduces them bit for bit without carrying the reference's
source files.  Output: skrample_b200/sampling/tableaux/coefficients.json
"""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

sys.path.insert(0, os.environ.get("SKRAMPLE_REF", "/root/reference"))

from skrample.sampling import tableaux  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "skrample_b200" / "sampling" / "tableaux" / "coefficients.json"


def encode(tab) -> dict:
    item = {
        "stages": [[float(c).hex(), [float(x).hex() for x in a]] for c, a in tab.stages],
        "weights": [float(x).hex() for x in tab.weights],
    }
    if hasattr(tab, "error_weights"):
        item["error_weights"] = [float(x).hex() for x in tab.error_weights]
    return item


def main() -> None:
    families = {}
    for name in ("RK1", "RK2", "RK3", "RK4", "RKZ", "RKE2", "RKE3", "RKE5", "SSP", "WSO", "Shanks1965"):
        enum = getattr(tableaux, name)
        families[name] = {member.name: encode(member.value) for member in enum}
    OUT.write_text(json.dumps(families, separators=(",", ":")))
    count = sum(len(v) for v in families.values())
    print(f"{count} tableaux in {len(families)} families -> {OUT} ({OUT.stat().st_size // 1024} KiB)")


if __name__ == "__main__":
    main()
