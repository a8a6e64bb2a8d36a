"""SASS evidence of the shipped library: per kernel, instruction count and the mnemonics that prove TMA staging
(UBLKCP: 1-D bulk copy) and mbarrier pipelines (SYNCS), written to profiles/<round>_sass_extract.txt (no GPU needed).

    python tools/sass_extract.py r02
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "skrample_b200" / "csrc" / "libskrample_b200.so"
round_name = sys.argv[1] if len(sys.argv) > 1 else "r02"
text = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
rows = []
total = collections.Counter()
for body in re.split(r"\n\s+Function : ", text)[1:]:
    name = body.split("\n", 1)[0].strip()
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
    short = collections.Counter(o.split(".")[0] for o in ops)
    rows.append((name, len(ops), short["UBLKCP"], short["SYNCS"], short["MUFU"], short["FFMA"], short["FMUL"] + short["FADD"], short["UTCHMMA"] + short["HMMA"] + short["UTCMMA"]))
    total.update({k: short[k] for k in ("UBLKCP", "SYNCS")})
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
arch = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True).stdout
out = [
    "# SASS evidence extracted from skrample_b200/csrc/libskrample_b200.so with `cuobjdump -sass` (tools/sass_extract.py)",
    f"# embedded cubins: {', '.join(sorted(set(re.findall(r'sm_[0-9]+a?', arch))))}",
    f"# kernels: {len(rows)}; UBLKCP (1-D TMA bulk copy) instructions: {total['UBLKCP']}; SYNCS (mbarrier) instructions: {total['SYNCS']}; "
    f"tensor-core instructions: {sum(r[7] for r in rows)} (none by design: the path is elementwise)",
    "",
    f"{'kernel':<150} {'instr':>6} {'UBLKCP':>7} {'SYNCS':>6} {'MUFU':>5} {'FFMA':>5} {'FMUL+FADD':>9}",
]
for (_, count, ublk, syncs, mufu, ffma, fp, _), pretty in sorted(zip(rows, names), key=lambda pair: -pair[0][1]):
    pretty = re.sub(r"\(.*$", "", pretty.replace("skr::", ""))
    out.append(f"{pretty[:148]:<150} {count:>6} {ublk:>7} {syncs:>6} {mufu:>5} {ffma:>5} {fp:>9}")
(ROOT / "profiles" / f"{round_name}_sass_extract.txt").write_text("\n".join(out) + "\n")
print("\n".join(out[:10]))
