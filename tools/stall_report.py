"""Top stall sites of one ncu report (run on the GPU box, where the report is; prints a short table).

    python tools/stall_report.py gpurun_out/x.ncu-rep [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys

report = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
header = next((r for r in rows if "Source" in r and "Instructions Executed" in r), None)
if header is None:
    print("no source page")
    sys.exit(1)
body = rows[rows.index(header) + 1 :]
i_src = header.index("Source")
i_exec = header.index("Instructions Executed")
i_samp = next((header.index(h) for h in header if h.startswith("Warp Stall Sampling (All")), None)
stall_cols = [(h, header.index(h)) for h in header if h.startswith("stall_")]
total = 0
lines = []
by_op = collections.Counter()
by_reason = collections.Counter()
for n, r in enumerate(body):
    try:
        samples = int(r[i_samp]) if i_samp is not None else 0
        executed = int(r[i_exec])
    except (ValueError, IndexError):
        continue
    total += samples
    lines.append((samples, executed, n, r[i_src].strip()))
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src].strip())
    by_op[m.group(2).split(".")[0] if m else "?"] += samples
    for name, col in stall_cols:
        try:
            by_reason[name] += int(r[col])
        except (ValueError, IndexError):
            pass
print(f"total stall samples {total}")
print("by opcode:", ", ".join(f"{k} {v / max(total, 1):.1%}" for k, v in by_op.most_common(12)))
if by_reason:
    print("by reason:", ", ".join(f"{k} {v}" for k, v in by_reason.most_common(10)))
print("top lines (samples, executed, line, sass):")
for samples, executed, n, src in sorted(lines, reverse=True)[:top]:
    print(f"{samples:7d} {executed:9d} {n:6d}  {src[:110]}")
