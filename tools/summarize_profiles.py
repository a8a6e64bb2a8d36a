"""Turn the artefacts of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01
"""
from __future__ import annotations

import collections
import csv
import io
import re
import shutil
import subprocess
import sys
from pathlib import Path

import os

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
# On the GPU box the summaries go under gpurun_out/ (the only directory that travels back; the .ncu-rep files are far
# over its 64 MiB limit and are deleted there after summarising): SKR_PROFILES_OUT=gpurun_out/r02_summary
PROFILES = Path(os.environ.get("SKR_PROFILES_OUT", ROOT / "profiles"))

KEEP = (
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "launch__block_size",
    "launch__grid_size",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg",
    "l1tex__t_sector_hit_rate.pct",
)


def launches(round_name: str, suffix: str = "", command: str = "python bench.py --quick --steps 400 --warmup 50 --no-cpu-baseline --streams 1 --inflight 1") -> None:
    path = OUT / f"{round_name}_launches{suffix}.csv"
    if not path.exists():
        return
    lines = [line for line in path.read_text().splitlines() if line.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    groups: dict[tuple[str, str], list[float]] = collections.defaultdict(list)
    for row in rows:
        if row["Metric Name"] == "gpu__time_duration.sum":
            name = re.sub(r"skr::", "", row["Kernel Name"])
            groups[(name, row["Grid Size"])].append(float(row["Metric Value"].replace(",", "")) / 1e3)
    # DRAM traffic per launch of the dominant step kernel (the default workload's steady-state shape): bench.py
    # reports it as roofline.traffic next to the algorithmic bytes
    traffic: dict[tuple[str, str], list[float]] = collections.defaultdict(list)
    per_id: dict[str, float] = collections.defaultdict(float)
    kernel_of: dict[str, tuple[str, str]] = {}
    for row in rows:
        if row["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1.0)
            per_id[row["ID"]] += float(row["Metric Value"].replace(",", "")) * scale
            kernel_of[row["ID"]] = (re.sub(r"skr::", "", row["Kernel Name"]), row["Grid Size"])
    for launch, nbytes in per_id.items():
        traffic[kernel_of[launch]].append(nbytes)
    total = sum(sum(v) for v in groups.values())
    if traffic:
        top = max((k for k in groups if "block_kernel" in k[0]), key=lambda k: sum(groups[k]), default=None)
        if top is not None and traffic.get(top):
            import json

            (PROFILES / f"{round_name}_traffic{suffix}.json").write_text(
                json.dumps({"kernel": top[0], "grid": top[1], "launches": len(traffic[top]), "dram_bytes_per_launch": sum(traffic[top]) / len(traffic[top]),
                            "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of this kernel in `{round_name}_launches{suffix}.csv` (`{command}`, 17 interleaved latent batches)"}, indent=1) + "\n"
            )
    out = [
        f"# ncu launch list of `{command}` (gpu__time_duration.sum, --clock-control none)",
        '# filter: -k regex:"block_kernel|step_kernel|fill_kernel" -c 3000; per-launch times are cold-cache and serialised: compare SHARES',
        "",
        f"{'kernel':<110} {'grid':<14} {'n':>6} {'avg us':>8} {'share':>6}",
    ]
    for (name, grid), values in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
        short = name if len(name) <= 108 else name[:105] + "..."
        out.append(f"{short:<110} {grid:<14} {len(values):>6} {sum(values) / len(values):>8.2f} {sum(values) / total:>6.3f}")
    (PROFILES / f"{round_name}_launches{suffix}_summary.txt").write_text("\n".join(out) + "\n")
    (PROFILES / f"{round_name}_launches{suffix}_head.csv").write_text("\n".join(path.read_text().splitlines()[:60]) + "\n")


def full_captures(round_name: str) -> None:
    for report in sorted(OUT.glob(f"{round_name}_*.ncu-rep")):
        raw = subprocess.run(["ncu", "-i", str(report), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        header, units, values = rows[0], rows[1], rows[2]
        out = [f"# ncu --set full --clock-control none --import-source on, one launch, report gpurun_out/{report.name}"]
        for name, unit, value in zip(header, units, values):
            if name == "Kernel Name":
                out.append(f"Kernel Name = {value}")
            elif name in KEEP or "issue_stalled" in name and name.endswith("per_issue_active.ratio"):
                out.append(f"{name} = {value} {unit}".rstrip())
        source = subprocess.run(["ncu", "-i", str(report), "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        src_rows = list(csv.reader(io.StringIO(source)))
        if len(src_rows) > 2 and "Instructions Executed" in src_rows[1]:
            hdr = src_rows[1]
            i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
            by_op: collections.Counter[str] = collections.Counter()
            for row in src_rows[2:]:
                try:
                    count = int(row[i_exec])
                except (ValueError, IndexError):
                    continue
                match = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", row[i_src].strip())
                by_op[match.group(2) if match else "?"] += count
            total = sum(by_op.values())
            out.append(f"# executed warp instructions by opcode (source page), total {total}")
            out.extend(f"  {op:<8} {n:>12} {n / total:6.1%}" for op, n in by_op.most_common(16))
        (PROFILES / f"{report.stem.replace(round_name + '_', round_name + '_ncu_')}.txt").write_text("\n".join(out) + "\n")


def bench_lines(round_name: str) -> None:
    for name in ("bench", "bench_reference", "e2e_breakdown", "hit_breakdown", "step_floor", "per_step_unipc3_flux_bf16", "per_step_unipc3_flux_bf16_contracted", "per_step_unipc3_flux64_bf16"):
        for suffix in (".json", ".txt"):
            src = OUT / f"{round_name}_{name}{suffix}"
            if src.exists() and src.stat().st_size and src.resolve().parent != PROFILES.resolve():
                shutil.copy(src, PROFILES / src.name)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "r01"
    PROFILES.mkdir(parents=True, exist_ok=True)
    bench_lines(which)
    launches(which)
    launches(which, "_kernel_only", "python tools/ab_bench.py unipc3_sde_sdxl_bf16")
    full_captures(which)
    print("\n".join(sorted(p.name for p in PROFILES.glob(f"{which}_*"))))
