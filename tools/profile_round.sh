#!/bin/bash
# Capture the round's measurement artefacts on the GPU box (run under gpurun from the repo root):
#   tools/profile_round.sh r01
# Writes gpurun_out/<round>_*.json (bench lines), <round>_launches.csv (ncu launch list of the bench command) and
# <round>_*.ncu-rep (one `ncu --set full` capture per top kernel).  Each ncu run follows a plain run of the same
# command that exited 0.  tools/summarize_profiles.py turns them into the tracked summaries under profiles/.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python bench.py --sweep > $O/${R}_bench.json 2> $O/${R}_bench.err || echo "bench failed"
python bench.py --impl reference > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err || echo "reference arm failed"
python bench.py --fused-noise --sweep --no-cpu-baseline > $O/${R}_bench_fused_noise.json 2>> $O/${R}_bench.err || echo "fused bench failed"

BENCH="python bench.py --steps 400 --warmup 50 --no-cpu-baseline"
$BENCH > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"block_kernel|step_kernel|fill_kernel" -c 3000 --csv --log-file $O/${R}_launches.csv $BENCH > $O/${R}_ncu_launch.log 2>&1

capture() {  # name, kernel skip count, prof_one arguments...
    local name=$1 skip=$2
    shift 2
    python tools/prof_one.py "$@" > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on \
        -k regex:block_kernel -s $skip -c 1 -f -o $O/${R}_$name python tools/prof_one.py "$@" > $O/${R}_ncu_$name.log 2>&1
}
capture euler_sde_f32_256x16x128x128 3 --sampler euler --dtype f32 --batch 256 --steps 6
capture unipc3_sde_bf16_16x16x128x128 6 --sampler unipc3 --dtype bf16 --batch 16 --steps 8
capture unipc3_sde_bf16_2x16x128x128 6 --sampler unipc3 --dtype bf16 --batch 2 --steps 8
capture adams9_sde_bf16_19x16x128x128 10 --sampler adams9 --dtype bf16 --batch 19 --steps 12
# the Brownian interval kernel on one 16x21x90x160 video latent (the first timed shape of tools/noise_bench.py)
python tools/noise_bench.py Brownian > $O/${R}_noise_brownian.txt 2>&1 && ncu --set full --clock-control none --import-source on \
    -k regex:brownian_kernel -s 30 -c 1 -f -o $O/${R}_brownian_f32_16x21x90x160 python tools/noise_bench.py Brownian > $O/${R}_ncu_brownian.log 2>&1
ls -la $O/${R}_*
