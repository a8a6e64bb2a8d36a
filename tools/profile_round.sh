#!/bin/bash
# Capture the round's measurement artefacts on the GPU box (run under gpurun from the repo root):
#   tools/profile_round.sh r02
# Writes gpurun_out/<round>_*.json (bench lines), <round>_launches.csv (ncu launch list of the bench command) and
# <round>_*.ncu-rep (one `ncu --set full` capture per kernel).  Each ncu run follows a plain run of the same
# command that exited 0.  tools/summarize_profiles.py turns them into the tracked summaries under profiles/.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err || echo "bench failed"
python bench.py --impl reference > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err || echo "reference arm failed"
python tools/e2e_breakdown.py > $O/${R}_e2e_breakdown.txt 2>&1 || echo "breakdown failed"
python tools/hit_breakdown.py > $O/${R}_hit_breakdown.txt 2>&1 || echo "hit breakdown failed"
tools/bin/step_floor > $O/${R}_step_floor.txt 2>&1 || echo "floor probe failed"
python tools/per_step.py unipc3_sde_flux_bf16 > $O/${R}_per_step_unipc3_flux_bf16.txt 2>&1 || echo "per-step failed"
python tools/per_step.py unipc3_sde_flux_bf16 --contracted > $O/${R}_per_step_unipc3_flux_bf16_contracted.txt 2>&1 || echo "per-step failed"
python tools/per_step.py unipc3_sde_flux64_bf16 > $O/${R}_per_step_unipc3_flux64_bf16.txt 2>&1 || echo "per-step failed"

BENCH="python bench.py --quick --steps 400 --warmup 50 --no-cpu-baseline --streams 1 --inflight 1"
$BENCH > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"block_kernel|step_kernel|fill_kernel" -c 3000 --csv --log-file $O/${R}_launches.csv $BENCH > $O/${R}_ncu_launch.log 2>&1

# the same step reading supplied noise tensors (bench.py: kernel_only / roofline_supplied_noise)
KONLY="python tools/ab_bench.py unipc3_sde_sdxl_bf16"
$KONLY > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"block_kernel|step_kernel|fill_kernel" -c 3000 --csv --log-file $O/${R}_launches_kernel_only.csv $KONLY > $O/${R}_ncu_launch_kernel_only.log 2>&1

capture() {  # name, kernel regex, kernel skip count, command...
    local name=$1 kernel=$2 skip=$3
    shift 3
    "$@" > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on \
        -k regex:$kernel -s $skip -c 1 -f -o $O/${R}_$name "$@" > $O/${R}_ncu_$name.log 2>&1
}
P="python tools/prof_one.py"
capture unipc3_sde_bf16_2x16x128x128 block_kernel 6 $P --sampler unipc3 --dtype bf16 --batch 2 --steps 8
capture unipc3_sde_bf16_2x16x128x128_philox block_kernel 6 $P --sampler unipc3 --dtype bf16 --batch 2 --steps 8 --lazy-noise
capture unipc3_sde_bf16_16x16x128x128 block_kernel 6 $P --sampler unipc3 --dtype bf16 --batch 16 --steps 8
capture unipc3_sde_bf16_16x16x128x128_contracted block_kernel 6 env SKR_ARITH=contracted $P --sampler unipc3 --dtype bf16 --batch 16 --steps 8
capture euler_sde_f32_256x16x128x128 block_kernel 3 $P --sampler euler --dtype f32 --batch 256 --steps 6
capture euler_sde_f32_256x16x128x128_philox block_kernel 3 $P --sampler euler --dtype f32 --batch 256 --steps 6 --lazy-noise
capture adams9_sde_bf16_19x16x128x128 block_kernel 10 $P --sampler adams9 --dtype bf16 --batch 19 --steps 12
capture interpreter_unipc3_sde_f32_16x16x128x128 step_kernel 6 env SKR_FORCE_INTERP=1 $P --sampler unipc3 --dtype f32 --batch 16 --steps 8
# noise kernels on one 16x21x90x160 video latent (the first timed shape of tools/noise_bench.py)
N="python tools/noise_bench.py"
capture noise_fill_f32_16x21x90x160 fill_kernel 25 $N Random
capture noise_pyramid_resident_f32_16x21x90x160 pyramid_resident 25 $N Pyramid
capture noise_pyramid_compose_f32_16x21x90x160 pyramid_compose 25 env SKR_NO_RESIDENT_PYRAMID=1 $N Pyramid
capture noise_pyramid_widen_f32_16x21x90x160 levels_widen 25 $N Pyramid
capture noise_pyramid_levels_f32_16x21x90x160 levels_fill 25 $N Pyramid
capture noise_scale_f32_16x21x90x160 scale_kernel 25 env SKR_NO_RESIDENT_PYRAMID=1 $N Pyramid
capture noise_colored_shape_16x21x90x160 colored_shape 25 $N Colored
capture noise_moments_f32_16x21x90x160 moments_kernel 25 $N Colored
capture noise_brownian_f32_16x21x90x160 brownian_kernel 30 $N Brownian
# summarise here: only gpurun_out/ travels back and the reports are far larger than its 64 MiB limit
SKR_PROFILES_OUT=$O/${R}_summary python tools/summarize_profiles.py $R > $O/${R}_summarize.log 2>&1
rm -f $O/${R}_*.ncu-rep
du -sh $O; ls $O/${R}_summary
