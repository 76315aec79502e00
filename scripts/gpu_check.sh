#!/bin/bash
# The measurement sequence behind profiles/ (run on a B200 box from the repository root, e.g. through
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'):
#   1. the GPU parity tests, 2. the stage-vmult / smoother micro-benchmark, 3. the bench line, 4. the ncu launch list of the
#   bench command, 5. one `ncu --set full` capture of the dominant kernel.  Every ncu pass runs only after the same command
#   has exited 0 without ncu; numbers printed under ncu are never bench values.
# Outputs go to gpurun_out/ (scratch); copy what should be kept to profiles/ (tools/summarize_launches.py turns the
# launch list into the per-kernel share table).
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 ) 2>&1 | tail -16
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag check 2>&1 | tail -3
python tools/bench_vmult.py --refine 7 --nb 1 --variants 0 --reps 10 --kernels apply cheb_step_own_dinv --tag check 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaling-reference > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaling-reference --profile > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv "bench.py --steps 2 --warmup 3 (IRK q=2, r=6)" 2>/dev/null | head -24
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_v3 -s 6 -c 2 -o gpurun_out/prof_cheb_apply \
    python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_cheb_apply.ncu-rep
