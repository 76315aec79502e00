( time python -m pytest tests -m gpu -x -q --durations=8 ) 2>&1 | tail -16
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag head_s 2>&1 | tail -3
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --opt v3_tail=0 --tag head_s_tail0 2>&1 | tail -3
python tools/bench_vmult.py --refine 7 --nb 1 --variants 0 --reps 10 --kernels apply cheb_step_own_dinv --tag head_s 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; tail -c 2600 gpurun_out/bench_s.json; tail -3 gpurun_out/bench_s.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaling-reference > gpurun_out/plain_s.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/launches_s.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaling-reference --profile > gpurun_out/ncu_s.log 2>&1
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/plain_s2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_v3 -s 6 -c 2 -o gpurun_out/prof_r02b_cheb_apply python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/ncu_s2.log 2>&1
ls -la gpurun_out/prof_r02b_cheb_apply.ncu-rep
