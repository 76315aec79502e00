python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -c 1500 gpurun_out/bench_p.json; tail -3 gpurun_out/bench_p.err
( time python bench.py --impl reference --steps 4 --warmup 1 --cpu-budget 90 ) > gpurun_out/bench_ref_p.json 2> gpurun_out/bench_ref_p.err; tail -c 900 gpurun_out/bench_ref_p.json; tail -4 gpurun_out/bench_ref_p.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaling-reference > gpurun_out/plain_p.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/launches_p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaling-reference --profile > gpurun_out/ncu_p.log 2>&1
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/plain_p2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_v3 -s 6 -c 4 -o gpurun_out/prof_r02_cheb_apply python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/ncu_p2.log 2>&1
ls -la gpurun_out/prof_r02_cheb_apply.ncu-rep
