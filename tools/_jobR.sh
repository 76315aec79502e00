python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/plain_r.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_v3 -s 6 -c 2 -o gpurun_out/prof_r02b_cheb_apply python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/ncu_r.log 2>&1
ls -la gpurun_out/prof_r02b_cheb_apply.ncu-rep
