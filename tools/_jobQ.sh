python -m pytest tests/test_gpu_abi.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag tail1 2>&1 | tail -3
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --opt v3_tail=0 --tag tail0 2>&1 | tail -3
python tools/bench_vmult.py --refine 6 --nb 1 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv --tag tail1_nb1 2>&1 | tail -2
python tools/bench_vmult.py --refine 7 --nb 1 --variants 0 --reps 10 --kernels apply cheb_step_own_dinv --tag tail1 2>&1 | tail -2
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-scaling-reference 2>&1 | tail -1 | cut -c1-2600
