python -m pytest tests/test_gpu_abi.py -x -q 2>&1 | tail -15
B="python tools/bench_vmult.py --refine 6 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual"
$B --tag lockstep --nb 1 --opt v3_schedule=1
$B --tag dyn8 --nb 1
$B --tag dyn4 --nb 1 --opt v3_chunk=4
$B --tag dyn16 --nb 1 --opt v3_chunk=16
$B --tag dyn2 --nb 1 --opt v3_chunk=2
$B --tag dyn8_nb2 --nb 2
$B --tag dyn4_nb2 --nb 2 --opt v3_chunk=4
python tools/bench_vmult.py --refine 5 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag r5_lockstep3 --nb 2 --opt v3_schedule=3
python tools/bench_vmult.py --refine 5 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag r5_dyn8 --nb 2
python tools/bench_vmult.py --refine 5 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag r5_dyn4 --nb 2 --opt v3_chunk=4
python tools/bench_vmult.py --refine 5 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag r5_dyn4_small --nb 2 --opt v3_chunk=4 v3_small_below=64
python tools/bench_vmult.py --refine 4 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag r4_old --nb 2 --opt v3_schedule=3
python tools/bench_vmult.py --refine 4 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag r4_dyn4 --nb 2 --opt v3_chunk=4
python tools/bench_vmult.py --refine 7 --variants 0 --reps 10 --kernels apply cheb_step_own_dinv residual --tag r7_dyn8 --nb 1
