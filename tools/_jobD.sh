python -m pytest tests/test_gpu_abi.py tests/test_gpu_multirank.py -x -q 2>&1 | tail -8
K="--kernels apply cheb_step_own_dinv residual"
B="python tools/bench_vmult.py --variants 0 --reps 20 $K"
$B --refine 6 --tag lockstep --nb 1 --opt v3_schedule=1
$B --refine 6 --tag lockstep --nb 2 --opt v3_schedule=1
for c in 2 4 8; do $B --refine 6 --tag rr$c --nb 1 --opt v3_chunk=$c; $B --refine 6 --tag rr$c --nb 2 --opt v3_chunk=$c; done
for c in 1 2 4 8; do $B --refine 5 --tag rr$c --nb 2 --opt v3_chunk=$c; done
$B --refine 5 --tag old --nb 2 --opt v3_schedule=3
for c in 1 2 4; do $B --refine 4 --tag rr$c --nb 2 --opt v3_chunk=$c; done
$B --refine 4 --tag old --nb 2 --opt v3_schedule=3
for c in 1 2 4; do $B --refine 3 --tag rr$c --nb 2 --opt v3_chunk=$c; done
$B --refine 3 --tag old --nb 2 --opt v3_schedule=3
$B --refine 7 --tag rr8 --nb 1 --reps 10
