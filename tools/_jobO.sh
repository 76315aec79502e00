python -m pytest tests/test_gpu_abi.py -x -q 2>&1 | tail -3
K="--kernels apply cheb_step_own_dinv residual"
B="python tools/bench_vmult.py --variants 0 --reps 20 $K"
for pf in 1 0; do
$B --refine 6 --tag pf$pf --nb 2 --opt v3_prefetch=$pf
$B --refine 6 --tag pf${pf}_c4 --nb 2 --opt v3_prefetch=$pf v3_chunk=4
$B --refine 5 --tag pf$pf --nb 2 --opt v3_prefetch=$pf
$B --refine 4 --tag pf$pf --nb 2 --opt v3_prefetch=$pf
done
$B --refine 7 --tag pf1 --nb 1 --reps 10
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-scaling-reference | tail -c 700
