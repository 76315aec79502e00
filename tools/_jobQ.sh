python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "residual or fast_path or v3 or slab" 2>&1 | tail -3
python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual --tag prepass2 2>&1 | tail -3
python tools/bench_vmult.py --refine 7 --nb 1 --variants 0 --reps 10 --kernels cheb_step_own_dinv --tag prepass2 2>&1 | tail -1
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio --clock-control none -k regex:k_v3 -s 6 -c 1 --csv --log-file gpurun_out/ncu_q_inst.csv python tools/bench_vmult.py --refine 6 --nb 2 --variants 0 --reps 1 --kernels cheb_step_own_dinv apply > gpurun_out/ncu_q.log 2>&1
cut -d, -f13- gpurun_out/ncu_q_inst.csv | tail -6
