nvidia-smi -L
python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 4 --warmup 2 2> gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.json
tail -c 1200 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
SPIRK_SYNC_TIMERS=1 $TR -m dealii_spirk_b200.launch --dim 3 json/bench_spirk_q8_r6.json > gpurun_out/launch_n2_q8_sync_timers.log 2>&1
tail -8 gpurun_out/launch_n2_q8_sync_timers.log
