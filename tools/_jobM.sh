nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multirank.py -x -q -k "four or grid" 2>&1 | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521"
$TR bench.py --gpus 4 --steps 4 --warmup 2 2> gpurun_out/bench_n4.err | tail -1 > gpurun_out/bench_n4_q8.json
python -c "import json; d=json.load(open('gpurun_out/bench_n4_q8.json')); print('N=4 q=8', d['ms_per_step'], d['outer_iterations'], d['e2e']['ms_per_step'], d['error_L2_final'])"
$TR bench.py --gpus 4 --stages 4 --steps 4 --warmup 2 2>> gpurun_out/bench_n4.err | tail -1 > gpurun_out/bench_n4_q4.json
python -c "import json; d=json.load(open('gpurun_out/bench_n4_q4.json')); print('N=4 q=4', d['ms_per_step'], d['outer_iterations'], d['e2e']['ms_per_step'], d['error_L2_final'])"
$TR bench.py --gpus 4 --stages 2 --steps 4 --warmup 2 2>> gpurun_out/bench_n4.err | tail -1 > gpurun_out/bench_n4_q2_slabs.json
python -c "import json; d=json.load(open('gpurun_out/bench_n4_q2_slabs.json')); print('N=4 q=2 (2x2 slabs)', d['ms_per_step'], d['outer_iterations'], d['e2e']['ms_per_step'], d['error_L2_final'], d['config']['parallelism'])"
tail -3 gpurun_out/bench_n4.err
