nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
run() { tag=$1; shift; $TR bench.py --gpus 8 "$@" 2>> gpurun_out/bench_n8.err | tail -1 > gpurun_out/bench_n8_$tag.json; python -c "import json; d=json.load(open('gpurun_out/bench_n8_$tag.json')); print('$tag', round(d['ms_per_step'],2), d['outer_iterations'], round(d['e2e']['ms_per_step'],2), d['error_L2_final'], d['config']['parallelism'])"; }
run q8 --steps 4 --warmup 2
run q4_4x2slabs --stages 4 --steps 4 --warmup 2
SPIRK_SLAB_MIN_CELLS=8 run q4_4x2slabs_min8 --stages 4 --steps 4 --warmup 2
run q4_4x2slabs_r7 --stages 4 --refine 7 --steps 3 --warmup 2
run q8_r7_config5 --refine 7 --steps 3 --warmup 2
run q2_2x4slabs --stages 2 --steps 4 --warmup 2
tail -3 gpurun_out/bench_n8.err
