python -m pytest tests/test_gpu_abi.py -x -q -k "slab" 2>&1 | tail -25
python -m pytest tests/test_gpu_abi.py -x -q -k "not slab" 2>&1 | tail -3
