python -m pytest tests/test_gpu_abi.py -x -q -k "slab" 2>&1 | grep -E "AssertionError|passed|failed" | head -5
