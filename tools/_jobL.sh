python -m pytest tests/test_gpu_abi.py -x -q -k "slab" 2>&1 | grep -E "AssertionError|passed|failed|Error" | head -8
python -m pytest tests/test_gpu_multirank.py -x -q -k "slabs" 2>&1 | tail -30
