( time python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "four" --durations=5 ) 2>&1 | tail -12 | cut -c1-300
