timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefetched or graphed or static or linear" 2>&1 | tail -6
