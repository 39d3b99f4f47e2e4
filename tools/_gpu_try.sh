timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kpconv or linear" 2>&1 | tail -4
timeout 300 python tools/kpconv_kernel_times.py 2>&1 | tail -6
timeout 600 python tools/sweep.py 1000000 2>&1 | grep kpconv
