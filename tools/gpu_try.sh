timeout 900 python -m pytest tests/test_gpu_spheres.py tests/test_gpu_parity.py -m gpu -x -q -k "sphere or augment or vote or voting or two_graph" 2>&1 | tail -15
