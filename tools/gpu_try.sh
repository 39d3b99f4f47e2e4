timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subsample" 2>&1 | tail -6
timeout 600 python tools/sweep.py 1000000 10000000 2>&1 | grep grid_subsample | python -c "
import sys,json
for l in sys.stdin: d=json.loads(l); print(d['order'], d['N'], round(d['ms'],3), round(d['Mpts_per_s']))"
