for f in 1 0; do echo "== WEASAL_RS_FAST=$f"; WEASAL_RS_FAST=$f timeout 600 python tools/sweep.py 100000 1000000 2>&1 | grep batch_query | python -c "
import sys,json
for l in sys.stdin: d=json.loads(l); print(d['N'], round(d['ms'],3), round(d['Mqueries_per_s']))"
WEASAL_RS_FAST=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-extra-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernels']['rs_search']['ms_per_step'])"; done
