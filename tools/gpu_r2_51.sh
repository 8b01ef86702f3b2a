#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -q -x -k "hessian or synthetic or golden or batch or shapes or substeps" ) > $O/pytest_51.log 2>&1
tail -2 $O/pytest_51.log
for c in 1 2 4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_51.json 2> $O/bench_c${c}_51.err; done
python - <<PY
import json
for n in ("bench_c1_51","bench_c2_51","bench_c4_51"):
    d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), 'iso', round(d['roofline'].get('isolated_launch_us'),2), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', ((d.get('objective') or {}).get('nlp_iterate') or {}).get('ms_per_iterate'))
PY
timeout 300 python tools/bench_batch.py --members 16 --configs 2,4 --iters 100 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], 'resjac batch', d['resjac']['graph_batch_us'], 'hess batch', d['hess']['graph_batch_us'])
"
