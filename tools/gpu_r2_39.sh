#!/bin/bash
# dmmah: 2 CTAs/SM for 8x8, hoisted base term + split coupling chains (also in knot_dmma)
O=gpurun_out/r2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_39.log 2>&1
tail -6 $O/pytest_39.log
timeout 600 python tools/bench_batch.py --members 16 --configs 1,2,4,6 > $O/batch_39.jsonl 2> $O/batch_39.err
cat $O/batch_39.jsonl; tail -3 $O/batch_39.err
for c in 1 2 4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_39.json 2> $O/bench_c${c}_39.err; done
python - <<PY
import json
for n in ("bench_c1_39","bench_c2_39","bench_c4_39"):
    try:
        d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), 'e2e', round(d['e2e']['value']), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), d['hessian']['kernel'][:20], (d.get('objective') or {}).get('nlp_iterate'))
    except Exception as e: print(n, 'ERR', e)
PY
