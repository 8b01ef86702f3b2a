#!/bin/bash
# general tensor-core Hessian (knot_dmmah), e2e small-first-chunk
O=gpurun_out/r2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -x -k "hessian or synthetic or golden or batch or shapes" ) > $O/pytest_38a.log 2>&1
tail -15 $O/pytest_38a.log
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_38.log 2>&1
tail -8 $O/pytest_38.log
timeout 600 python tools/bench_batch.py --members 16 --configs 1,2,4,6 > $O/batch_38.jsonl 2> $O/batch_38.err
cat $O/batch_38.jsonl; tail -3 $O/batch_38.err
for c in 2 4 3; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_38.json 2> $O/bench_c${c}_38.err; done
python - <<PY
import json
for n in ("bench_c2_38","bench_c4_38","bench_c3_38"):
    try:
        d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), 'e2e', round(d['e2e']['value']), 'hess', d['hessian'], (d.get('objective') or {}).get('nlp_iterate',{}).get('ms_per_iterate'))
    except Exception as e: print(n, 'ERR', e)
PY
PB2_E2E_TIMELINE=1 python tools/e2e_timeline.py 2>&1 | tail -4
