#!/bin/bash
# batch grid shape (one wave of persistent CTAs over all members), chunked host-pointer pipeline
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_35.log 2>&1
timeout 600 python tools/bench_batch.py --members 16 --configs 1,2,4,6 > $O/batch_35.jsonl 2> $O/batch_35.err
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_35.json 2> $O/bench_c3_35.err
PB2_E2E_PIPE=0 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_35_nopipe.json 2> $O/bench_c3_35_nopipe.err
PB2_D2H_CHUNKS=16 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_35_16ch.json 2> $O/bench_c3_35_16ch.err
python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_35.json 2> $O/bench_c5_35.err
tail -8 $O/pytest_35.log; cat $O/batch_35.jsonl; tail -3 $O/batch_35.err
python - <<PY
import json
for n in ("bench_c3_35","bench_c3_35_nopipe","bench_c3_35_16ch","bench_c5_35"):
    try:
        d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), (d.get('objective') or {}).get('nlp_iterate'))
    except Exception as e: print(n, 'ERR', e)
PY
