#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -q ) > $O/pytest_63.log 2>&1
grep -E "FAILED|passed|failed" $O/pytest_63.log | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_63.json 2> $O/bench_c3_63.err
python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_63.json 2> $O/bench_c5_63.err
python - <<PY
import json
for n in ("bench_c3_63","bench_c5_63"):
    d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline']['isolated_launch_us'], 'e2e', round(d['e2e']['value']))
PY
