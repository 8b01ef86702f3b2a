#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for c in 3 2; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_59.json 2> $O/bench_c${c}_59.err; tail -2 $O/bench_c${c}_59.err; done
python - <<PY
import json
for n in ("bench_c3_59","bench_c2_59"):
    try:
        d=json.load(open("$O/"+n+".json")); it=d['objective']['nlp_iterate']; print(n, round(d['ms_per_step']*1e3,3), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', round(it['ms_per_iterate']*1e3,1), 'concurrent', round(it['ms_per_iterate_concurrent']*1e3,1))
    except Exception as e: print(n, 'ERR', e)
PY
