#!/bin/bash
# concurrent NLP iterate with SMs reserved for the non-Hessian callbacks
O=gpurun_out/r2; mkdir -p $O
for r in 23 19 23; do
PB2_BENCH_HESS_RESERVE=$r python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_61_$r.json 2> $O/bench_c3_61_$r.err
python - <<PY
import json
d=json.load(open("$O/bench_c3_61_$r.json")); it=d['objective']['nlp_iterate']; print('reserve $r:', round(d['ms_per_step']*1e3,3), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', round(it['ms_per_iterate']*1e3,1), 'concurrent', round(it['ms_per_iterate_concurrent']*1e3,1))
PY
done
