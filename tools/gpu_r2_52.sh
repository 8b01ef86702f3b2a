#!/bin/bash
# knot_u8h with separate barriers for the forward and the adjoint tiles
O=gpurun_out/r2; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -q -x -k "hessian or synthetic or full_size_properties or u8" ) > $O/pytest_52.log 2>&1
tail -2 $O/pytest_52.log
for c in 3 5; do python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_52.json 2> $O/bench_c${c}_52.err; done
python - <<PY
import json
for n in ("bench_c3_52","bench_c5_52"):
    d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', ((d.get('objective') or {}).get('nlp_iterate') or {}).get('ms_per_iterate'))
PY
