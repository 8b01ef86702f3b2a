#!/bin/bash
# C3 / C5 Hessian: knot_u8h against the general tensor-core kernel with 20 warps per knot
O=gpurun_out/r2; mkdir -p $O
PB2_HESS_DMMAH=1 timeout 600 python -m pytest tests -m gpu -q -x -k "hessian or synthetic or full_size_properties" > $O/pytest_44.log 2>&1
tail -3 $O/pytest_44.log
for v in 0 1 0 1; do
PB2_HESS_DMMAH=$v python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_44_$v.json 2> $O/bench_c3_44_$v.err
python - <<PY
import json
d=json.load(open("$O/bench_c3_44_$v.json")); print('C3 dmmah=$v', round(d['ms_per_step']*1e3,3), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), d['hessian']['kernel'][:18], 'iter', d['objective']['nlp_iterate']['ms_per_iterate'])
PY
done
for v in 0 1; do
PB2_HESS_DMMAH=$v python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_44_$v.json 2> $O/bench_c5_44_$v.err
python - <<PY
import json
d=json.load(open("$O/bench_c5_44_$v.json")); print('C5 dmmah=$v', round(d['ms_per_step']*1e3,3), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), d['hessian']['kernel'][:18])
PY
done
