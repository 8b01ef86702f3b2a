#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_20.log 2>&1
tail -n 5 $O/pytest_20.log
for c in 3 5 1 2 4; do
  python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_r20_c$c.json 2> $O/bench_r20_c$c.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_r20_c$c.json")); print("C$c", round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline']['isolated_launch_us'], 'e2e', round(d['e2e']['value']), 'hess', (d.get('hessian') or {}).get('ms_per_callback'))
except Exception as e: print("C$c ERR", e)
PY
done
PB2_BENCH_PIPELINED=0 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_r20_c3_nopipe.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_r20_c3_nopipe.json')); print('C3 early-z only', d['ms_per_step']*1e3, d['roofline']['frac'])"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_r20_ref.json 2>/dev/null; head -c 400 $O/bench_r20_ref.json
