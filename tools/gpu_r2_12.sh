#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time PB2_U8Q_NS=2 python -m pytest tests -m gpu -x -q -k "single_round or two_cta" ) > $O/pytest_12a.log 2>&1
( time PB2_U8Q_NS=2 PB2_U8Q_SPACE=3000 python -m pytest tests -m gpu -x -q -k "single_round or two_cta" ) > $O/pytest_12b.log 2>&1
tail -3 $O/pytest_12a.log $O/pytest_12b.log
run() { # name, env...
  n=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_$n.json 2> $O/bench_$n.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$n.json")); print("$n", round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), d['roofline']['isolated_launch_us'])
except Exception as e: print("$n ERR", e)
PY
}
run q4 PB2_U8Q_NS=4
run q2 PB2_U8Q_NS=2
for sp in 1500 3000 4500 6000; do run q4_s$sp PB2_U8Q_NS=4 PB2_U8Q_SPACE=$sp; run q2_s$sp PB2_U8Q_NS=2 PB2_U8Q_SPACE=$sp; done
run q2_s3000_p2000 PB2_U8Q_NS=2 PB2_U8Q_SPACE=3000 PB2_U8Q_PRO=2000
run q2_s3000_p4000 PB2_U8Q_NS=2 PB2_U8Q_SPACE=3000 PB2_U8Q_PRO=4000
