#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
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
run nw_q4 PB2_NOWAIT=1 PB2_U8Q_NS=4
run nw_q2 PB2_NOWAIT=1 PB2_U8Q_NS=2
for sp in 1000 2000 3000; do run nw_q4_s$sp PB2_NOWAIT=1 PB2_U8Q_NS=4 PB2_U8Q_SPACE=$sp; run nw_q2_s$sp PB2_NOWAIT=1 PB2_U8Q_NS=2 PB2_U8Q_SPACE=$sp; done
run nw_q2_c5 PB2_NOWAIT=1 PB2_U8Q_NS=2 PB2_U8Q=2 PB2_BENCH_CONFIG=5
PB2_NOWAIT=1 PB2_U8Q_NS=2 PB2_U8Q=2 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_nw_q2.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_c5_nw_q2.json')); print('c5 nw q2', d['ms_per_step']*1e3, d['roofline']['frac'])"
PB2_NOWAIT=1 PB2_U8Q_NS=2 PB2_U8Q=2 PB2_U8Q_SPACE=2000 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_nw_q2s.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_c5_nw_q2s.json')); print('c5 nw q2 space2000', d['ms_per_step']*1e3, d['roofline']['frac'])"
PB2_U8Q_NS=2 PB2_U8Q=2 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_q2.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_c5_q2.json')); print('c5 q2', d['ms_per_step']*1e3, d['roofline']['frac'])"
