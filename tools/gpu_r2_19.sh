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
for sp in 1500 2500 3500 5000; do
  run v3_s${sp} PB2_U8Q_SPACE=$sp PB2_U8Q_PRO=3500
  run v3_s${sp}_nw PB2_U8Q_SPACE=$sp PB2_U8Q_PRO=3500 PB2_NOWAIT=1
done
run v3_s2500_p2500_nw PB2_U8Q_SPACE=2500 PB2_U8Q_PRO=2500 PB2_NOWAIT=1
run v3_q4_s5000_nw PB2_U8Q_NS=4 PB2_U8Q_SPACE=5000 PB2_U8Q_PRO=3500 PB2_NOWAIT=1
run v3_q4_s7000_nw PB2_U8Q_NS=4 PB2_U8Q_SPACE=7000 PB2_U8Q_PRO=3500 PB2_NOWAIT=1
PB2_U8Q_SPACE=2500 PB2_U8Q_PRO=3500 PB2_NOWAIT=1 timeout 300 python tools/trace_u8q.py > $O/trace_u8q_v3_s2500_nw.txt 2>&1
sed -n 1,30p $O/trace_u8q_v3_s2500_nw.txt
