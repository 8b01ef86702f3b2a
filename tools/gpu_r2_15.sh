#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -x -q  ) > $O/pytest_18.log 2>&1
( time PB2_U8Q_NS=4 python -m pytest tests -m gpu -x -q -k "single_round or two_cta" ) > $O/pytest_18b.log 2>&1
tail -n 5 $O/pytest_18.log; tail -n 5 $O/pytest_18b.log
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
run v3_q2 PB2_U8Q_NS=2
run v3_q4 PB2_U8Q_NS=4
run v3_q2_nw PB2_U8Q_NS=2 PB2_NOWAIT=1
run v3_q2_again PB2_U8Q_NS=2
PB2_U8Q_NS=2 PB2_U8Q=2 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_v3_q2.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_c5_v3_q2.json')); print('c5 v2 q2', d['ms_per_step']*1e3, d['roofline']['frac'])"
PB2_U8Q_NS=2 timeout 300 python tools/trace_u8q.py > $O/trace_u8q_v3_ns2.txt 2>&1
tail -13 $O/trace_u8q_v3_ns2.txt
