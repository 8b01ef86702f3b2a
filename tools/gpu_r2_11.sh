#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -k "single_round or two_cta or u8 or full_size or compact or sharded_integrator_single" ) > $O/pytest_11.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8q.json 2> $O/bench_c3_u8q.err
PB2_U8Q=0 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8q_off.json 2> $O/bench_c3_u8q_off.err
PB2_BENCH_EARLY_Z=0 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8q_noez.json 2> $O/bench_c3_u8q_noez.err
PB2_U8Q=2 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_u8q.json 2> $O/bench_c5_u8q.err
python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_u8.json 2> $O/bench_c5_u8.err
tail -4 $O/pytest_11.log
for f in bench_c3_u8q bench_c3_u8q_off bench_c3_u8q_noez bench_c5_u8q bench_c5_u8; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), d['roofline']['isolated_launch_us'], d['e2e']['value'])
except Exception as e: print("$f ERR", e)
PY
done
