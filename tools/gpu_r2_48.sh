#!/bin/bash
# u8q with one knot per CTA (64 threads, 8 CTAs per SM) against the default two
O=gpurun_out/r2; mkdir -p $O
PB2_U8Q_NS=1 timeout 600 python -m pytest tests -m gpu -q -x -k "single_round or full_size or synthetic or pipelined or two_cta" > $O/pytest_48.log 2>&1
tail -3 $O/pytest_48.log
for ns in 2 1 2 1; do
PB2_U8Q_NS=$ns python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_48_$ns.json 2> $O/bench_c3_48_$ns.err
python - <<PY
import json
d=json.load(open("$O/bench_c3_48_$ns.json")); print('C3 ns=$ns', round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline']['isolated_launch_us'])
PY
done
for ns in 2 1; do
PB2_U8Q_NS=$ns python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_48_$ns.json 2> $O/bench_c5_48_$ns.err
python - <<PY
import json
d=json.load(open("$O/bench_c5_48_$ns.json")); print('C5 ns=$ns', round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4))
PY
done
