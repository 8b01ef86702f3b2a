#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:knot_u8q -s 6 -c 2 -o $O/prof_u8q python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_u8q.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file $O/launches_u8q.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_launches_u8q.log 2>&1
ncu --set full --clock-control none -k regex:knot_u8h -s 2 -c 1 -o $O/prof_u8h python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_u8h.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/bench_r33_c3_full.json 2> $O/bench_r33_c3_full.err
PB2_HOST_NT=1 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_r33_c3_nt.json 2>/dev/null
python - <<PY
import json
for n in ("bench_r33_c3_full","bench_r33_c3_nt"):
    d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline']['isolated_launch_us'], 'e2e', round(d['e2e']['value']), d.get('cpu_baseline',{}).get('value'))
PY
ls -la $O/*.ncu-rep
