#!/bin/bash
# round-2 final single-GPU evidence: tests, bench lines C1-C5, ensemble throughput, e2e timeline, ncu of the general Hessian kernel
O=gpurun_out/r2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_40.log 2>&1
tail -6 $O/pytest_40.log
python bench.py --steps 20 --warmup 5 > $O/bench_c3_40.json 2> $O/bench_c3_40.err
for c in 1 2 4 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_40.json 2> $O/bench_c${c}_40.err; done
timeout 600 python tools/bench_batch.py --members 16 --configs 1,2,4,6 > $O/batch_40.jsonl 2> $O/batch_40.err
PB2_E2E_TIMELINE=1 python tools/e2e_timeline.py > $O/e2e_timeline_40.txt 2>&1
PB2_E2E_TIMELINE=1 PB2_E2E_PIPE=0 python tools/e2e_timeline.py >> $O/e2e_timeline_40.txt 2>&1
PB2_E2E_TIMELINE=1 python tools/e2e_timeline.py 8000 >> $O/e2e_timeline_40.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:knot_dmmah -s 4 -c 1 -o $O/prof_dmmah_c2 python bench.py --config 2 --steps 4 --warmup 3 --no-cpu > $O/ncu_dmmah_c2.log 2>&1
ncu --set full --clock-control none -k regex:knot_dmmah -s 4 -c 1 -o $O/prof_dmmah_c4 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu > $O/ncu_dmmah_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file $O/launches_c2_40.csv python bench.py --config 2 --steps 2 --warmup 1 --no-cpu > $O/ncu_launches_c2.log 2>&1
python - <<PY
import json
for n in ("bench_c1_40","bench_c2_40","bench_c3_40","bench_c4_40","bench_c5_40"):
    try:
        d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline'].get('isolated_launch_us'), 'e2e', round(d['e2e']['value']), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', ((d.get('objective') or {}).get('nlp_iterate') or {}).get('ms_per_iterate'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(n, 'ERR', e)
PY
cat $O/batch_40.jsonl; cat $O/e2e_timeline_40.txt; ls -la $O/*.ncu-rep
