#!/bin/bash
# final evidence with knot_dmmaq as the default general kernel
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_47.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_47.log 2>&1
tail -5 $O/pytest_47.log
for c in 1 2 4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_47.json 2> $O/bench_c${c}_47.err; done
timeout 600 python tools/bench_batch.py --members 16 --configs 1,2,4,6 > $O/batch_47.jsonl 2> $O/batch_47.err
ncu --set full --clock-control none --import-source on -k regex:knot_dmmaq -s 6 -c 1 -o $O/prof_dmmaq_c2 python bench.py --config 2 --steps 4 --warmup 3 --no-cpu > $O/ncu_dmmaq_c2.log 2>&1
ncu --set full --clock-control none -k regex:knot_dmmaq -s 6 -c 1 -o $O/prof_dmmaq_c4 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu > $O/ncu_dmmaq_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file $O/launches_c2_47.csv python bench.py --config 2 --steps 2 --warmup 1 --no-cpu > $O/ncu_launches_c2_47.log 2>&1
python - <<PY
import json
for n in ("bench_c1_47","bench_c2_47","bench_c4_47"):
    d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline'].get('isolated_launch_us'), 'e2e', round(d['e2e']['value']), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', ((d.get('objective') or {}).get('nlp_iterate') or {}).get('ms_per_iterate'))
PY
cat $O/batch_47.jsonl; ls -la $O/prof_dmmaq*.ncu-rep
