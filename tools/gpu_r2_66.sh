#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_c3_66.json 2> $O/bench_c3_66.err; echo "rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_66.json 2> $O/bench_ref_66.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_c3_66.json")); it=d['objective']['nlp_iterate']
print('C3', round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline']['isolated_launch_us'], 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', round(it['ms_per_iterate']*1e3,1), round(it['ms_per_iterate_concurrent']*1e3,1), d['clocks'])
r=json.load(open("$O/bench_ref_66.json")); print('ref', round(r['value']), r['steps'], round(r['ms_per_step'],1))
PY
