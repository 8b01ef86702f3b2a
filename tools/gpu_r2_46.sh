#!/bin/bash
# knot_dmmaq (one small CTA per knot for the general residual+Jacobian) against knot_dmma
O=gpurun_out/r2; mkdir -p $O
( time PB2_DMMAQ=1 timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_46.log 2>&1
grep -E "FAILED|passed|failed" $O/pytest_46.log | tail -12
for v in 0 1 0 1; do
  for c in 1 2 4; do
    PB2_DMMAQ=$v timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_46_$v.json 2> $O/bench_c${c}_46_$v.err
  done
  python - <<PY
import json
out=[]
for c in (1,2,4):
    d=json.load(open("$O/bench_c%d_46_$v.json" % c)); out.append("C%d %.2f us (iso %.1f) iter %s" % (c, d['ms_per_step']*1e3, d['roofline']['isolated_launch_us'], ((d.get('objective') or {}).get('nlp_iterate') or {}).get('ms_per_iterate')))
print('dmmaq=$v', ' | '.join(out))
PY
done
for v in 0 1; do echo "== batch dmmaq=$v"; PB2_DMMAQ=$v timeout 300 python tools/bench_batch.py --members 16 --configs 1,2,4,6 --iters 100 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], 'resjac batch', d['resjac']['graph_batch_us'], 'per-member', d['resjac']['graph_per_member_us'], 'evals/s', d['resjac']['batch_knot_evals_per_s'])
"; done
