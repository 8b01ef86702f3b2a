#!/bin/bash
# A/B of the coupling-sum forms in knot_dmmah (compile-time variants), same box
O=gpurun_out/r2; mkdir -p $O
for v in default chain0 chain1 default chain0 chain1; do
  if [ $v = default ]; then unset PB2_LIB; else export PB2_LIB=$PWD/piccolo.jl_b200/libpb2_$v.so; fi
  echo "== $v"
  timeout 300 python tools/bench_batch.py --members 16 --configs 2,4 --iters 100 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], 'hess batch', d['hess']['graph_batch_us'], 'per-member', d['hess']['graph_per_member_us'])
"
done
