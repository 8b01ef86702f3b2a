#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
N=${1:-4}
run() { n=$1; shift; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/bench_$n.json 2> $O/bench_$n.err; python - <<PY
import json
try:
    d=json.load(open("$O/bench_$n.json")); print("$n step_us", round(d['ms_per_step']*1e3,2), "value", round(d['value']), d['scaling'], "knots/gpu", d['config']['knots_per_gpu'], "e2e", round(d['e2e']['value']))
except Exception as e: print("$n ERR", e)
PY
}
run strong_c5_n$N --config 5 --strong
if [ "$2" == "both" ]; then run weak_c3_n$N; fi
