#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
N=${1:-4}
run() { n=$1; shift; env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_$n.json 2> $O/bench_$n.err; python - <<PY
import json
try:
    d=json.load(open("$O/bench_$n.json")); print("$n step_us", round(d['ms_per_step']*1e3,2), "value", round(d['value']), "kernel_us", round(d['roofline']['kernel_ms']*1e3,2), "xk_us", d['detail'].get('exchange_kernel_ms_without_barrier'), "e2e", round(d['e2e']['value']), d['detail']['pipelined'])
except Exception as e: print("$n ERR", e)
PY
}
run n${N}_rot PB2_BENCH_DIAG=1
