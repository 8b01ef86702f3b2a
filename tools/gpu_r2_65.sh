#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_c3_n2_65.json 2> $O/bench_c3_n2_65.err
echo "rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_c3_n2_65.json")); print('N=2', round(d['ms_per_step']*1e3,2), f"{d['value']:.4g}", 'e2e', round(d['e2e']['value']))
PY
timeout 200 python -m pytest tests -m gpu -q -k "two_gpus or sharded" 2>&1 | tail -2
