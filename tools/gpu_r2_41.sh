#!/bin/bash
# two GPUs: the sharded tests and the N=2 bench line with the final code
O=gpurun_out/r2; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -k "two_gpus or sharded or exchange" ) > $O/pytest_41.log 2>&1
tail -5 $O/pytest_41.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_c3_n2_41.json 2> $O/bench_c3_n2_41.err
tail -3 $O/bench_c3_n2_41.err
python - <<PY
import json
d=json.load(open("$O/bench_c3_n2_41.json")); print('N=2', round(d['ms_per_step']*1e3,2), f"{d['value']:.3g}", 'e2e', round(d['e2e']['value']))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2_41.json 2> $O/bench_ref_n2_41.err
head -c 400 $O/bench_ref_n2_41.json
