#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > $O/sharded_check_n2.log 2>&1; tail -5 $O/sharded_check_n2.log
python -m pytest tests -m gpu -x -q -k "two_gpus" > $O/pytest_23.log 2>&1; tail -3 $O/pytest_23.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err; python - <<PY
import json
d=json.load(open("$O/bench_n2.json")); print("N=2", d['ms_per_step']*1e3, d['value'], d['e2e']['value'], d['detail']['collective'][:80])
PY
PB2_U8Q_NO_PEERS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_old.json 2> $O/bench_n2_old.err; python - <<PY
import json
d=json.load(open("$O/bench_n2_old.json")); print("N=2 old kernel", d['ms_per_step']*1e3, d['value'])
PY
