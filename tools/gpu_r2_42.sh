#!/bin/bash
# N=2 e2e A/B of the chunked pipeline on one box; u8h step restructure (single GPU part)
O=gpurun_out/r2; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -x -k "hessian or u8 or full_size" ) > $O/pytest_42.log 2>&1
tail -4 $O/pytest_42.log
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_42.json 2> $O/bench_c3_42.err
python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_42.json 2> $O/bench_c5_42.err
for pipe in 1 0 1 0; do
PB2_E2E_PIPE=$pipe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$pipe bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_c3_n2_42_p$pipe.json 2> $O/bench_c3_n2_42_p$pipe.err
python - <<PY
import json
d=json.load(open("$O/bench_c3_n2_42_p$pipe.json")); print('N=2 pipe=$pipe', round(d['ms_per_step']*1e3,2), f"{d['value']:.3g}", 'e2e', round(d['e2e']['value']))
PY
done
python - <<PY
import json
for n in ("bench_c3_42","bench_c5_42"):
    d=json.load(open("$O/"+n+".json")); print(n, round(d['ms_per_step']*1e3,3), 'e2e', round(d['e2e']['value']), 'hess', round(d['hessian']['ms_per_callback']*1e3,2), 'iter', d['objective']['nlp_iterate']['ms_per_iterate'])
PY
