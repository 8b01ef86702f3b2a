#!/bin/bash
# N=8 and N=4 lines with the final code (one knot per CTA in the fused exchange kernel)
O=gpurun_out/r2; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_c3_n8_57.json 2> $O/bench_c3_n8_57.err
echo "N=8 rc=$?"; tail -2 $O/bench_c3_n8_57.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_c3_n4_57.json 2> $O/bench_c3_n4_57.err
echo "N=4 rc=$?"
python - <<PY
import json
for n in ("bench_c3_n8_57","bench_c3_n4_57"):
    try:
        d=json.load(open("$O/"+n+".json")); print(n, 'N', d['n_gpus'], round(d['ms_per_step']*1e3,2), f"{d['value']:.4g}", 'e2e', round(d['e2e']['value']), d['detail'].get('exchange_kernel_ms_without_barrier'))
    except Exception as e: print(n, 'ERR', e)
PY
