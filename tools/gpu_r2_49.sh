#!/bin/bash
# u8q one knot per CTA as the default: full suite, N=2 sharded tests + bench, final C3 / C5 lines
O=gpurun_out/r2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_49.log 2>&1
grep -E "FAILED|passed|failed" $O/pytest_49.log | tail -8
python bench.py --steps 20 --warmup 5 > $O/bench_c3_49.json 2> $O/bench_c3_49.err
python bench.py --config 5 --steps 20 --warmup 5 --no-cpu > $O/bench_c5_49.json 2> $O/bench_c5_49.err
PB2_BENCH_EARLY_Z=0 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_49_noopt.json 2> /dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_c3_n2_49.json 2> $O/bench_c3_n2_49.err
tail -2 $O/bench_c3_n2_49.err
python - <<PY
import json
for n in ("bench_c3_49","bench_c5_49","bench_c3_49_noopt","bench_c3_n2_49"):
    try:
        d=json.load(open("$O/"+n+".json")); print(n, 'N', d['n_gpus'], round(d['ms_per_step']*1e3,3), round(d['roofline']['frac'],4), 'iso', d['roofline'].get('isolated_launch_us'), 'e2e', round(d['e2e']['value']), f"{d['value']:.4g}", d['detail'].get('early_z'), d['detail'].get('pipelined'))
    except Exception as e: print(n, 'ERR', e)
PY
ncu --set full --clock-control none --import-source on -k regex:knot_u8q -s 6 -c 2 -o $O/prof_u8q_ns1 python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_u8q_ns1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file $O/launches_u8q_ns1.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_launches_u8q_ns1.log 2>&1
ls -la $O/prof_u8q_ns1.ncu-rep
