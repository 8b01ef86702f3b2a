#!/bin/bash
# batch axis, dense_blocks, C program, dmma launch bounds: full GPU suite + ensemble throughput + C2/C4 re-check
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_34.log 2>&1
timeout 600 python tools/bench_batch.py --members 16 --configs 1,2,4,6 > $O/batch_34.jsonl 2> $O/batch_34.err
for c in 2 4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_34.json 2> $O/bench_c${c}_34.err; done
tail -15 $O/pytest_34.log; cat $O/batch_34.jsonl; tail -3 $O/batch_34.err; head -c 700 $O/bench_c2_34.json; echo; head -c 700 $O/bench_c4_34.json
