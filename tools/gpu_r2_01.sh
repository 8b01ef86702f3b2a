#!/bin/bash
# round-2 GPU call 1: baseline tests (with the accumulator pre-load in knot_dmma), the single-round kernel behind PB2_U8S=1
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_base.log 2>&1
( time PB2_U8S=1 timeout 600 python -m pytest tests -m gpu -q -k "u8 or full_size or synthetic" ) > $O/pytest_u8s.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_base.json 2> $O/bench_c3_base.err
PB2_U8S=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8s.json 2> $O/bench_c3_u8s.err
for c in 1 2 4 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > $O/bench_c${c}_base.json 2> $O/bench_c${c}_base.err; done
tail -3 $O/pytest_base.log $O/pytest_u8s.log
cat $O/bench_c3_base.json | head -c 600; echo; cat $O/bench_c3_u8s.json | head -c 600
