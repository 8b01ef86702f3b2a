#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_03.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8p.json 2> $O/bench_c3_u8p.err
PB2_BENCH_EARLY_Z=0 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8p_noez.json 2> $O/bench_c3_u8p_noez.err
timeout 300 python tools/trace_u8p.py > $O/trace_u8p.txt 2>&1
tail -5 $O/pytest_03.log; head -c 1500 $O/bench_c3_u8p.json; echo; tail -3 $O/bench_c3_u8p.err; head -30 $O/trace_u8p.txt
