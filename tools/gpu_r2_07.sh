#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -k "single_round or u8 or full_size or compact or sharded_integrator_single" ) > $O/pytest_09.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8p7.json 2> $O/bench_c3_u8p7.err
PB2_NO_UNIT=1 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8p7_nounit.json 2> $O/bench_c3_u8p7_nounit.err
timeout 300 python tools/trace_u8p.py > $O/trace_u8p7.txt 2>&1
tail -4 $O/pytest_09.log; head -c 1300 $O/bench_c3_u8p7.json; echo; grep -A16 "block 0 smid" $O/trace_u8p7.txt | head -18;  grep "^block 0 w" $O/trace_u8p7.txt
