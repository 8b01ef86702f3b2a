#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/trace_u8p.py > $O/trace_u8p4.txt 2>&1
PB2_NO_UNIT=1 timeout 300 python tools/trace_u8p.py > $O/trace_u8p4_nounit.txt 2>&1
PB2_NO_UNIT=1 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8p4_nounit.json 2> $O/bench_c3_u8p4_nounit.err
python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_c3_u8p4.json 2> $O/bench_c3_u8p4.err
tail -40 $O/trace_u8p4.txt
