#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_02.log 2>&1
PB2_U8S=1 timeout 300 python tools/trace_u8p.py > $O/trace_u8s.txt 2>&1
tail -5 $O/pytest_02.log; head -40 $O/trace_u8s.txt
