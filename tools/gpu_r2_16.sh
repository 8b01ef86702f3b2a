#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
PB2_U8Q_NS=2 timeout 300 python tools/trace_u8q.py > $O/trace_u8q_v2b_ns2.txt 2>&1
tail -13 $O/trace_u8q_v2b_ns2.txt
