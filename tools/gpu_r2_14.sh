#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
PB2_U8Q_NS=2 timeout 300 python tools/trace_u8q.py > $O/trace_u8q_ns2.txt 2>&1
PB2_U8Q_NS=4 timeout 300 python tools/trace_u8q.py > $O/trace_u8q_ns4.txt 2>&1
PB2_U8Q_NS=2 PB2_NOWAIT=1 PB2_U8Q_SPACE=3000 timeout 300 python tools/trace_u8q.py > $O/trace_u8q_ns2_nw_s3000.txt 2>&1
head -60 $O/trace_u8q_ns2.txt
