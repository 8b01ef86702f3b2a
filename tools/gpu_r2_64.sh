#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
ncu --set full --clock-control none -k regex:knot_u8h -s 2 -c 1 -o $O/prof_u8h_split python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_u8h_split.log 2>&1
ls -la $O/prof_u8h_split.ncu-rep
