#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
ncu --set full --clock-control none -k regex:knot_dmmah -s 2 -c 1 -o $O/prof_dmmah_b18 python tools/bench_shapes.py > $O/ncu_dmmah_b18.log 2>&1
ncu --set full --clock-control none -k regex:knot_dmmaq -s 2 -c 1 -o $O/prof_dmmaq_b18 python tools/bench_shapes.py > $O/ncu_dmmaq_b18.log 2>&1
ls -la $O/prof_dmma*_b18.ncu-rep
