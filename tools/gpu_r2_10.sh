#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:knot_u8p -s 6 -c 2 -o $O/prof_u8p python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_u8p.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 200 --csv --log-file $O/launches_u8p.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_launches.log 2>&1
ls -la $O/*.ncu-rep; tail -3 $O/ncu_u8p.log
