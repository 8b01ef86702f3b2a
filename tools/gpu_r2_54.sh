#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
python tools/bench_shapes.py > $O/shapes_54.txt 2>&1; cat $O/shapes_54.txt | tail -5
