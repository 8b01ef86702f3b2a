#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -q -x -k "hessian or shapes" ) > $O/pytest_56.log 2>&1
tail -2 $O/pytest_56.log
python tools/bench_shapes.py > $O/shapes_56.txt 2>&1; cat $O/shapes_56.txt | tail -5
