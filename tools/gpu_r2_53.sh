#!/bin/bash
# tensor-core general kernels extended to b <= 24 (NT = 3)
O=gpurun_out/r2; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_53.log 2>&1
grep -E "FAILED|passed|failed" $O/pytest_53.log | tail -8
python tools/bench_shapes.py > $O/shapes_53.txt 2>&1; cat $O/shapes_53.txt | tail -5
