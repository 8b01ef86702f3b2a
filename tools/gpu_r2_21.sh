#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_21.log 2>&1
tail -n 25 $O/pytest_21.log
