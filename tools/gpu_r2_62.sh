#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -q ) > $O/pytest_62.log 2>&1
grep -E "FAILED|passed|failed" $O/pytest_62.log | tail -5
