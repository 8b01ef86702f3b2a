#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_50.log 2>&1
grep -E "FAILED|passed|failed" $O/pytest_50.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_50.log 2>&1; echo "smoke rc=$?"
