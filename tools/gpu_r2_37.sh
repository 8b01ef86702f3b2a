#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q -x -k "pipeline or occupancy or full_size or synthetic or golden or c_program" ) > $O/pytest_37.log 2>&1
{
PB2_E2E_TIMELINE=1 python tools/e2e_timeline.py
PB2_E2E_TIMELINE=1 PB2_E2E_H2D=2 python tools/e2e_timeline.py
PB2_E2E_TIMELINE=1 PB2_D2H_CHUNKS=6 python tools/e2e_timeline.py
PB2_E2E_TIMELINE=1 python tools/e2e_timeline.py 8000
PB2_E2E_TIMELINE=1 PB2_HOST_NT=0 python tools/e2e_timeline.py 8000
} > $O/e2e_timeline_37.txt 2>&1
tail -5 $O/pytest_37.log; cat $O/e2e_timeline_37.txt
