#!/bin/bash
# where the host-pointer callback's time goes: timeline stamps for the upload modes / chunk counts
O=gpurun_out/r2; mkdir -p $O
{
for mode in 0 1 2; do PB2_E2E_TIMELINE=1 PB2_E2E_H2D=$mode python tools/e2e_timeline.py; done
for ch in 4 6 12; do PB2_E2E_TIMELINE=1 PB2_E2E_H2D=2 PB2_D2H_CHUNKS=$ch python tools/e2e_timeline.py; done
PB2_E2E_TIMELINE=1 PB2_E2E_PIPE=0 python tools/e2e_timeline.py
PB2_E2E_TIMELINE=1 PB2_HOST_NT=1 PB2_E2E_H2D=2 python tools/e2e_timeline.py
PB2_E2E_TIMELINE=1 PB2_HOST_THREADS=4 PB2_E2E_H2D=2 python tools/e2e_timeline.py
PB2_E2E_TIMELINE=1 PB2_E2E_H2D=2 python tools/e2e_timeline.py 8000
} > $O/e2e_timeline_36.txt 2>&1
cat $O/e2e_timeline_36.txt; nproc
