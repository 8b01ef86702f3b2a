#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > $O/memcheck_58.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|Error" $O/memcheck_58.log | head -20; tail -12 $O/memcheck_58.log
