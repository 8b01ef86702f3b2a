#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_45.log 2>&1; echo "smoke rc=$?"; tail -8 $O/smoke_45.log
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_45.log 2>&1
tail -5 $O/pytest_45.log
python bench.py > $O/bench_default_45.json 2> $O/bench_default_45.err; echo "bench rc=$?"; head -c 600 $O/bench_default_45.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_45.json 2> $O/bench_ref_45.err; echo "ref rc=$?"; head -c 300 $O/bench_ref_45.json
