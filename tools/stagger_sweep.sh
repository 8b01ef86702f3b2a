# sweep of the start delay (cycles) of a group that owns one knot less (PB2_STAGGER), C3 bench
for s in 0 2500 5000 8000; do
  PB2_STAGGER=$s timeout 100 python bench.py --no-cpu --steps 400 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger', $s, 'ms', d['ms_per_step'])"
done
