for cfg in 3 5; do for d in 0 3; do
st=1000; [ $cfg = 5 ] && st=200
PB2_LIB=libpiccolo_b200_trace.so PB2_DRY=$d timeout 200 python bench.py --no-cpu --config $cfg --steps $st 2>gpurun_out/dry.err | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
if not t: print('no output'); sys.exit(0)
d=json.loads(t[-1]); print('cfg',$cfg,'dry',$d, d['ms_per_step'], d['roofline']['kernel_ms'])"
done; done; tail -3 gpurun_out/dry.err
