TAG=${1:-r2h}
B="python bench.py --scenario-cache /tmp/scn --no-cpu-baseline --no-e2e"
show() { python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_${TAG}_$1.json')); print('$1', round(d['ms_per_step'],4), 'after_reset', round(d['after_reset']['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items() if not isinstance(v,list)})
except Exception as e: print('$1 failed', e)
"; }
timeout 300 $B > gpurun_out/bench_${TAG}_base.json 2> gpurun_out/bench_${TAG}_base.err; show base
for v in $2; do
  AUV_B200_LIB=$PWD/gym_auv_b200/variants/lib_$v.so timeout 300 $B > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err; show $v
done
