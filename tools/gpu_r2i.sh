TAG=${1:-r2i}
B="python bench.py --scenario-cache /tmp/scn --no-cpu-baseline --no-e2e"
show() { python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_${TAG}_$1.json')); print('$1', round(d['ms_per_step'],4), 'after_reset', round(d['after_reset']['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items() if not isinstance(v,list)})
except Exception as e: print('$1 failed', e)
"; }
for v in split cap5 cap4e16 cap5s; do
 for c in "2 2" "4 2" "8 2" "8 4" "16 4"; do
  set -- $c
  AUV_B200_LIB=$PWD/gym_auv_b200/variants/lib_$v.so timeout 300 $B --chunks $1 --chunk-streams $2 > gpurun_out/bench_${TAG}_${v}_c$1s$2.json 2> gpurun_out/bench_${TAG}_${v}_c$1s$2.err; show ${v}_c$1s$2
 done
done
