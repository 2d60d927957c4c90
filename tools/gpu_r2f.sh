# tests + default bench + A/B of obs_nz and two lidar variants
TAG=${1:-r2f}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -12 gpurun_out/test_gpu_$TAG.log
B="python bench.py --scenario-cache /tmp/scn --no-cpu-baseline"
timeout 600 $B > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-1200
tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
show() { python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_${TAG}_$1.json')); print('$1', round(d['ms_per_step'],4), 'after_reset', round(d['after_reset']['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items() if not isinstance(v,list)})
except Exception as e: print('$1 failed', e)
"; }
AUV_B200_NO_OBS_NZ=1 timeout 300 $B --no-e2e > gpurun_out/bench_${TAG}_nonz.json 2> gpurun_out/bench_${TAG}_nonz.err; show nonz
for v in le16 lmb5; do
  AUV_B200_LIB=$PWD/gym_auv_b200/variants/lib_$v.so timeout 300 $B --no-e2e > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err; show $v
  AUV_B200_NO_OBS_NZ=1 AUV_B200_LIB=$PWD/gym_auv_b200/variants/lib_$v.so timeout 300 $B --no-e2e > gpurun_out/bench_${TAG}_${v}nonz.json 2> gpurun_out/bench_${TAG}_${v}nonz.err; show ${v}nonz
done
for c in 1 2 8; do
  timeout 300 $B --no-e2e --chunks $c > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; show c$c
done
