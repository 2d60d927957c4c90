# tests, default bench (sustained headline incl. e2e), kernel variants.  Usage: bash tools/gpu_r2b.sh TAG "variant list"
TAG=${1:-r2b}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -12 gpurun_out/test_gpu_$TAG.log
B="python bench.py --scenario-cache /tmp/scn --no-cpu-baseline"
timeout 600 $B > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-700
tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
for v in $2; do
  AUV_B200_LIB=$PWD/gym_auv_b200/variants/lib_$v.so timeout 300 $B --no-e2e > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_${TAG}_$v.json')); print('$v', round(d['ms_per_step'],4), 'after_reset', round(d['after_reset']['ms_per_step'],4), d['roofline']['kernel_ms'])
except Exception as e: print('$v failed', e)
"
done
