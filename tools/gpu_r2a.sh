# round-2 first GPU pass: tests, default bench, kernel variants (steady-state per-kernel times)
TAG=${1:-r2a}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -15 gpurun_out/test_gpu_$TAG.log
B="python bench.py --scenario-cache /tmp/scn --no-cpu-baseline --no-e2e --gpu-scenarios --steps 50 --preroll-steps 600"
timeout 300 $B > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-300
for v in lg8 lg32 lg16e1 lg16e4 lg8e1 nav6 nav8; do
  AUV_B200_LIB=$PWD/gym_auv_b200/variants/lib_$v.so timeout 300 $B > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_${TAG}_$v.json')); s=d['steady_state']; print('$v', round(d['ms_per_step'],4), 'steady', round(s['ms_per_step'],4), s['kernel_ms'])
except Exception as e: print('$v failed', e)
"
done
python -c "
import json
d=json.load(open('gpurun_out/bench_$TAG.json')); s=d['steady_state']; print('default', round(d['ms_per_step'],4), 'steady', round(s['ms_per_step'],4), s['kernel_ms'])
"
