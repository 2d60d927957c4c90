TAG=${1:-r2g}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -6 gpurun_out/test_gpu_$TAG.log
B="python bench.py --scenario-cache /tmp/scn --no-cpu-baseline"
timeout 600 $B > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-1500
show() { python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_${TAG}_$1.json')); print('$1', round(d['ms_per_step'],4), 'after_reset', round(d['after_reset']['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items() if not isinstance(v,list)}, d['run'].get('records_per_env_step'))
except Exception as e: print('$1 failed', e)
"; }
for v in le16 lmb5; do
  AUV_B200_LIB=$PWD/gym_auv_b200/variants/lib_$v.so timeout 300 $B --no-e2e > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err; show $v
done
timeout 600 python bench.py --workload land --envs 131072 --n-moving 0 --n-static 0 --n-polygons 512 --n-paths 256 --no-cpu-baseline --no-e2e --steps 30 --preroll-steps 300 > gpurun_out/bench_${TAG}_land.json 2> gpurun_out/bench_${TAG}_land.err; show land; tail -2 gpurun_out/bench_${TAG}_land.err | cut -c1-300
timeout 900 python tools/sweep.py --out gpurun_out/sweep_$TAG.jsonl > gpurun_out/sweep_$TAG.log 2>&1; tail -30 gpurun_out/sweep_$TAG.log | cut -c1-200
