# multi-GPU pass: default bench, config 4 (land, 131072 envs per GPU), reduced config-5 sweep.  Usage: bash tools/gpu_multi.sh TAG NGPU
TAG=${1:-m8}; N=${2:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; grep '^\[bench\]' gpurun_out/bench_${TAG}.err | cut -c1-600
python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}.json')); e=d['e2e']; print('N=$N value', round(d['value']/1e6,1),'M ms', round(d['ms_per_step'],4), 'e2e', round(e['value']/1e6,1), 'M expand_ms', e.get('host_expand_ms_per_step'), 'threads', e.get('host_threads'), 'setup_s', d['run']['setup_s'])
"
timeout 900 $T bench.py --gpus $N --workload land --envs 131072 --n-moving 0 --n-static 0 --n-polygons 512 --n-paths 256 --no-cpu-baseline --no-e2e --steps 30 --preroll-steps 300 > gpurun_out/bench_${TAG}_land.json 2> gpurun_out/bench_${TAG}_land.err
python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}_land.json')); print('land N=$N total envs', d['config']['envs_per_gpu']*$N, 'value', round(d['value']/1e6,1),'M ms', round(d['ms_per_step'],4), d['roofline']['kernel_ms']['nav_pair_min_med_max'], d['run']['records_per_env_step'])
"
tail -2 gpurun_out/bench_${TAG}_land.err | cut -c1-300
timeout 900 $T tools/sweep.py --envs 16384 262144 1048576 --rays 64 180 360 --out gpurun_out/sweep_${TAG}.jsonl > gpurun_out/sweep_${TAG}.log 2>&1; grep '^{' gpurun_out/sweep_${TAG}.log | cut -c1-230
