# final multi-GPU pass: the default bench under torchrun.  Usage: bash tools/gpu_multi_final.sh TAG NGPU
TAG=${1:-m8f}; N=${2:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - gpurun_out/bench_${TAG}.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
f=lambda r:{k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in('value','ms_per_step','host_transfer','groups','d2h_bytes_per_step','host_expand_ms_per_step','host_blocked_ms_per_step','host_threads')}
print('n_gpus',d['n_gpus'],'value',round(d['value']/1e6,1),'M ms',round(d['ms_per_step'],4),'e2e',f(e),'other',[f(r) for r in e.get('other',[])],'sync',f(e['sync']))
PY
