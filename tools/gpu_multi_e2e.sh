# multi-GPU e2e pass: default bench (delta + compact e2e) and the other delta chunk sizes.  Usage: bash tools/gpu_multi_e2e.sh TAG NGPU
TAG=${1:-m8e}; N=${2:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1; lscpu | grep -i -E "numa|^CPU\(s\)|model name|socket" >> gpurun_out/topo_${TAG}.txt
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
f=lambda r:{k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in('value','ms_per_step','host_transfer','d2h_bytes_per_step','host_expand_ms_per_step','host_blocked_ms_per_step','host_threads')}
print('value',round(d['value']/1e6,1),'M  e2e',f(e),'other',[f(r) for r in e.get('other',[])],'sync',f(e['sync']))
PY
}
timeout 600 $T bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; show gpurun_out/bench_${TAG}.json
for g in 8 32; do
timeout 600 $T bench.py --gpus $N --no-cpu-baseline --host-transfer delta --delta-gran $g > gpurun_out/bench_${TAG}_d$g.json 2> gpurun_out/bench_${TAG}_d$g.err; show gpurun_out/bench_${TAG}_d$g.json
done
