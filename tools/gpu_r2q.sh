TAG=${1:-r2q}
python -m pytest tests/test_gpu_v2.py tests/test_gpu_parity.py -m gpu -x -q -k "ring or delta or step_host or async or compact" > gpurun_out/test_gpu_$TAG.log 2>&1; tail -3 gpurun_out/test_gpu_$TAG.log
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
f=lambda r:{k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in('value','ms_per_step','host_transfer','groups','d2h_bytes_per_step','host_blocked_ms_per_step')}
print(f(e)); [print(f(r)) for r in e['other']]; print('sync',f(e['sync']))
PY
}
for c in 0 148 592 1184; do
echo "ctas=$c"; AUV_B200_DELTA_CTAS=$c timeout 600 python bench.py --host-transfer delta --e2e-groups 2 4 --no-cpu-baseline > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; show gpurun_out/bench_${TAG}_c$c.json
done
