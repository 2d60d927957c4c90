TAG=${1:-r2p}
python -m pytest tests/test_gpu_v2.py tests/test_gpu_parity.py -m gpu -x -q -k "ring or delta or step_host or async or compact" > gpurun_out/test_gpu_$TAG.log 2>&1; tail -4 gpurun_out/test_gpu_$TAG.log
( time timeout 900 python bench.py --e2e-groups 2 3 4 6 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2p.json').read().strip().splitlines()[-1]); e=d['e2e']
f=lambda r:{k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in('value','ms_per_step','host_transfer','groups','d2h_bytes_per_step','host_expand_ms_per_step','host_blocked_ms_per_step')}
print('value',round(d['value']/1e6,1)); print(f(e)); [print(f(r)) for r in e['other']]; print('sync',f(e['sync']))
PY
