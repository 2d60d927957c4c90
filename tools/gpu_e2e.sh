TAG=${1:-e2e}
for t in 4 8 16 32; do
  timeout 300 python bench.py --scenario-cache /tmp/scn --no-cpu-baseline --steps 20 --preroll-steps 300 --host-threads $t > gpurun_out/bench_${TAG}_t$t.json 2> gpurun_out/bench_${TAG}_t$t.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}_t$t.json')); e=d['e2e']; print('threads $t', 'e2e', round(e['value']/1e6,1), 'M  ms', round(e['ms_per_step'],3), 'expand_ms', round(e.get('host_expand_ms_per_step',0),3), 'sync', round(e['sync']['ms_per_step'],3), 'value', round(d['value']/1e6,1))
"
done
nproc; cat /proc/cpuinfo | grep "model name" | head -1; python -c "import os; print(len(os.sched_getaffinity(0)))"
