python -m pytest tests -m gpu -x -q -k "chunked or step_host" > gpurun_out/test_gpu_e2e.log 2>&1; tail -3 gpurun_out/test_gpu_e2e.log
for hc in 1 4 8 16 32; do
  python bench.py --scenario-cache /tmp/scn --steps 30 --no-cpu-baseline --host-chunks $hc > gpurun_out/bench_hc$hc.json 2> gpurun_out/bench_hc$hc.err
  echo "hc=$hc $(grep '^\[bench\]' gpurun_out/bench_hc$hc.err | sed 's/.*e2e=//' | cut -c1-200)"
done
