TAG=${1:-x}; shift
python -m pytest tests -m gpu -x -q -k "async or chunked or step_host" > gpurun_out/test_gpu_k.log 2>&1; tail -4 gpurun_out/test_gpu_k.log
python bench.py --scenario-cache /tmp/scn "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | sed 's/.*e2e=/e2e=/' | cut -c1-1200; tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
