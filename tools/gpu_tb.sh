# GPU tests + default bench.  Usage: bash tools/gpu_tb.sh TAG [bench args]
TAG=${1:-x}; shift
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -4 gpurun_out/test_gpu.log
python bench.py --scenario-cache /tmp/scn "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-900
