# end-of-round pass: tests, default bench (with cpu baseline), ncu captures.  Usage: bash tools/gpu_final.sh TAG
TAG=${1:-r1x}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -4 gpurun_out/test_gpu.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; grep '^\[bench\]' gpurun_out/bench_${TAG}.err | cut -c1-300
bash tools/gpu_ncu.sh $TAG
