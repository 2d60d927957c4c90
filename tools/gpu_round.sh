# full GPU pass: tests, default bench (chunk settings 1 and 4), ncu captures.  Usage: bash tools/gpu_round.sh TAG
TAG=${1:-r1x}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -5 gpurun_out/test_gpu.log
python bench.py --scenario-cache /tmp/scn --chunks 1 > gpurun_out/bench_${TAG}_c1.json 2> gpurun_out/bench_${TAG}_c1.err; grep '^\[bench\]' gpurun_out/bench_${TAG}_c1.err | cut -c1-200
python bench.py --scenario-cache /tmp/scn > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; grep '^\[bench\]' gpurun_out/bench_${TAG}.err | cut -c1-200
bash tools/gpu_ncu.sh $TAG
