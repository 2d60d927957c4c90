TAG=${1:-r2j}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -6 gpurun_out/test_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-1500
bash tools/gpu_ncu_r2.sh $TAG
