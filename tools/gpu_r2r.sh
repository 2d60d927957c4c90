TAG=${1:-r2r}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -4 gpurun_out/test_gpu_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-600
( time timeout 900 python bench.py --dense --no-cpu-baseline > gpurun_out/bench_${TAG}_dense.json 2> gpurun_out/bench_${TAG}_dense.err ) 2>&1 | grep real; grep '^\[bench\]' gpurun_out/bench_${TAG}_dense.err | cut -c1-600
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_${TAG}_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
