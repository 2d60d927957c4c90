TAG=${1:-r2u}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -3 gpurun_out/test_gpu_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-420
timeout 900 python bench.py --workload land --envs 131072 --n-moving 0 --n-static 0 --n-polygons 512 --n-paths 256 --no-cpu-baseline --no-e2e --steps 30 --preroll-steps 300 > gpurun_out/bench_${TAG}_land.json 2> gpurun_out/bench_${TAG}_land.err; grep '^\[bench\]' gpurun_out/bench_${TAG}_land.err | cut -c1-420
timeout 600 python bench.py --dense --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_dense.json 2> gpurun_out/bench_${TAG}_dense.err; grep '^\[bench\]' gpurun_out/bench_${TAG}_dense.err | cut -c1-300
