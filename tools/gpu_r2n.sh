TAG=${1:-r2n}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -6 gpurun_out/test_gpu_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-1800
AUV_B200_LIB=gym_auv_b200/variants/lib_nogs.so timeout 600 python bench.py --no-e2e > gpurun_out/bench_${TAG}_nogs.json 2> gpurun_out/bench_${TAG}_nogs.err; grep '^\[bench\]' gpurun_out/bench_${TAG}_nogs.err | cut -c1-300
timeout 600 python bench.py --no-e2e > gpurun_out/bench_${TAG}_b.json 2> gpurun_out/bench_${TAG}_b.err; grep '^\[bench\]' gpurun_out/bench_${TAG}_b.err | cut -c1-300
for g in 8 32; do
timeout 600 python bench.py --host-transfer delta --delta-gran $g > gpurun_out/bench_${TAG}_d$g.json 2> gpurun_out/bench_${TAG}_d$g.err; grep '^\[bench\]' gpurun_out/bench_${TAG}_d$g.err | sed 's/.*e2e=/e2e=/' | cut -c1-700
done
