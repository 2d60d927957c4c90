TAG=${1:-r2t}
for v in "" nolong; do
  if [ -n "$v" ]; then export AUV_B200_LIB=gym_auv_b200/variants/lib_$v.so; fi
  timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err; echo "variant=$v"; grep '^\[bench\]' gpurun_out/bench_${TAG}_$v.err | cut -c1-330
done
unset AUV_B200_LIB
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -3 gpurun_out/test_gpu_$TAG.log
