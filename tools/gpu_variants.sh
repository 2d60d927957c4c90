# bench every library build in build_variants/ (tuning sweeps; AUV_B200_LIB override).  Usage: bash tools/gpu_variants.sh [--tests]
if [ "$1" = "--tests" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -5 gpurun_out/test_gpu.log; fi
for lib in build_variants/lib_*.so; do
  name=$(basename $lib .so)
  AUV_B200_LIB=$PWD/$lib python bench.py --steps 50 --warmup 5 --chunks 1 --no-cpu-baseline --no-e2e --gpu-scenarios --scenario-cache /tmp/scn \
     > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name $(grep '^\[bench\]' gpurun_out/bench_$name.err | sed "s/kernel_ms={'k_obstacle_update': [0-9.e-]*, //" | cut -c1-130)"
done
