# bench every library build in build_variants/ (tuning sweeps; AUV_B200_LIB override)
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -5 gpurun_out/test_gpu.log
for lib in build_variants/lib_*.so; do
  name=$(basename $lib .so)
  AUV_B200_LIB=$PWD/$lib python bench.py --steps 30 --warmup 5 --chunks 1 --no-cpu-baseline --no-e2e --scenario-cache /tmp/scn \
     > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name $(grep '^\[bench\]' gpurun_out/bench_$name.err | cut -c1-330)"
done
