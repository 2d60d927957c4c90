TAG=${1:-r3d}
run() { name=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "== $name"; grep '^\[bench\]' gpurun_out/bench_${TAG}_$name.err | cut -c1-135; tail -2 gpurun_out/bench_${TAG}_$name.err | grep -i error | cut -c1-200; }
run base
AUV_B200_LIB=gym_auv_b200/variants/lib_oldcull.so run oldcull
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -2
