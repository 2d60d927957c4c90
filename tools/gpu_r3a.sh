TAG=${1:-r3a}
run() { name=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "== $name"; grep '^\[bench\]' gpurun_out/bench_${TAG}_$name.err | cut -c1-135; tail -2 gpurun_out/bench_${TAG}_$name.err | grep -i error | cut -c1-200; }
run base
AUV_B200_LIB=gym_auv_b200/variants/lib_nocoop.so run nocoop
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -3 gpurun_out/test_gpu_$TAG.log
