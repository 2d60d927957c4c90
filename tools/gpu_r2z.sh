TAG=${1:-r2z}
run() { name=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "== $name"; grep '^\[bench\]' gpurun_out/bench_${TAG}_$name.err | cut -c1-135; tail -2 gpurun_out/bench_${TAG}_$name.err | grep -i error | cut -c1-200; }
run base
for v in pu1 pu4 pu8; do AUV_B200_LIB=gym_auv_b200/variants/lib_$v.so run $v; done
