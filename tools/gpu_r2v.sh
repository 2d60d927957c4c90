TAG=${1:-r2v}
run() { name=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "== $name"; grep '^\[bench\]' gpurun_out/bench_${TAG}_$name.err | cut -c1-130; }
run base
AUV_B200_STREAM_PRIO=1 run prio
run c3 --chunks 3 --chunk-streams 3
AUV_B200_STREAM_PRIO=1 run c3prio --chunks 3 --chunk-streams 3
run c4 --chunks 4 --chunk-streams 4
run c1 --chunks 1
