# tests, default bench (sustained headline incl. e2e), ncu of both kernels in steady state.  Usage: bash tools/gpu_r2c.sh TAG
TAG=${1:-r2c}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -12 gpurun_out/test_gpu_$TAG.log
B="python bench.py --scenario-cache /tmp/scn --no-cpu-baseline"
timeout 600 $B > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-900
tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
# steady-state capture: launches before = reset cache 2 + reset 2 + (3 warm-up + 50 after-reset + pre-roll) x 2 + refresh
P="python bench.py --scenario-cache /tmp/scn --steps 4 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e --preroll-steps 1500 --refresh-every 100000"
ncu --set full --clock-control none --import-source on -k regex:'k_lidar|k_vessel_nav' --launch-skip 3004 --launch-count 2 \
    -o gpurun_out/prof_$TAG -f $P > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
