TAG=${1:-r3b}
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; grep '^\[bench\]' gpurun_out/bench_$TAG.err | cut -c1-420
# k_nav_cull alone (the one kernel that changed since the r2z capture): 1500th step after reset, --set full with source
P="python bench.py --scenario-cache /tmp/scn --steps 4 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e --preroll-steps 1500 --refresh-every 100000"
( time timeout 420 ncu --set full --clock-control none --import-source on -k regex:'k_nav_cull' --launch-skip 1503 --launch-count 1 \
    -o gpurun_out/prof_$TAG -f $P > gpurun_out/ncu_$TAG.log 2>&1 ) 2>&1 | grep real
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
