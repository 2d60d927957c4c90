# one ncu --set full capture of k_vessel_nav<true> and k_lidar from a short bench run (single stream),
# plus the launch list of this library's kernels.  Usage: bash tools/gpu_ncu.sh TAG
TAG=${1:-r1x}
B="python bench.py --steps 2 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_$TAG.log 2>&1 || exit 1
# matched launches before the capture: reset cache (nav,lidar) + reset (nav,lidar) + 3 warm-up steps (nav,lidar) x 3
ncu --set full --clock-control none --import-source on -k regex:'k_lidar|k_vessel_nav' --launch-skip 8 --launch-count 2 \
    -o gpurun_out/prof_$TAG -f $B > gpurun_out/ncu_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|auv' -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    $B > gpurun_out/ncu_l_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
