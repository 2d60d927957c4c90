# ncu captures of the step kernels in STEADY STATE (1500 pre-roll steps): --set full with source, and the launch list.
# Usage: bash tools/gpu_ncu_r2.sh TAG
TAG=${1:-r2}
P="python bench.py --scenario-cache /tmp/scn --steps 4 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e --preroll-steps 1500 --refresh-every 100000"
$P > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || exit 1
# launches before the capture (k_vessel_nav, k_nav_cull, k_lidar per step): reset cache 2 x 3 + reset 3 + (3 warm-up + 4 after-reset + 1493 pre-roll) x 3
ncu --set full --clock-control none --import-source on -k regex:'k_lidar|k_vessel_nav|k_nav_cull' --launch-skip 4509 --launch-count 3 \
    -o gpurun_out/prof_$TAG -f $P > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
L="python bench.py --scenario-cache /tmp/scn --steps 2 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e --preroll-steps 40"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|auv' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $L > gpurun_out/ncu_l_$TAG.log 2>&1
tail -3 gpurun_out/launches_$TAG.csv | cut -c1-200
