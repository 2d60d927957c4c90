# ncu --set full capture of the two step kernels in STEADY STATE (600th pre-roll step).  Usage: bash tools/gpu_ncu_r2.sh TAG [LIB]
TAG=${1:-r2}
[ -n "$2" ] && export AUV_B200_LIB=$PWD/$2
B="python bench.py --scenario-cache /tmp/scn --steps 2 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e --gpu-scenarios --preroll-steps 600"
$B > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_lidar|k_vessel_nav' --launch-skip 1208 --launch-count 2 \
    -o gpurun_out/prof_$TAG -f $B > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
