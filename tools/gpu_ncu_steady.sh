# ncu --set full capture of the two step kernels in STEADY STATE: matched launches before the capture =
# reset cache 2 + reset 2 + warm-up 3 x 2 + 1799 pre-roll steps x 2 = 3608, i.e. the 1800th pre-roll step is captured.  Usage: bash tools/gpu_ncu_steady.sh TAG
TAG=${1:-steady}
ncu --set full --clock-control none --import-source on -k regex:'k_lidar|k_vessel_nav' --launch-skip 3608 --launch-count 2 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e --gpu-scenarios --preroll-steps 1800 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
