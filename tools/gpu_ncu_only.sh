bash tools/gpu_ncu.sh ${1:-r1x}
