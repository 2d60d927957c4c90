TAG=${1:-r1x}
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -5 gpurun_out/test_gpu.log
bash tools/gpu_ncu.sh $TAG
