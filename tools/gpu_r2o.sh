TAG=${1:-r2o}
./tools/probes/pcie_store_probe > gpurun_out/pcie_probe_$TAG.txt 2>&1; head -50 gpurun_out/pcie_probe_$TAG.txt
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_$TAG.log 2>&1; tail -4 gpurun_out/test_gpu_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; grep '^\[bench\]' gpurun_out/bench_$TAG.err | sed 's/.*e2e=/e2e=/' | cut -c1-1800
