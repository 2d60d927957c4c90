# GPU tests + one bench per chunk setting given as arguments (default: 1 4)
python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; tail -15 gpurun_out/test_gpu.log
for ch in ${@:-1 4}; do
  python bench.py --steps 50 --warmup 5 --chunks $ch --no-cpu-baseline > gpurun_out/bench_q$ch.json 2> gpurun_out/bench_q$ch.err; tail -2 gpurun_out/bench_q$ch.err
done
