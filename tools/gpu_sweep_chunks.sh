set -x

for ch in 1 2 4 8; do
  python bench.py --steps 50 --warmup 5 --chunks $ch --no-cpu-baseline > gpurun_out/bench_c$ch.json 2> gpurun_out/bench_c$ch.err; tail -2 gpurun_out/bench_c$ch.err
done
python bench.py --steps 50 --warmup 5 --chunks 8 --chunk-streams 8 --no-cpu-baseline > gpurun_out/bench_c8s8.json 2> gpurun_out/bench_c8s8.err
python bench.py --steps 50 --warmup 5 --chunks 16 --chunk-streams 4 --no-cpu-baseline > gpurun_out/bench_c16s4.json 2> gpurun_out/bench_c16s4.err
python bench.py --steps 50 --warmup 5 --chunks 4 --chunk-streams 2 --no-cpu-baseline > gpurun_out/bench_c4s2.json 2> gpurun_out/bench_c4s2.err
