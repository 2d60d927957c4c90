for cfg in "1 1" "2 2" "4 4" "6 6" "8 8" "8 4" "12 6" "16 8"; do
  set -- $cfg
  python bench.py --scenario-cache /tmp/scn --gpu-scenarios --steps 50 --no-cpu-baseline --no-e2e --chunks $1 --chunk-streams $2 > gpurun_out/bench_cc$1_$2.json 2> gpurun_out/bench_cc$1_$2.err
  echo "chunks=$1 streams=$2 $(grep '^\[bench\]' gpurun_out/bench_cc$1_$2.err | cut -c1-60)"
done
