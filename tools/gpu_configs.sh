# BASELINE configs 2, 4 and 5 with the current kernels
python bench.py --workload pathfollow --envs 4096 --no-cpu-baseline --steps 50 > gpurun_out/bench_pathfollow.json 2> gpurun_out/bench_pathfollow.err; grep '^\[bench\]' gpurun_out/bench_pathfollow.err | cut -c1-160
python bench.py --workload pathfollow --envs 65536 --no-cpu-baseline --steps 50 > gpurun_out/bench_pathfollow_64k.json 2> gpurun_out/bench_pathfollow_64k.err; grep '^\[bench\]' gpurun_out/bench_pathfollow_64k.err | cut -c1-160
python bench.py --workload land --envs 131072 --n-moving 0 --n-static 0 --n-polygons 512 --no-cpu-baseline --steps 30 > gpurun_out/bench_land.json 2> gpurun_out/bench_land.err; grep '^\[bench\]' gpurun_out/bench_land.err | cut -c1-160; tail -2 gpurun_out/bench_land.err | cut -c1-200
python tools/sweep.py --out gpurun_out/sweep.jsonl > gpurun_out/sweep.log 2>&1; tail -30 gpurun_out/sweep.log | cut -c1-160
