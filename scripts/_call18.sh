set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -5 gpurun_out/r2_gputest.txt
python scripts/time_kernels.py 20
python bench.py --workload arterial --cells-per-edge 4 --generations 16 --steps 10 --warmup 3 --strong-generations 0 --no-cpu-baseline > gpurun_out/r2_bench_art4.json 2> gpurun_out/r2_bench_art4.err; tail -c 600 gpurun_out/r2_bench_art4.err; cut -c1-300 gpurun_out/r2_bench_art4.json
