set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -8 gpurun_out/r2_gputest.txt
python scripts/time_kernels.py 20
python bench.py --workload arterial --steps 10 --warmup 3 --strong-generations 0 --no-cpu-baseline > gpurun_out/r2_bench_art.json 2> gpurun_out/r2_bench_art.err; tail -c 800 gpurun_out/r2_bench_art.err
