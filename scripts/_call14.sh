set -x
python scripts/time_kernels.py 20
python scripts/time_kernels.py 20
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -5 gpurun_out/r2_gputest.txt
python bench.py --steps 20 --warmup 5 --strong-generations 0 --no-cpu-baseline > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; cut -c1-250 gpurun_out/r2_bench_e.json
