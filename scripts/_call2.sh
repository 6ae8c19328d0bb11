set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -15 gpurun_out/r2_gputest.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 400 gpurun_out/r2_bench_a.err
python scripts/time_kernels.py 20 > gpurun_out/r2_tk_a.txt 2>&1; cat gpurun_out/r2_tk_a.txt
