set -x
python scripts/time_kernels.py 20
python scripts/time_kernels.py 20
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -3 gpurun_out/r2_gputest.txt
