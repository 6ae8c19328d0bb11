set -x
for i in 1 2; do
NXFX_POLL_NORMS=0 timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
