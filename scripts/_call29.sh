set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "higher_order" 2>&1 | tail -3
timeout 600 python scripts/time_generic.py 19 4 2>&1 | tail -3
NXFX_COND_SMEM=0 timeout 600 python scripts/time_generic.py 19 4 2>&1 | tail -3
