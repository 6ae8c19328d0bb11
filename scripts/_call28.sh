set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "higher_order" 2>&1 | tail -5
timeout 300 python scripts/time_generic.py 16 4 2>&1 | tail -3
