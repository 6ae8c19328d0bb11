set -x
NXFX_LIB=build/variants/lib_nosort.so timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
NXFX_LIB=build/variants/lib_sort_pf.so timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python scripts/tree_stamps.py 20 2>&1 | tail -23
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
