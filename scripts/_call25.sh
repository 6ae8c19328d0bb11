set -x
for v in lib_base lib_pf lib_pair; do NXFX_LIB=gpurun_out/variants/$v.so timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1; done
timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python scripts/tree_stamps.py 20 2>&1 | tail -24
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
