set -x
NXFX_SPLIT_ASSEMBLY=0 timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
NXFX_LIB=build/variants/lib_nopf.so timeout 300 python scripts/time_kernels.py 20 2>&1 | tail -1
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python scripts/tree_stamps.py 20 2>&1 | tail -23
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --strong-generations 0 --no-cpu-baseline 2>gpurun_out/r2_bench_split.err | tee gpurun_out/r2_bench_split.json | cut -c1-400
