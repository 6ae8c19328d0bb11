set -x
nvidia-smi -L
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err; tail -c 600 gpurun_out/r2_base_bench.err
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python scripts/tree_stamps.py 20 > gpurun_out/r2_stamps20.txt 2>&1
cat gpurun_out/r2_stamps20.txt
