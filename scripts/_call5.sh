set -x
for i in 1 2; do
python scripts/time_kernels.py 20
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_rhsstream.so python scripts/time_kernels.py 20
done
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python scripts/tree_stamps.py 20
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -c 300 gpurun_out/r2_bench_c.err; cut -c1-400 gpurun_out/r2_bench_c.json
