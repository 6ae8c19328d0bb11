set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -15 gpurun_out/r2_gputest.txt
for lib in libnxfx_b200_stamps.so libnxfx_b200_stamps512.so; do
NXFX_LIB=networks_fenicsx_b200/csrc/$lib timeout 300 python scripts/tree_stamps.py 20 > gpurun_out/r2_stamps_$lib.txt 2>&1; cat gpurun_out/r2_stamps_$lib.txt
NXFX_LIB=networks_fenicsx_b200/csrc/$lib python scripts/time_kernels.py 20
done
python scripts/time_kernels.py 20
for mb in 16 34 48 64; do NXFX_ASM_PERSIST_MB=$mb python scripts/time_kernels.py 20; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
grep -v "^==" gpurun_out/r2_launches_b.csv | tail -16 | cut -d, -f5,15 
