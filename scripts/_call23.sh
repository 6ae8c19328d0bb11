set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mirror or accumulated or alias or rhs_only" 2>&1 | tail -3
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python scripts/tree_stamps.py 20 > gpurun_out/r2_stamps20.txt 2>&1; cat gpurun_out/r2_stamps20.txt
python bench.py --steps 3 --warmup 3 --strong-generations 0 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --strong-generations 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/step_once.py 20 4 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:assemble_tiles|tree_factor_solve|edge_backsub|spmv_pipe" -s 8 -c 4 -o gpurun_out/r2_step_prof -f python scripts/step_once.py 20 4 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
