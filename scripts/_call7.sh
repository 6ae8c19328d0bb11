set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 1500 gpurun_out/r2_bench_d.err
python bench.py --workload arterial --steps 10 --warmup 3 --strong-generations 0 --no-cpu-baseline > gpurun_out/r2_bench_art.json 2> gpurun_out/r2_bench_art.err; tail -c 1500 gpurun_out/r2_bench_art.err
python scripts/step_once.py 20 4 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:assemble_tiles|tree_factor_solve|edge_backsub|spmv_pipe" -s 8 -c 4 -o gpurun_out/r2_step_prof -f python scripts/step_once.py 20 4 > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log
