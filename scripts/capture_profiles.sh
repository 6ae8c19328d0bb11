# Single-GPU capture recipe of the round (run under gpurun): bench line, reference arm, ncu launch list, ncu --set full of the
# in-step kernels and of one higher-order step.  Summaries: profiles/summarize_ncu.py, profiles/make_traffic.py.
set -x
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_bench_n1_final.err > gpurun_out/r2_bench_n1_final.json; cut -c1-200 gpurun_out/r2_bench_n1_final.json
python bench.py --impl reference --steps 1 --warmup 0 2>gpurun_out/r2_bench_ref.err > gpurun_out/r2_bench_ref.json; cut -c1-300 gpurun_out/r2_bench_ref.json
python bench.py --steps 3 --warmup 3 --strong-generations 0 --higher-order-generations 0 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --strong-generations 0 --higher-order-generations 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/step_once.py 20 4 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:assemble_tiles|tree_factor_solve|edge_backsub|spmv_pipe" -s 8 -c 4 -o gpurun_out/r2_step_prof -f python scripts/step_once.py 20 4 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
python scripts/step_once_generic.py 16 3 > gpurun_out/plain_generic.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:cond_|btree_|generic" -s 26 -c 13 -o gpurun_out/r2_generic_prof -f python scripts/step_once_generic.py 16 3 > gpurun_out/ncu_generic.log 2>&1
tail -2 gpurun_out/ncu_generic.log; cat gpurun_out/plain_generic.log
