set -x
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python scripts/tree_stamps.py 20
python scripts/time_kernels.py 20
python scripts/step_once.py 20 4 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tree_factor_solve_coop -s 2 -c 1 -o gpurun_out/r2_tree_prof -f python scripts/step_once.py 20 4 > gpurun_out/ncu_tree.log 2>&1
tail -3 gpurun_out/ncu_tree.log
