set -x
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 scripts/dist_stamps.py 21 2>&1 | grep -v "^W\|^\*\|OMP_NUM\|warn\|colors ="
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 --strong-generations 0 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err; cut -c1-300 gpurun_out/r2_bench_n2_c.json
for a in "11 1 64 tree peer" "10 1 64 arterial peer" "14 1 2048 tree peer"; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/dist_check.py $a > gpurun_out/dc.log 2>&1; grep "dist_check\|Error:\|rror" gpurun_out/dc.log | grep -v "errors.html\|error_file" | tail -3
done
python scripts/time_kernels.py 20
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_t512.so python scripts/time_kernels.py 20
