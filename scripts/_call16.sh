set -x
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 scripts/dist_stamps.py 21 2>&1 | grep -v "^W\|^\*\|OMP_NUM\|warn\|colors =" | grep "rank\|exchange\|ticket\|staged\|flag"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 --strong-generations 0 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err; cut -c1-300 gpurun_out/r2_bench_n2_c.json
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_demos.py -x -q 2>&1 | tail -3
