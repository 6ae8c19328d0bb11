set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; grep -v "^W\|^\*\|OMP_NUM\|warn\|colors =" gpurun_out/r2_bench_n8.err | tail -5; cut -c1-400 gpurun_out/r2_bench_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; cut -c1-400 gpurun_out/r2_bench_n4.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29653 tests/dist_check.py 13 1 64 arterial peer 2>&1 | grep dist_check
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29654 tests/dist_check.py 12 4 64 arterial nccl 2>&1 | grep dist_check
