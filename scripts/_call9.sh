set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -c 1200 gpurun_out/r2_bench_n8.err; cut -c1-400 gpurun_out/r2_bench_n8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 tests/dist_check.py 13 1 64 arterial peer 2>&1 | grep dist_check
