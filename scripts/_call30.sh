set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2_bench_n2_final.err | tee gpurun_out/r2_bench_n2_final.json | cut -c1-300
