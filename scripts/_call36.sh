set -x
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --steps 20 --warmup 5 2>gpurun_out/r2_bench_n4_final.err | tee gpurun_out/r2_bench_n4_final.json | cut -c1-300
