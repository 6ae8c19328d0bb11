set -x
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_demos.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --workload arterial --cells-per-edge 4 --generations 16 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench_art2_final.err | tee gpurun_out/r2_bench_art2_final.json | cut -c1-250
