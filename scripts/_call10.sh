set -x
for a in "11 1 64 tree peer" "11 1 64 tree nccl" "10 1 64 arterial peer" "9 4 32 tree auto" "8 4 32 arterial nccl" "14 1 2048 tree peer"; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/dist_check.py $a 2>&1 | grep "dist_check\|Error\|error" | tail -4
done
timeout 900 python -m pytest tests/test_gpu_demos.py tests/test_gpu_distributed.py -x -q 2>&1 | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; tail -c 600 gpurun_out/r2_bench_n2_b.err; cut -c1-300 gpurun_out/r2_bench_n2_b.json
