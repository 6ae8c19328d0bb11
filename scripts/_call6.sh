set -x
nvidia-smi -L
for a in "11 1 64 tree peer" "11 1 64 tree nccl" "10 1 64 arterial peer" "9 4 32 tree auto" "8 4 32 arterial nccl"; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/dist_check.py $a 2>&1 | grep -v "^W\|warn" | tail -4
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_peer.json 2> gpurun_out/r2_bench_n2_peer.err; tail -c 600 gpurun_out/r2_bench_n2_peer.err; cut -c1-300 gpurun_out/r2_bench_n2_peer.json
