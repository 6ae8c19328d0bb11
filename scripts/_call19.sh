set -x
timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -3
for a in "9 4 32 tree peer" "8 4 32 arterial peer" "12 3 128 arterial auto"; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/dist_check.py $a > gpurun_out/dc.log 2>&1; grep "dist_check\|Error:\|rror" gpurun_out/dc.log | grep -v "errors.html\|error_file" | tail -3
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --workload arterial --cells-per-edge 4 --generations 16 --steps 10 --warmup 3 --strong-generations 0 > gpurun_out/r2_bench_art4_n2.json 2> gpurun_out/r2_bench_art4_n2.err; grep -v "^W\|^\*\|OMP_NUM\|warn\|colors =" gpurun_out/r2_bench_art4_n2.err | tail -4; cut -c1-300 gpurun_out/r2_bench_art4_n2.json
