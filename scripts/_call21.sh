set -x
for a in "14 1 4 tree peer" "13 2 4 arterial peer" "11 1 64 tree peer"; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/dist_check.py $a > gpurun_out/dc.log 2>&1; grep "dist_check\|Error:\|rror" gpurun_out/dc.log | grep -v "errors.html\|error_file" | tail -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_d.json 2> gpurun_out/r2_bench_n2_d.err; grep -v "^W\|^\*\|OMP_NUM\|warn\|colors =" gpurun_out/r2_bench_n2_d.err | tail -4; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_d.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['gpu_launches'], d['strong']['ms_per_step'], d['strong']['gpu_launches_per_step'], d['strong']['exchange'], d['strong']['parity'])
PY
