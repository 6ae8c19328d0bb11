set -x
python scripts/time_kernels.py 20
python scripts/time_kernels.py 22
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -3 gpurun_out/r2_gputest.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; tail -c 300 gpurun_out/r2_bench_f.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['gpu_launches'], d['strong']['ms_per_step'], d['strong']['gpu_launches_per_step'], d['strong']['parity'])
PY
