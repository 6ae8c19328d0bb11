set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.txt 2>&1; tail -3 gpurun_out/r2_gputest.txt
python bench.py --steps 20 --warmup 5 --strong-generations 0 --no-cpu-baseline > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -c 300 gpurun_out/r2_bench_g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_g.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['gpu_launches'], 'e2e', d['e2e']['value']/1e9, 3670012/d['e2e']['value']*1e3, 'ms')
PY
