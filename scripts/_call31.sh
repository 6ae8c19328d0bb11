set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_bench_n1_final.err | tee gpurun_out/r2_bench_n1_final.json | cut -c1-250
python bench.py --workload arterial --cells-per-edge 4 --generations 16 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench_art_final.err | tee gpurun_out/r2_bench_art_final.json | cut -c1-250
