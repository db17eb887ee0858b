set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench1.json
tail -5 gpurun_out/bench_err.log
python bench.py --steps 3 --warmup 3 --fast --no-cpu-baseline 2>> gpurun_out/bench_err.log | tee gpurun_out/bench1_fast.json
