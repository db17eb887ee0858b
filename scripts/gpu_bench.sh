set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench5.json | cut -c1-200
tail -3 gpurun_out/bench_err.log
python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench5_ref.json | cut -c1-200
