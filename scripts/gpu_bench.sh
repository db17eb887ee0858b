set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_err.log | tee gpurun_out/bench2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline'])"
python bench.py --steps 3 --warmup 3 --fast --no-cpu-baseline 2>> gpurun_out/bench_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('FAST', d['value'], d['ms_per_step'], d['roofline']['frac'])"
tail -3 gpurun_out/bench_err.log
