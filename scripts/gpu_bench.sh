python -m pytest tests/test_rel_gpu.py tests/test_full_size_gpu.py tests/test_robustness_gpu.py tests/test_files_gpu.py -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4', d['value'], d['ms_per_step'], d['e2e']['value'])"
python scripts/bench_paths.py 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print(d['config'][:12], round(d['gpu_ms'],4), '%.3e'%d['gpu_units_per_s'], d['max_abs_err'])"
python scripts/exp/size_sweep.py 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print(d['P'],d['NE'],d['ms'],'%.3e'%d['units_per_s'])"
