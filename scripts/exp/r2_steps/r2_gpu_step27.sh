#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rel_gpu.py tests/test_full_size_gpu.py tests/test_files_gpu.py tests/test_robustness_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'parity', d.get('parity_max_abs_err_rad'))"
