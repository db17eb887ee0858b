python -m pytest tests/test_rel_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rel_energy_axis -c 4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "gpu__time" | head -4
