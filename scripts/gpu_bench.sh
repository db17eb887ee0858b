python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rel_energy_axis -c 3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "gpu__time|dram__bytes" | head -9
