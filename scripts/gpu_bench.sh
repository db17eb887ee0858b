set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rel_integrate -s 3 -c 1 -o gpurun_out/prof_rel_r1_full $CMD > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log | cut -c1-200
