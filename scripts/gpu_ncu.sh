set -x
CMD="python bench.py --steps 2 --warmup 3 --potentials 512 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rel_integrate -s 3 -c 1 -o gpurun_out/prof_rel $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_plain.log
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
