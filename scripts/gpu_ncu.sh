set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --potentials 512 --no-cpu-baseline"
$CMD2 > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rel_integrate -s 3 -c 1 -o gpurun_out/prof_rel_r1 $CMD2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
