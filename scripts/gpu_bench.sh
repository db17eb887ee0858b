set -x
python scripts/bench_paths.py > gpurun_out/secondary.jsonl 2> gpurun_out/secondary.err
cat gpurun_out/secondary.jsonl; tail -3 gpurun_out/secondary.err
CMD="python scripts/bench_paths.py"
ncu --set full --clock-control none --import-source on -k regex:wil_integrate -s 2 -c 1 -o gpurun_out/prof_wil_r1 $CMD > gpurun_out/ncu_wil.log 2>&1
tail -2 gpurun_out/ncu_wil.log | cut -c1-200
