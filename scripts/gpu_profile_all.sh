# Regenerates every artefact under profiles/ for the current tree (1 GPU), one ncu pass per call.
# Usage: gpurun -- bash scripts/gpu_profile_all.sh 1   (tests, bench lines, secondary paths, launch list)
#        gpurun -- bash scripts/gpu_profile_all.sh 2   (ncu --set full captures: rel_integrate on the C4 bench command,
#                                                        rel_energy_axis, cav_ps, wil_integrate, cavpot)
set -x
mkdir -p gpurun_out
if [ "${1:-1}" = 1 ]; then
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.log > gpurun_out/final_bench_n1.json
cp gpurun_out/fp64_peak_n1.json gpurun_out/final_fp64_peak.json  # before the profiled runs overwrite it
python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/bench_err.log > gpurun_out/final_bench_n1_reference.json
python scripts/bench_paths.py > gpurun_out/final_secondary.jsonl 2>> gpurun_out/bench_err.log
python scripts/exp/chain_latency.py > gpurun_out/final_chain_latency.json 2>> gpurun_out/bench_err.log
python scripts/exp/size_sweep.py > gpurun_out/final_size_sweep.jsonl 2>> gpurun_out/bench_err.log
python scripts/exp/cav_scale_probe.py > gpurun_out/final_cav_scale.txt 2>> gpurun_out/bench_err.log
PHSH_B200_GUARD=1 python scripts/guard_run.py > gpurun_out/final_guard_run.txt 2>> gpurun_out/bench_err.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
cut -c1-200 gpurun_out/final_bench_n1.json
else
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rel_integrate -s 3 -c 1 -o gpurun_out/final_prof_rel $CMD2 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:rel_energy_axis -s 3 -c 1 -o gpurun_out/final_prof_axis $CMD2 > gpurun_out/ncu_axis.log 2>&1
python scripts/exp/cav_one_launch.py > gpurun_out/cav_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cav_ps_kernel -s 1 -c 1 -o gpurun_out/final_prof_cav python scripts/exp/cav_one_launch.py > gpurun_out/ncu_cav.log 2>&1
CMD3="python scripts/exp/cavpot_dev_time.py 4096 2"
$CMD3 > gpurun_out/cavpot_plain.log 2>&1 &&
for k in prep sumax classify mtz_sum finish; do
ncu --set full --clock-control none --import-source on -k regex:cavpot_${k} -s 1 -c 1 -f -o gpurun_out/final_prof_cavpot_$k $CMD3 > gpurun_out/ncu_cavpot_$k.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_cavpot_launches.csv $CMD3 > gpurun_out/ncu_cavpot_launch.log 2>&1
python scripts/exp/wil_time.py > gpurun_out/wil_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wil_integrate -s 1 -c 1 -f -o gpurun_out/final_prof_wil python scripts/exp/wil_time.py > gpurun_out/ncu_wil.log 2>&1
tail -2 gpurun_out/ncu_full.log gpurun_out/ncu_axis.log gpurun_out/ncu_cav.log gpurun_out/ncu_cavpot_*.log gpurun_out/ncu_wil.log | cut -c1-200; cat gpurun_out/cavpot_plain.log
fi
