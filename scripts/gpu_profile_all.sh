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
# ncu --set full captures; the reports are turned into raw-metric and SASS-source CSVs on the box (gpurun_out/ is capped at
# 64 MiB), only the rel_integrate report itself is kept
cap() {  # cap <name> <kernel regex> <skip> <command...>
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/final_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/final_prof_$name.ncu-rep --page raw --csv > gpurun_out/final_prof_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/final_prof_$name.ncu-rep --page source --csv > gpurun_out/final_prof_${name}_source.csv 2>/dev/null
  [ "$name" = rel ] || rm -f gpurun_out/final_prof_$name.ncu-rep
  tail -1 gpurun_out/ncu_$name.log | cut -c1-160
}
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/ncu_plain2.log 2>&1 || exit 1
cap rel rel_integrate 3 $CMD2
cap axis rel_energy_axis 3 $CMD2
python scripts/exp/cav_one_launch.py > gpurun_out/cav_plain.log 2>&1 && cap cav cav_ps_kernel 1 python scripts/exp/cav_one_launch.py
CMD3="python scripts/exp/cavpot_dev_time.py 4096 2"
$CMD3 > gpurun_out/cavpot_plain.log 2>&1 || exit 1
for k in prep sumax classify mtz_sum; do cap cavpot_$k cavpot_$k 1 $CMD3; done
cap cavpot_finish cavpot_finish 2 $CMD3   # launches alternate light / heavy: the third one is the light instantiation of the second call
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_cavpot_launches.csv $CMD3 > gpurun_out/ncu_cavpot_launch.log 2>&1
python scripts/exp/wil_time.py > gpurun_out/wil_plain.log 2>&1 && cap wil wil_integrate 1 python scripts/exp/wil_time.py
cat gpurun_out/cavpot_plain.log gpurun_out/wil_plain.log
du -sh gpurun_out
fi
