#!/bin/bash
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:hartfock -c 1 -f -o gpurun_out/r2_prof_hartfock python scripts/exp/hartfock_one.py > gpurun_out/ncu_hartfock.log 2>&1
ncu -i gpurun_out/r2_prof_hartfock.ncu-rep --page source --csv > gpurun_out/r2_prof_hartfock_source.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_hartfock.ncu-rep --page raw --csv > gpurun_out/r2_prof_hartfock_raw.csv 2>/dev/null
rm -f gpurun_out/r2_prof_hartfock.ncu-rep
tail -2 gpurun_out/ncu_hartfock.log; ls -la gpurun_out/r2_prof_hartfock_*
