#!/bin/bash
# round-2 pass 1 (tests, bench lines, secondary paths, launch list) + cavpot occupancy variants
bash scripts/gpu_profile_all.sh 1
for v in "" scripts/exp/libvar_minb5.so scripts/exp/libvar_minb6.so scripts/exp/libvar_minb8.so; do
  for S in 4096 1; do
    PHSH_B200_LIB=${v:+$PWD/$v} timeout 300 python scripts/exp/cavpot_dev_time.py $S 2>&1 | tail -1 | tee -a gpurun_out/cavpot_variants.txt
  done
done
