#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "cav or c2" 2>&1 | tail -5
: > gpurun_out/r2_cav_variants.txt
for v in "1 7" "1 6" "1 5" "2 7" "2 6" "2 5" "2 4" "3 6" "3 5" "3 4" "4 5" "4 4"; do
  set -- $v
  echo "NCH=$1 MINB=$2: $(PHSH_B200_CAV_NCH=$1 PHSH_B200_CAV_MINB=$2 python scripts/exp/cav_scale_probe.py 4096 512 2>&1 | tr '\n' ' ')" >> gpurun_out/r2_cav_variants.txt
done
cat gpurun_out/r2_cav_variants.txt
