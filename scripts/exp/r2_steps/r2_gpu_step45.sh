#!/bin/bash
# diagnostic: cycles per block of four of a consumer that is alone on its scheduler (CTA 15, Re's outermost orbital), with
# the producers at work (a) and with producers that stop after the second tile of every round (b: wrong results, same loop)
mkdir -p gpurun_out
for v in a b; do
  echo "== variant $v"
  PHSH_B200_LIB=$PWD/phaseshifts_b200/libphsh_b200_dbg$v.so timeout 90 python scripts/exp/hartfock_one.py 2>&1 | grep "consumer of CTA\|Re batch" | tail -2
done | tee gpurun_out/step45_diag.txt
