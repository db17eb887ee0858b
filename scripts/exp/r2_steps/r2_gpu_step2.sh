#!/bin/bash
# round-2 GPU step 2: cav kernel after the sqrt/r^2/Lagrange hoisting -- parity tests, scale probe, ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "cav or c2 or full_size or files" 2>&1 | tail -8 > gpurun_out/r2_gputests2.txt
cat gpurun_out/r2_gputests2.txt
python scripts/exp/cav_scale_probe.py > gpurun_out/r2_cav_scale.txt 2>&1; cat gpurun_out/r2_cav_scale.txt
python scripts/exp/cav_one_launch.py > gpurun_out/r2_cav_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cav_ps_kernel -s 1 -c 1 -o gpurun_out/r2_prof_cav python scripts/exp/cav_one_launch.py > gpurun_out/r2_ncu_cav.log 2>&1
tail -3 gpurun_out/r2_ncu_cav.log | cut -c1-200
python bench.py --potentials 512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p512_new.json 2> gpurun_out/r2_p512_new_err.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_p512_new.json").read().strip().splitlines()[-1])
print("P512 value %.4g ms %.3f e2e %.4g pageable %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pageable"]["value"]))
PY
