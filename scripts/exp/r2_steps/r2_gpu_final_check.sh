#!/bin/bash
# what the driver runs at round end: smoke, the GPU suite, both bench arms with default flags; plus the one ncu capture
# pass 2 missed (the light finish kernel)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.txt
( time python bench.py ) 2> gpurun_out/final_default_bench.err | cut -c1-300
tail -4 gpurun_out/final_default_bench.err
( time python bench.py --impl reference ) 2> gpurun_out/final_default_ref.err | cut -c1-300
tail -4 gpurun_out/final_default_ref.err
CMD3="python scripts/exp/cavpot_dev_time.py 4096 2"
ncu --set full --clock-control none --import-source on -k regex:cavpot_finish -s 2 -c 1 -f -o gpurun_out/final_prof_cavpot_finish $CMD3 > gpurun_out/ncu_cavpot_finish.log 2>&1
ncu -i gpurun_out/final_prof_cavpot_finish.ncu-rep --page raw --csv > gpurun_out/final_prof_cavpot_finish_raw.csv 2>/dev/null
rm -f gpurun_out/final_prof_cavpot_finish.ncu-rep
tail -2 gpurun_out/ncu_cavpot_finish.log
