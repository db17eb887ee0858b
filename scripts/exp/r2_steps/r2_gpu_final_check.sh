#!/bin/bash
# what the driver runs at round end: smoke, the GPU suite, both bench arms with default flags
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.txt
( time python bench.py ) 2> gpurun_out/final_default_bench.err | cut -c1-300
tail -4 gpurun_out/final_default_bench.err
( time python bench.py --impl reference ) 2> gpurun_out/final_default_ref.err | cut -c1-300
tail -4 gpurun_out/final_default_ref.err
