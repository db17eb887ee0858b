#!/bin/bash
mkdir -p gpurun_out
PHSH_B200_TRACE=1 timeout 300 python scripts/exp/chain_latency.py 2> gpurun_out/trace_all.txt > /dev/null
tail -60 gpurun_out/trace_all.txt > gpurun_out/r2_chain_trace.txt
cat gpurun_out/r2_chain_trace.txt
