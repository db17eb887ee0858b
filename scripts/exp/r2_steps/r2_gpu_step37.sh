#!/bin/bash
mkdir -p gpurun_out
python scripts/exp/hartfock_batch_ab.py 148 2>&1 | tee gpurun_out/step37_ab.txt
