#!/bin/bash
PHSH_B200_TRACE=1 timeout 300 python scripts/exp/hartfock_one.py batches 2>&1 | grep -v "Re file" | tail -40
