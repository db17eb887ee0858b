#!/bin/bash
timeout 300 python scripts/exp/hartfock_one.py
for c in 1 8; do PHSH_B200_HARTFOCK_CLUSTER=$c timeout 300 python scripts/exp/hartfock_one.py | tail -1; done
