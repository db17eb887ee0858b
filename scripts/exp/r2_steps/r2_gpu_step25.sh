#!/bin/bash
timeout 300 python scripts/exp/hartfock_one.py 2>&1 | tail -6
