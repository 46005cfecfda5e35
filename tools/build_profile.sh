#!/bin/bash
export LS_B200_PROFILE=1
for W in "$@"; do echo "== $W"; timeout 300 python tools/profile_workload.py $W 1 2>&1 | grep -v "^$" | tail -6; done
