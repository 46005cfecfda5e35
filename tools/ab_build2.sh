#!/bin/bash
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --workload ${WL:-kagome36} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -o '"build": {[^}]*}'; }
run "two-phase K=8" A=1
run "two-phase K=8 again" A=1
run "two-phase K=4" LS_B200_BUILD_FILTER_ROWS=4
run "two-phase K=16" LS_B200_BUILD_FILTER_ROWS=16
run "onepass" LS_B200_BUILD=onepass
run "onepass again" LS_B200_BUILD=onepass
