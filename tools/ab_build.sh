#!/bin/bash
export LS_B200_PROFILE=1
run() { echo "== $1"; shift; env "$@" timeout 300 python tools/profile_workload.py ${WL:-kagome36} 1 2>&1 | grep "ls_b200\] enumerate\|ls_b200\] build_rep\|matvec ms"; }
run "two-phase K=8" A=1
run "two-phase K=4" LS_B200_BUILD_FILTER_ROWS=4
run "two-phase K=12" LS_B200_BUILD_FILTER_ROWS=12
run "two-phase K=16" LS_B200_BUILD_FILTER_ROWS=16
run "onepass" LS_B200_BUILD=onepass
