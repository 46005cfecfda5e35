#!/bin/bash
TAG=${1:-ab}; WL=${2:-kagome36}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export LS_B200_PROFILE=1
run() { echo "== $1"; shift; env "$@" timeout 300 python tools/profile_workload.py $WL 3 2>&1 | tail -1; }
{
run "split" A=1
run "split, no xs gather" LS_B200_MV_SKIP=4
run "split, no index search" LS_B200_MV_SKIP=8
run "split, neither" LS_B200_MV_SKIP=12
} > $OUT/ab_skip_$WL.txt 2>&1
cat $OUT/ab_skip_$WL.txt
