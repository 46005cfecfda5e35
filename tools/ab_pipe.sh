#!/bin/bash
TAG=${1:-ab}; WL=${2:-kagome36}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export LS_B200_PROFILE=1
run() { echo "== $1"; shift; env "$@" timeout 300 python tools/profile_workload.py $WL 3 2>&1 | tail -1; }
{
run "split sequential" LS_B200_MV_PIPELINE=0
run "split pipelined, B high priority" LS_B200_MV_PIPELINE=1 LS_B200_MV_PRIORITY=1
run "split pipelined, B low priority" LS_B200_MV_PIPELINE=1 LS_B200_MV_PRIORITY=0
} > $OUT/ab_pipe_$WL.txt 2>&1
cat $OUT/ab_pipe_$WL.txt
