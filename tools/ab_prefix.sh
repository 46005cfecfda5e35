#!/bin/bash
TAG=${1:-ab}; WL=${2:-kagome36}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export LS_B200_PROFILE=1
run() { echo "== $1"; shift; env "$@" timeout 300 python tools/profile_workload.py $WL 3 2>&1 | tail -1; }
{
for P in 10 12 14 16 18 20; do
run "split prefix $P" LS_B200_MATVEC=split LS_B200_INDEX_PREFIX=$P
done
run "fused prefix 12" LS_B200_INDEX_PREFIX=12
run "fused prefix 16" LS_B200_INDEX_PREFIX=16
} > $OUT/ab_prefix_$WL.txt 2>&1
cat $OUT/ab_prefix_$WL.txt
