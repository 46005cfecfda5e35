#!/bin/bash
TAG=${1:-ab}; WL=${2:-kagome36}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export LS_B200_PROFILE=1
run() { echo "== $1"; shift; env "$@" timeout 300 python tools/profile_workload.py $WL 3 2>&1 | tail -1; }
{
run "split" LS_B200_MATVEC=split
run "split prefix 12" LS_B200_MATVEC=split LS_B200_INDEX_PREFIX=12
run "fused" LS_B200_MATVEC=fused
for v in lattice_symmetries_b200/variants/*.so; do [ -f "$v" ] && run "variant $v split" LS_B200_LIBRARY=$PWD/$v LS_B200_MATVEC=split;  done
for v in lattice_symmetries_b200/variants/*.so; do [ -f "$v" ] && run "variant $v split prefix 12" LS_B200_LIBRARY=$PWD/$v LS_B200_MATVEC=split LS_B200_INDEX_PREFIX=12;  done
} > $OUT/ab_split_$WL.txt 2>&1
cat $OUT/ab_split_$WL.txt
