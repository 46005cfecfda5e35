#!/bin/bash
# ncu --set full of named kernels, exported to CSV on the box (the .ncu-rep files are too large to travel):
#   tools/ncu_kernels.sh TAG WORKLOAD "ENV=.. ENV=.." kernel_regex...
TAG=$1; WL=$2; ENVS=$3; shift 3
OUT=gpurun_out/$TAG; mkdir -p $OUT
for K in "$@"; do
  REP=/tmp/${K}_$WL.ncu-rep
  env $ENVS timeout 500 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o ${REP%.ncu-rep} \
    python tools/profile_workload.py $WL 2 > $OUT/ncu_$K.log 2>&1
  tail -2 $OUT/ncu_$K.log
  ncu -i $REP --page details --csv > $OUT/${K}_$WL.details.csv 2>/dev/null
  ncu -i $REP --page raw --csv > $OUT/${K}_$WL.raw.csv 2>/dev/null
  ncu -i $REP --page source --csv --print-source sass > $OUT/${K}_$WL.sass.csv 2>/dev/null
done
ls -la $OUT
