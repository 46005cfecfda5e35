#!/bin/bash
TAG=${1:-r02i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "sorted or variants" > $OUT/pytest_sorted.log 2>&1; tail -5 $OUT/pytest_sorted.log
for S in 0 1; do
  LS_B200_MV_SORT=$S timeout 600 python tools/footprint_probe.py ${SITES:-40} > $OUT/probe_plain_sort$S.log 2>&1; grep -v build_ranges $OUT/probe_plain_sort$S.log | tail -6
  LS_B200_MV_SORT=$S timeout 600 python tools/footprint_probe.py ${SITES:-40} --wide > $OUT/probe_wide_sort$S.log 2>&1; grep -v build_ranges $OUT/probe_wide_sort$S.log | tail -6
done
