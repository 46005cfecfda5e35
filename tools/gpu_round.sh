#!/bin/bash
# One gpurun call: GPU parity tests, bench line, reference arm, ncu launch list, ncu --set full of the hot kernels
# (exported to CSV on the box; the .ncu-rep files are too large to travel), bench lines of the other workloads.
# Usage: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh TAG [workload]'
TAG=${1:-r01}
WL=${2:-kagome36}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "bench exit $?"
cat $OUT/bench_$WL.json
if [ -z "$SKIP_REF" ]; then
  timeout 400 python bench.py --impl reference --workload $WL --steps 3 --warmup 1 > $OUT/bench_ref_$WL.json 2> $OUT/bench_ref_$WL.err
  cat $OUT/bench_ref_$WL.json
fi
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
  for K in orbit_kernel rank_gather_kernel row_sum_kernel build_flags_bitsliced_kernel; do
    REP=/tmp/${K}_$WL
    timeout 500 ncu --set full --clock-control none --import-source on -k regex:^$K -s 1 -c 1 -f -o $REP \
      python tools/profile_workload.py $WL 2 > $OUT/ncu_$K.log 2>&1
    tail -1 $OUT/ncu_$K.log
    ncu -i $REP.ncu-rep --page details --csv > $OUT/${K}_$WL.details.csv 2>/dev/null
    ncu -i $REP.ncu-rep --page raw --csv > $OUT/${K}_$WL.raw.csv 2>/dev/null
  done
fi
for W in $OTHER_WORKLOADS; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$W.json 2> $OUT/bench_$W.err; echo "bench $W exit $?"
  cat $OUT/bench_$W.json
done
ls -la $OUT
