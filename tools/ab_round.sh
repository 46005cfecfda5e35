#!/bin/bash
# One gpurun call: GPU parity tests, then A/B timings of the matvec pipeline variants (env knobs), bench line,
# ncu launch list + full capture of the fused kernel.
# Usage: gpurun --timeout 1500 -- 'bash tools/ab_round.sh TAG [workload]'
TAG=${1:-r01b}
WL=${2:-kagome36}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -15 $OUT/pytest_gpu.log
fi
export LS_B200_PROFILE=1
run() { echo "== $1"; shift; env "$@" timeout 300 python tools/profile_workload.py $WL 3 2>&1 | tail -2; }
{
run "default (fused, two-level index)" A=1
run "fused, flat index" LS_B200_INDEX_FLAT=1
run "fused, no gather (orbit only)" LS_B200_MV_SKIP=2
run "unfused, two-level index" LS_B200_MATVEC=unfused
run "split" LS_B200_MATVEC=split
run "split, flat index" LS_B200_MATVEC=split LS_B200_INDEX_FLAT=1
for v in lattice_symmetries_b200/variants/*.so; do [ -f "$v" ] && run "variant $v split" LS_B200_LIBRARY=$PWD/$v LS_B200_MATVEC=split; done
} > $OUT/ab_$WL.txt 2>&1
cat $OUT/ab_$WL.txt
unset LS_B200_PROFILE
timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "bench exit $?"
cat $OUT/bench_$WL.json
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
  for K in orbit_gather_kernel row_sum_kernel; do
    timeout 500 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/${K}_$WL \
      python tools/profile_workload.py $WL 2 > $OUT/ncu_$K.log 2>&1
    tail -2 $OUT/ncu_$K.log
  done
fi
ls -la $OUT
