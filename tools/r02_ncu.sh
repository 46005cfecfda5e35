#!/bin/bash
# ncu evidence of the round: launch list of a short bench run + --set full of the hot kernels (CSV exports, summarised)
TAG=${1:-r02r}; WL=${2:-kagome36}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "bench exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-checks > $OUT/ncu_launches.log 2>&1
for K in ${KERNELS:-orbit_kernel rank_gather_kernel row_sum_kernel}; do
  REP=/tmp/${K}_$WL
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:^$K -s 1 -c 1 -f -o $REP \
      python tools/profile_workload.py $WL 2 > $OUT/ncu_$K.log 2>&1
  tail -2 $OUT/ncu_$K.log
  ncu -i $REP.ncu-rep --page details --csv > $OUT/${K}_$WL.details.csv 2>/dev/null
  ncu -i $REP.ncu-rep --page raw --csv > $OUT/${K}_$WL.raw.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/${K}_$WL > $OUT/${K}_$WL.summary.txt 2>&1
  rm -f $OUT/${K}_$WL.raw.csv
done
python - <<PY
import json
d = json.loads(open("$OUT/bench_$WL.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k: round(v["ms_per_step"], 2) for k, v in d["roofline"]["kernels"].items()})
PY
ls -la $OUT
for W in $OTHER_WORKLOADS; do
  timeout 200 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$W.json 2> $OUT/bench_$W.err; echo "bench $W exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$W.json").read().strip().splitlines()[-1])
    print("$W", d["config"]["dim"], d["ms_per_step"], d["value"], d["dtype"], "e2e", (d["e2e"] or {}).get("ms_per_step"), d["checks"].get("sampled_rows"), d["checks"].get("hermiticity_rel_err"), d["roofline"]["bound"], round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$W: no line", e)
PY
done
