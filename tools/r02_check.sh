#!/bin/bash
TAG=${1:-r02q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
if [ -z "$SKIP_BENCH" ]; then
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench_kagome36.json 2> $OUT/bench_kagome36.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("$OUT/bench_kagome36.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "build", d["build"]["samples_ms"], d["checks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print({k: (round(v["ms_per_step"], 2), round(v.get("int_frac", 0), 3), round(v["hbm_frac"], 3)) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["bound"], d["roofline"]["frac"])
PY
fi
