#!/bin/bash
# One 8-GPU call: real-rank parity tests, the kagome-36 bench at N=8 (both product forms), then kagome-42:
# sharded build + products + Lanczos ground state under a time limit.  Each part under its own timeout.
TAG=${1:-r02g}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt; free -g >> $OUT/gpu.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_TESTS" ]; then
  LS_B200_DIST_MIN_BLOCK=65536 timeout 300 $RUN --master-port 29601 tests/nccl_worker.py > $OUT/nccl_worker.log 2>&1; echo "nccl_worker exit $?"
  grep -E "NCCL_WORKER|FAILED" $OUT/nccl_worker.log | head
fi
for M in ${MODES:-auto alltoall}; do
  timeout 300 $RUN --master-port 29602 bench.py --gpus $N --steps 10 --warmup 3 --workload kagome36 --mode $M \
     > $OUT/bench_kagome36_${N}gpu_$M.json 2> $OUT/bench_kagome36_${N}gpu_$M.err; echo "bench kagome36 $M exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_kagome36_${N}gpu_$M.json").read().strip().splitlines()[-1])
    print("$M", d["ms_per_step"], "ms/step; build", d["build"]["samples_ms"], "e2e", d["e2e"]["ms_per_step"], d["checks"].get("sampled_rows_max_rel_err"))
except Exception as e:
    print("no bench line:", e)
PY
  grep dist_build $OUT/bench_kagome36_${N}gpu_$M.err | tail -2
done
if [ -n "$K42" ]; then
  timeout ${K42_TIMEOUT:-900} $RUN --master-port 29603 tools/ground_state.py kagome42 --mode ${K42_MODE:-auto} --matvecs 3 \
     --time-limit ${K42_LANCZOS_S:-300} --tol 1e-9 --max-iters 300 --out $OUT/kagome42.json ${K42_ARGS} > $OUT/kagome42.log 2>&1; echo "kagome42 exit $?"
  grep -v "^\[W\|^W0\|^\*\*\*" $OUT/kagome42.log | tail -40
fi
