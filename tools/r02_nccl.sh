#!/bin/bash
# N-GPU pass: real NCCL ranks (parity vs the oracle), then the bench at N GPUs
TAG=${1:-r02c}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 600 python -m pytest tests/test_gpu_distributed.py -q --timeout 600 -k "nccl or complex or invalid" > $OUT/pytest_nccl.log 2>&1; echo "exit $?" >> $OUT/pytest_nccl.log
tail -30 $OUT/pytest_nccl.log
for W in ${WORKLOADS:-kagome36}; do
  for M in ${MODES:-auto}; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 5 --warmup 3 --workload $W --mode $M > $OUT/bench_${W}_${N}gpu_$M.json 2> $OUT/bench_${W}_${N}gpu_$M.err; echo "bench $W $M exit $?"
    tail -1 $OUT/bench_${W}_${N}gpu_$M.json | cut -c 1-1800; tail -4 $OUT/bench_${W}_${N}gpu_$M.err
  done
done
