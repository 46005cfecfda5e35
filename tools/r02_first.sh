#!/bin/bash
# first GPU pass of round 2: emulated-rank tests, the whole GPU suite, one bench line
TAG=${1:-r02a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py -q -x --timeout 600 > $OUT/pytest_dist.log 2>&1; echo "dist exit $?" >> $OUT/pytest_dist.log
tail -25 $OUT/pytest_dist.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_distributed.py > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_kagome36.json 2> $OUT/bench_kagome36.err; echo "bench exit $?"
cat $OUT/bench_kagome36.json; tail -5 $OUT/bench_kagome36.err
