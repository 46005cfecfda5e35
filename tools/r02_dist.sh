#!/bin/bash
TAG=${1:-r02b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_distributed.py -q --timeout 600 > $OUT/pytest_dist.log 2>&1; echo "dist exit $?" >> $OUT/pytest_dist.log
tail -40 $OUT/pytest_dist.log
