#!/bin/bash
for i in 1 2 3 4 5 6 7 8; do timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "build_ranges: 4537\|enumerate:\|build_representatives\|\"build\"" | sed 's/.*"build": \({[^}]*}\).*/\1/' | cut -c1-260; echo; done
