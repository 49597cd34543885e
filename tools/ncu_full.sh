#!/bin/bash
# usage: tools/ncu_full.sh <out-prefix> [kernel regex]  — one ncu --set full capture (with source counters) of one launch of a kernel in a short bench run
OUT=$1; K=${2:-k_prune_tc5}
timeout 280 ncu --set full --import-source on --clock-control none -k regex:$K -c 1 -f -o $OUT python bench.py --no-cpu-baseline --steps 1 --warmup 3 --cols 2097152 > $OUT.log 2>&1
echo "ncu rc=$?"
