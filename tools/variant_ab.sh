#!/bin/bash
# A/B of library variants on one box: resident columns/s of the config-3 step (bench.py, other legs off), two rounds
# usage: tools/variant_ab.sh <variant|-> ...     ('-' = the product library)
mkdir -p gpurun_out
for round in 1 2; do
for v in "$@"; do
  if [ "$v" = "-" ]; then unset PCSF_LIB_VARIANT; else export PCSF_LIB_VARIANT=$v; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 6 --warmup 3 --config4-cols 0 --config5-alignments 0 --cli-cols 0 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$v.json'))
print('variant %-8s round $round value %.2f M col/s  prune %.2f ms  e2e %.2f M  clocks %s' % ('$v', d['value']/1e6, d['stages_ms']['ms_prune'], d['e2e']['value']/1e6, d['clocks']['sm_mhz']))
PY
done
done
