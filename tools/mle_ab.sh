#!/bin/bash
# A/B of library variants on the config-5 leg (MLE): alignments/s and per-kernel rooflines.  usage: tools/mle_ab.sh <variant|-> ...
for v in "$@"; do
  if [ "$v" = "-" ]; then unset PCSF_LIB_VARIANT; else export PCSF_LIB_VARIANT=$v; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 3 --cols 2097152 --config4-cols 0 --config5-alignments 262144 --cli-cols 0 > gpurun_out/mle_$v.json 2> gpurun_out/mle_$v.err || tail -n 3 gpurun_out/mle_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/mle_$v.json'))["config5"]
r=d["roofline"]
print('variant %-8s %.0f alignments/s  expm %.1f ms (%.2f of DMMA peak)  prune %.1f ms (%.2f)' % ('$v', d['alignments_per_s'], r['k_mle_expm']['ms'], r['k_mle_expm']['frac'], r['k_prune<true>']['ms'], r['k_prune<true>']['frac']))
PY
done
