#!/bin/bash
# A/B of the per-handle stream (default) against the legacy default stream (PCSF_LEGACY_STREAM=1): build-tracks --precision tc5 on a
# 25 M-column synthetic MAF, 16 host threads, three runs each
python - <<PY
import sys, os
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from make_synth_maf import write_synth_maf
from phylocsfpp_b200.models import load_model
print(write_synth_maf("/tmp/t.maf", load_model("58mammals"), 25000000, seed=7))
PY
cat /tmp/t.maf > /dev/null
for i in 1 2 3; do
  PCSF_HOST_STATS=1 timeout 60 phylocsfpp_b200/bin/phylocsf_b200 build-tracks --threads 16 --precision tc5 --output /tmp/o_tc5 58mammals /tmp/t.maf | grep "^{" | sed "s/^/own-stream run $i /" | cut -c1-260
  PCSF_LEGACY_STREAM=1 PCSF_HOST_STATS=1 timeout 60 phylocsfpp_b200/bin/phylocsf_b200 build-tracks --threads 16 --precision tc5 --output /tmp/o_tc5 58mammals /tmp/t.maf | grep "^{" | sed "s/^/legacy     run $i /" | cut -c1-260
done
