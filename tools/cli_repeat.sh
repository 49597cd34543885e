#!/bin/bash
# repeats build-tracks --precision tc5 on a 10 M-column synthetic MAF with 1, 2 and 16 host threads; prints the tool statistics
python - <<PY
import sys, os
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from make_synth_maf import write_synth_maf
from phylocsfpp_b200.models import load_model
print(write_synth_maf("/tmp/t.maf", load_model("58mammals"), 10000000, seed=7))
PY
cat /tmp/t.maf > /dev/null
for t in 1 16; do
  for i in 1 2 3; do
    PCSF_HOST_TIMING=1 PCSF_HOST_STATS=1 timeout 60 phylocsfpp_b200/bin/phylocsf_b200 build-tracks --threads $t --precision tc5 --output /tmp/o_tc5 58mammals /tmp/t.maf | grep "^{" | sed "s/^/threads $t run $i /" 
  done
done
