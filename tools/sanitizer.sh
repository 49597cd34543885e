#!/bin/bash
# compute-sanitizer leg (the analogue of the reference's ASAN CI leg, azure-pipelines.yml:9-14): memcheck and racecheck over smoke()
# (FP64 + tcgen05 build-tracks paths on 300 columns, checked against the oracle) and over a small score-msa batch (FIXED, MLE, OMEGA).
# usage: tools/sanitizer.sh <out.txt>
OUT=${1:-gpurun_out/sanitizer.txt}
: > $OUT
cat > /tmp/san_msa.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from phylocsfpp_b200 import capi
from phylocsfpp_b200.models import load_model
model = load_model("29mammals", "Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat")
rng = np.random.default_rng(3)
alns = [rng.choice(np.frombuffer(b"ACGTN-", np.uint8), size=(model.nl, L), p=[.22, .22, .22, .22, .06, .06]).astype(np.uint8) for L in (30, 61, 93)]
dm = capi.DeviceModel(model, 0)
for st in (capi.STRATEGY_FIXED, capi.STRATEGY_MLE, capi.STRATEGY_OMEGA):
    p, a, b = dm.score_msa(alns, st)
    print(st, p, b)
dm.close()
PY
for tool in memcheck racecheck; do
  for what in "python __graft_entry__.py --smoke" "python /tmp/san_msa.py"; do
    echo "=== compute-sanitizer --tool $tool $what" | tee -a $OUT
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 $what 2>&1 | grep -v "^$" | (head -30; echo "..."; tail -6) >> $OUT
    echo "rc=$?" >> $OUT
  done
done
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|===" $OUT
