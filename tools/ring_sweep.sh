#!/bin/bash
# sensitivity of k_prune_tc5 to its two TMA ring depths (inner-edge tiles / leaf tables); prints ms per 8 Mi columns
for cfg in "3 6" "2 6" "3 4" "2 3"; do
  set -- $cfg
  PCSF_TC5_NSTAGE=$1 PCSF_TC5_NLSTAGE=$2 timeout 60 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > /tmp/r.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("/tmp/r.json")); print("tile stages $1 leaf stages $2: ms_prune", round(d["stages_ms"]["ms_prune"],2), "columns/s", round(d["value"]))
PY
done
