#!/bin/bash
# sweeps the host's columns-per-library-call group size on the 10 M-column synthetic file (tool seconds: tc5, f64)
for g in 1048576 524288 262144; do
  PCSF_HOST_GROUP_COLS=$g timeout 100 python tools/e2e_cli_bench.py 10000000 0 1 2>/dev/null > /tmp/gs.json
  python - <<PY
import json
d=json.load(open("/tmp/gs.json"))
print("group", $g, d["build_tracks_tc5"]["tool_stats"]["seconds"], d["build_tracks_f64"]["tool_stats"]["seconds"], d["build_tracks_tc5"]["tool_stats"])
PY
done
