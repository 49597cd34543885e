#!/bin/bash
# the command-line tests + the command-line leg alone (bench.py --cli-cols ...), other legs skipped
mkdir -p gpurun_out
T=${TAG:-cli}
timeout 900 python -m pytest tests/test_host_cli.py tests/test_reference_build.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 3 --config4-cols 0 --config5-alignments 0 --cli-cols ${COLS:-100000000} > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(json.dumps(d.get('cli')))
PY
