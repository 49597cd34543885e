#!/bin/bash
# round-2 session 3: the whole -m gpu suite + the default bench line (all legs) + the reference arm
mkdir -p gpurun_out
T=${TAG:-k}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${T}_rc.txt
tail -3 gpurun_out/${T}_pytest.log
( time timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2> gpurun_out/${T}_bench.time; echo "bench rc=$?" | tee -a gpurun_out/${T}_rc.txt
tail -5 gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.time
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
for k in ('config4','config5','cli'):
    v=d.get(k); print(k, json.dumps(v)[:1500] if v else None)
PY
