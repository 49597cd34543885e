#!/bin/bash
# round-2 closing run: the whole -m gpu suite, the default bench line (all legs), the reference arm
mkdir -p gpurun_out
T=${TAG:-final2}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref arm rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline'].get('traffic'), 'cpu', d.get('cpu_baseline',{}).get('value'))
print('stages', d['stages_ms'])
for k in ('config4','config5','cli'):
    v=d.get(k); print(k, {x:v[x] for x in v if x in ('columns_per_s','alignments_per_s','seconds','process_seconds','tool_seconds')} if v else None)
print(open('gpurun_out/${T}_bench_ref.json').read()[:300])
PY
