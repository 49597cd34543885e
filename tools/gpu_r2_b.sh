#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-b}
timeout 600 python -m pytest tests/test_gpu_tracks.py -x -q -m gpu > gpurun_out/${T}_pytest_tracks.log 2>&1; echo "pytest tracks rc=$?" | tee gpurun_out/${T}_rc.txt
timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${T}_rc.txt
timeout 120 python tools/tc5_trace_run.py > gpurun_out/${T}_trace.txt 2>&1; echo "trace rc=$?" | tee -a gpurun_out/${T}_rc.txt
tools/ncu_quick.sh gpurun_out/${T}_ncu_quick > gpurun_out/${T}_ncu_quick.txt 2>&1
tail -3 gpurun_out/${T}_pytest_tracks.log; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'stages', d['stages_ms'], 'clocks', d['clocks'])"
grep -E "tensor|issue_active|time_duration|wavefronts" gpurun_out/${T}_ncu_quick.txt
