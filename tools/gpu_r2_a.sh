#!/bin/bash
# round 2, call A: parity of the cherry-table tcgen05 kernel, bench line, per-step trace, light ncu pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tracks.py -x -q -m gpu > gpurun_out/a_pytest_tracks.log 2>&1; echo "pytest tracks rc=$?" | tee -a gpurun_out/a_rc.txt
timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" | tee -a gpurun_out/a_rc.txt
timeout 120 python tools/tc5_trace_run.py > gpurun_out/a_trace.txt 2>&1; echo "trace rc=$?" | tee -a gpurun_out/a_rc.txt
M2=lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
timeout 170 ncu --metrics $M2 --clock-control none -k regex:k_prune_tc5 -c 1 --csv --log-file gpurun_out/a_ncu_l2.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 --cols 2097152 > /dev/null 2> gpurun_out/a_ncu_l2.err
tools/ncu_quick.sh gpurun_out/a_ncu_quick > gpurun_out/a_ncu_quick.txt 2>&1
tail -3 gpurun_out/a_pytest_tracks.log; cat gpurun_out/a_bench.json | head -c 1500
