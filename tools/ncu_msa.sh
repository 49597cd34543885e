#!/bin/bash
# ncu --set full captures of the score-msa kernels in the middle of a run: MLE (k_mle_step, k_mle_plan, k_mle_expm, k_prune<true>) on
# 16384 alignments, OMEGA (k_omega_*) on 2048.   usage: tools/ncu_msa.sh <out-prefix>
OUT=${1:-gpurun_out/ncu_msa}
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_mle_expm|k_prune" --launch-skip 80 --launch-count 4 -f -o ${OUT}_mle python tools/ncu_msa.py mle 16384 > ${OUT}_mle.log 2>&1; echo "mle rc=$?"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_omega" --launch-skip 40 --launch-count 6 -f -o ${OUT}_omega python tools/ncu_msa.py omega 2048 > ${OUT}_omega.log 2>&1; echo "omega rc=$?"
tail -n 2 ${OUT}_mle.log; tail -n 2 ${OUT}_omega.log
