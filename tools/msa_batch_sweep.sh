#!/bin/bash
# score-msa --strategy mle on ~100 k single-block alignments (config 5 shape) for several alignments-per-call batch sizes
python - <<PY
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from make_synth_maf import write_synth_maf
from phylocsfpp_b200.models import load_model
print(write_synth_maf("/tmp/blocks.maf", load_model("29mammals"), 19000000, seed=3, loguniform_blocks=(30, 600)))
PY
cat /tmp/blocks.maf > /dev/null
S=Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat
for b in 4096 16384 32768 131072; do
  PCSF_HOST_MSA_BATCH=$b PCSF_HOST_STATS=1 timeout 120 phylocsfpp_b200/bin/phylocsf_b200 score-msa --strategy mle --comp-anc 1 --species $S --output /tmp/o_mle 29mammals /tmp/blocks.maf | grep "^{" | sed "s/^/batch $b /"
done
md5sum /tmp/o_mle/blocks.maf.scores
