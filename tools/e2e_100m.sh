#!/bin/bash
# BASELINE config 3 at full size through the command line host: 100 M reference columns of the synthetic hg38.100way-shaped MAF
# (58mammals, 30 % missing cells) as four 25 M-column chromosomes -> 7 wig files on one B200; then the tcgen05 tracks against the FP64
# tracks on one chromosome (largest |difference| of every emitted value).
# usage: tools/e2e_100m.sh [columns per chromosome] [chromosomes] [out.json]
set -e
N=${1:-25000000}; C=${2:-4}; OUT=${3:-gpurun_out/e2e_100m.json}
D=$(mktemp -d /tmp/pcsf100m.XXXX)
df -h /tmp | tail -1
t0=$(date +%s.%N)
python - <<PY
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from make_synth_maf import write_synth_maf
from phylocsfpp_b200.models import load_model
print(write_synth_maf("$D/chr1.maf", load_model("58mammals"), $N, seed=7))
PY
for c in $(seq 2 $C); do sed "s/\.chr1 /.chr$c /" $D/chr1.maf > $D/chr$c.maf; done
t1=$(date +%s.%N)
ls -la $D | head; cat $D/*.maf > /dev/null
FILES=$(ls $D/chr*.maf | sort -V | tr '\n' ' ')
PCSF_HOST_STATS=1 phylocsfpp_b200/bin/phylocsf_b200 build-tracks --threads $(nproc) --precision tc5 --output $D/o_tc5 58mammals $FILES | grep "^{" > $D/tc5.json
cat $D/tc5.json
PCSF_HOST_STATS=1 phylocsfpp_b200/bin/phylocsf_b200 build-tracks --threads $(nproc) --precision f64 --output $D/o_f64 58mammals $D/chr1.maf | grep "^{" > $D/f64.json
cat $D/f64.json
PCSF_HOST_STATS=1 phylocsfpp_b200/bin/phylocsf_b200 build-tracks --threads $(nproc) --precision tc5 --output $D/o_tc5_1 58mammals $D/chr1.maf | grep "^{" > $D/tc5_1.json
# largest difference between the two precisions over every emitted value of chromosome 1 (headers must be identical)
for f in PhyloCSFRaw+1 PhyloCSFRaw+2 PhyloCSFRaw+3 PhyloCSFRaw-1 PhyloCSFRaw-2 PhyloCSFRaw-3 PhyloCSFpower; do
  paste $D/o_tc5_1/$f.wig $D/o_f64/$f.wig | awk -v f=$f -F'\t' 'BEGIN{m=0;n=0;h=0} { if ($1 ~ /^fixedStep/) { if ($1 != $2) h++; next } d=$1-$2; if (d<0) d=-d; if (d>m) m=d; n++ } END{printf("%s values %d max_abs_diff %.4f header_mismatches %d\n", f, n, m, h)}'
done | tee $D/diff.txt
python - <<PY
import json
tc5=json.load(open("$D/tc5.json")); f64=json.load(open("$D/f64.json")); tc51=json.load(open("$D/tc5_1.json"))
diff=[l.split() for l in open("$D/diff.txt")]
json.dump({"config": "BASELINE config 3: build-tracks 58mammals, synthetic hg38.100way-shaped MAF, $C chromosomes x $N columns, 30% missing cells, 1 B200",
           "generate_seconds": $t1 - $t0, "build_tracks_tc5_all": tc5, "build_tracks_f64_chr1": f64, "build_tracks_tc5_chr1": tc51,
           "tc5_vs_f64_chr1": {d[0]: {"values": int(d[2]), "max_abs_diff": float(d[4]), "header_mismatches": int(d[6])} for d in diff}}, open("$OUT", "w"), indent=1)
print(open("$OUT").read()[:1500])
PY
rm -rf $D
