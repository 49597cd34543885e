#!/bin/bash -e
# Copies the reference's own golden inputs/outputs for the hot path into tests/golden/.
# Run in the build container only (needs /root/reference); the copies are committed because
# /root/reference does not exist on the GPU box.  These are data fixtures, not source code.
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/build-tracks" "$HERE/score-msa"
# config 1: build-tracks on the example MAF with the external 53birds model (test/tests.sh:15-19)
cp "$REF/example/galGal6_chr22_25_28_each_30k_bases.maf.gz" "$HERE/build-tracks/"
cp "$REF"/test/53birds.nh "$REF"/test/53birds_coding.ECM "$REF"/test/53birds_noncoding.ECM "$HERE/build-tracks/"
for f in PhyloCSFRaw+1 PhyloCSFRaw+2 PhyloCSFRaw+3 PhyloCSFRaw-1 PhyloCSFRaw-2 PhyloCSFRaw-3 PhyloCSFpower; do
  gzip -9 -n -c "$REF/test/expected_results/build-tracks/$f.wig" > "$HERE/build-tracks/$f.wig.gz"
done
# config 2: score-msa goldens (test/tests.sh:35-42 and the 516-alignment set)
cp "$REF/test/maf-file-small/chr22.50alignments.maf" "$HERE/score-msa/"
cp "$REF"/test/maf-file-small/PhyloCSFpp-results/chr22.50alignments.fixed.scores "$HERE/score-msa/"
cp "$REF"/test/maf-file-small/PhyloCSFpp-results/chr22.50alignments.mle.scores "$HERE/score-msa/"
gzip -9 -n -c "$REF/test/maf-file-medium/chr22.516alignments.maf" > "$HERE/score-msa/chr22.516alignments.maf.gz"
cp "$REF"/test/maf-file-medium/chr22.516alignments.maf.fixed.scores "$HERE/score-msa/"
cp "$REF"/test/maf-file-medium/chr22.516alignments.maf.mle.scores "$HERE/score-msa/"
chmod -R u+w "$HERE"
# f-4: annotate-with-tracks (test/tests.sh:23-26): the example tracks, the three example annotations and the expected output
mkdir -p "$HERE/annotate-with-tracks"
cp "$REF"/example/tracks/PhyloCSF*.bw "$HERE/annotate-with-tracks/"
for f in ensGene ncbiRefSeq refGene; do
  gzip -9 -n -c "$REF/example/galGal6_chr22_25_28_subset_$f.gtf" > "$HERE/annotate-with-tracks/galGal6_chr22_25_28_subset_$f.gtf.gz"
  gzip -9 -n -c "$REF/test/expected_results/annotate-with-tracks/galGal6_chr22_25_28_subset_$f.PhyloCSF++.gtf" > "$HERE/annotate-with-tracks/galGal6_chr22_25_28_subset_$f.PhyloCSF++.gtf.gz"
done
chmod -R u+w "$HERE"
