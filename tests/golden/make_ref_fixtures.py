#!/usr/bin/env python3
"""Generates tests/golden/ref-generated/: small synthetic MAF inputs together with the outputs of the REFERENCE ITSELF
(oracle/_ref/phylocsf_ref = the reference's unmodified sources compiled against oracle/ref/gsl, see oracle/ref/Makefile)
run on them in the build container.  The GPU box has neither /root/reference nor a reason to trust a rebuilt binary, so
the inputs and the reference's outputs are committed (a few hundred KB, gzip) next to this script.

  python tests/golden/make_ref_fixtures.py            # needs oracle/_ref/phylocsf_ref (python __graft_entry__.py)

Fixtures (what each one pins that the reference's own goldens do not):
  tracks12   build-tracks, 12flies, 15 000 reference columns starting at 992 001: a chain that crosses the 1 Mb
             BREAKPOINT_POS (+2-base read-ahead, cursor rewind), holes, reference-gap columns, absent species, rows of an
             unknown species, soft-masked blocks -> 7 wig files
  smooth53   build-tracks --output-phylo 1 --output-regions 1 on the reference's own example MAF (53birds) with coding exons
             taken from example/galGal6_chr22_25_28_subset_ncbiRefSeq.gtf by the README's awk line (README.rst:124; the
             exon list its tests.sh names is not shipped) -> 6 smoothed wigs + 6 region BED files
  smooth12   the same on tracks12 with tests/util.py:write_synthetic_exons (46 000 exons: gap subsampling path)
  msa29      score-msa, 29mammals reduced with --species to 12 leaves, 60 single-block alignments of 30..600 columns
             (BASELINE config 5 shape), strategies fixed / mle (--comp-anc 1), omega and fixed_mean -> .scores files
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = os.path.join(ROOT, "oracle", "_ref", "phylocsf_ref")
OUT = os.path.join(HERE, "ref-generated")
SPECIES29 = "Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat"


def gz(src, dst):
    with open(src, "rb") as fi, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as fo:
        shutil.copyfileobj(fi, fo)


def ref(*args):
    subprocess.run([REF] + list(args), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def main():
    os.environ["PCSF_SYNTH_CPU"] = "1"
    from make_synth_maf import write_synth_maf
    from phylocsfpp_b200.models import load_model
    if not os.path.exists(REF):
        raise SystemExit(f"{REF} missing: build it with `make -C oracle/ref` (needs /root/reference)")
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        # ---- tracks12
        maf = os.path.join(tmp, "tracks12.maf")
        print(write_synth_maf(maf, load_model("12flies"), 15000, seed=11, start0=992000, mean_block=90, hole_p=1 / 40.0, ref_gap=0.02,
                              alien_p=0.05))
        ref("build-tracks", "--threads", "4", "--output", os.path.join(tmp, "t12"), "12flies", maf)
        gz(maf, os.path.join(OUT, "tracks12.maf.gz"))
        for n in ["PhyloCSFpower.wig"] + [f"PhyloCSFRaw{s}{f}.wig" for s in "+-" for f in (1, 2, 3)]:
            gz(os.path.join(tmp, "t12", n), os.path.join(OUT, "tracks12." + n + ".gz"))
        # ---- smooth53 / smooth12
        from tests.util import write_synthetic_exons
        G = os.path.join(HERE, "build-tracks")
        maf53 = os.path.join(tmp, "in53.maf")
        with gzip.open(os.path.join(G, "galGal6_chr22_25_28_each_30k_bases.maf.gz"), "rb") as fi, open(maf53, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        exons = os.path.join(tmp, "smooth53.coding_exons.txt")
        with open(exons, "w") as fo:
            for ln in open("/root/reference/example/galGal6_chr22_25_28_subset_ncbiRefSeq.gtf"):
                f = ln.rstrip("\n").split("\t")
                if len(f) > 7 and f[2] == "CDS":
                    fo.write("\t".join([f[0], f[6], f[7], f[3], f[4]]) + "\n")
        gz(exons, os.path.join(OUT, "smooth53.coding_exons.txt.gz"))
        ref("build-tracks", "--threads", "8", "--output-phylo", "1", "--output-regions", "1", "--genome-length", "1065365434", "--coding-exons", exons,
            "--output", os.path.join(tmp, "s53"), os.path.join(G, "53birds"), maf53)
        exons12 = os.path.join(tmp, "exons12.txt")
        write_synthetic_exons(exons12)
        ref("build-tracks", "--threads", "4", "--output-phylo", "1", "--output-regions", "1", "--genome-length", "400000000", "--coding-exons", exons12,
            "--output", os.path.join(tmp, "s12"), "12flies", os.path.join(tmp, "tracks12.maf"))
        for tag, d in (("smooth53", "s53"), ("smooth12", "s12")):
            for k in ("+1", "+2", "+3", "-1", "-2", "-3"):
                gz(os.path.join(tmp, d, f"PhyloCSF{k}.wig"), os.path.join(OUT, f"{tag}.PhyloCSF{k}.wig.gz"))
                gz(os.path.join(tmp, d, f"PhyloCSF{k}Regions.bed"), os.path.join(OUT, f"{tag}.PhyloCSF{k}Regions.bed.gz"))
        # ---- msa29
        maf = os.path.join(tmp, "msa29.maf")
        print(write_synth_maf(maf, load_model("29mammals"), 14000, seed=12, loguniform_blocks=(30, 600), alien_p=0.05))
        gz(maf, os.path.join(OUT, "msa29.maf.gz"))
        for strat, anc in (("fixed", "1"), ("mle", "1"), ("omega", "0")):
            ref("score-msa", "--threads", "8", "--strategy", strat, "--comp-phylo", "1", "--comp-anc", anc, "--species", SPECIES29,
                "--output", os.path.join(tmp, "m29_" + strat), "29mammals", maf)
            shutil.copy(os.path.join(tmp, "m29_" + strat, "msa29.maf.scores"), os.path.join(OUT, f"msa29.{strat}.scores"))
        # FIXED_MEAN: per-codon scores through the PhyloCSF-HMM (parameters from the synthetic exon list of smooth12)
        ref("score-msa", "--threads", "8", "--strategy", "fixed_mean", "--comp-phylo", "1", "--comp-anc", "0", "--species", SPECIES29,
            "--genome-length", "400000000", "--coding-exons", exons12, "--output", os.path.join(tmp, "m29_fm"), "29mammals", maf)
        shutil.copy(os.path.join(tmp, "m29_fm", "msa29.maf.scores"), os.path.join(OUT, "msa29.fixed_mean.scores"))
        # ---- the reference's own medium MLE input, scored by the reference's current sources: the shipped golden of that input
        # (test/maf-file-medium/...mle.scores) predates v1.2.0 — 135 of its 516 rows are not reproduced by the unmodified sources
        # either (same rows as on the GPU) — so the GPU test pins against this file and reports the shipped one (32 CPU-minutes)
        if os.environ.get("PCSF_FIXTURES_MLE516"):
            maf516 = os.path.join(tmp, "chr22.516alignments.maf")
            with gzip.open(os.path.join(HERE, "score-msa", "chr22.516alignments.maf.gz"), "rb") as fi, open(maf516, "wb") as fo:
                shutil.copyfileobj(fi, fo)
            ref("score-msa", "--threads", "16", "--strategy", "mle", "--comp-anc", "1", "--output", os.path.join(tmp, "m516"), "100vertebrates", maf516)
            shutil.copy(os.path.join(tmp, "m516", "chr22.516alignments.maf.scores"), os.path.join(OUT, "chr22.516alignments.mle.refbuilt.scores"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
