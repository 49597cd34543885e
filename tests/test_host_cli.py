"""The C++ host (phylocsfpp_b200/host -> bin/phylocsf_b200): its MAF reader against the Python mirror on the reference's
fixtures and on a synthetic file with breakpoints, holes, reference gaps and unknown species (CPU), and the two tools end
to end against the reference's golden outputs (GPU)."""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from phylocsfpp_b200.maf import MafReader
from phylocsfpp_b200.models import load_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phylocsfpp_b200", "bin", "phylocsf_b200")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _need_bin():
    if not os.path.exists(BIN):
        pytest.fail(f"{BIN} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")


def _fnv(a: np.ndarray) -> str:
    h = 1469598103934665603
    for b in a.reshape(-1).tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _dump(model_arg, maf, concatenate, threads, species=""):
    cmd = [BIN, "dump-alignments", "--concatenate", "1" if concatenate else "0", "--threads", str(threads)]
    if species:
        cmd += ["--species", species]
    out = subprocess.run(cmd + [model_arg, maf], check=True, capture_output=True, text=True).stdout
    return [ln.split("\t") for ln in out.splitlines()]


def _gunzip(src, tmp_path):
    dst = os.path.join(tmp_path, os.path.basename(src)[:-3])
    with gzip.open(src, "rb") as fi, open(dst, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return dst


def _compare(rows, alns, hash_limit=200000):
    alns = [a for a in alns]
    assert len(rows) == len(alns)
    for r, a in zip(rows, alns):
        assert r[0] == a.chrom and int(r[1]) == a.start_pos and int(r[2]) == a.chrom_len and r[3] == a.strand and int(r[4]) == a.L
        if a.seqs.size <= hash_limit:
            assert r[5] == _fnv(a.seqs)


def test_reader_matches_mirror_on_reference_fixtures(golden_dir, tmp_path):
    _need_bin()
    G = os.path.join(golden_dir, "build-tracks")
    maf = _gunzip(os.path.join(G, "galGal6_chr22_25_28_each_30k_bases.maf.gz"), str(tmp_path))
    model = load_model(os.path.join(G, "53birds"))
    ref = list(MafReader(maf, model.seqid_to_phyloid, model.nl, True, warn=False))
    for threads in (1, 3, 8):
        _compare(_dump(os.path.join(G, "53birds"), maf, True, threads), ref, hash_limit=10 ** 7 if threads == 1 else 0)
    S = os.path.join(golden_dir, "score-msa")
    model = load_model("100vertebrates")
    maf = os.path.join(S, "chr22.50alignments.maf")
    ref = list(MafReader(maf, model.seqid_to_phyloid, model.nl, False, warn=False))
    _compare(_dump("100vertebrates", maf, False, 4), ref)


def test_reader_synthetic_breakpoints_holes_gaps(tmp_path):
    """2.3 M reference bases starting at 10 000: two 1 Mb breakpoints (the +2-base read-ahead and the rewind), a few holes,
    1 % reference-gap columns, absent species, a --species reduction with rows of species outside the subset."""
    _need_bin()
    from make_synth_maf import write_synth_maf
    model = load_model("12flies")
    maf = os.path.join(str(tmp_path), "synth.maf")
    info = write_synth_maf(maf, model, 2_300_000, seed=5, hole_p=1 / 4000.0)
    assert info["blocks"] > 10000
    ref = list(MafReader(maf, model.seqid_to_phyloid, model.nl, True, warn=False))
    assert any(a.L > 900000 for a in ref) and len(ref) > 5
    for threads in (1, 5):
        _compare(_dump("12flies", maf, True, threads), ref, hash_limit=3_000_000 if threads == 5 else 0)
    sub = "dmel,dsim,dyak,dpse"
    msub = load_model("12flies", sub)
    ref = list(MafReader(maf, msub.seqid_to_phyloid, msub.nl, False, warn=False))
    rows = _dump("12flies", maf, False, 4, species=sub)
    ref = [a for a in ref if a.L > 0 or a.chrom]
    _compare(rows[:2000], ref[:2000])


def test_reader_odd_line_formats(tmp_path):
    """The fast line path of read_chain_into (species + text taken without tokenising the fields in between; species of line k
    compared with line k of the previous block) against the full tokeniser and the Python mirror on lines a strtok reader accepts
    but a careless one trips over: runs of spaces between fields, trailing spaces, an eighth token, a longer first
    token, `i` / `e` / `q` lines, species order changing from block to block, upper-case species, bare `a` lines."""
    _need_bin()
    import random
    from make_synth_maf import write_synth_maf
    model = load_model("12flies")
    plain = os.path.join(str(tmp_path), "plain.maf")
    write_synth_maf(plain, model, 60_000, seed=11, hole_p=1 / 2000.0)
    rnd = random.Random(3)
    out = []
    block = []

    def flush():
        if block:
            head, rows = block[0], block[1:]
            if len(rows) > 2 and rnd.random() < 0.3:          # species order changes (the reference row stays first)
                first, rest = rows[0], rows[1:]
                rnd.shuffle(rest)
                rows = [first] + rest
            out.append(head)
            out.extend(rows)
            block.clear()
    for ln in open(plain).read().split("\n"):
        if ln.startswith("a"):
            flush()
            block.append("a" if rnd.random() < 0.2 else ln)
        elif ln.startswith("s "):
            tok = ln.split()
            k = rnd.random()
            if k < 0.15:
                ln = "s  " + tok[1] + "   " + "  ".join(tok[2:6]) + "    " + tok[6]          # runs of spaces
            elif k < 0.30:
                ln = ln + "  "                                                               # trailing spaces
            elif k < 0.45:
                ln = tok[0] + " " + tok[1] + " " + " ".join(tok[2:6]) + " " + tok[6] + " "   # single trailing space
            elif k < 0.55:
                ln = ln + " extra"                                                           # an eighth token (ignored by strtok readers)
            elif k < 0.60:
                ln = "sx" + ln[1:]                                                           # first token longer than one character
            elif k < 0.70:
                sp, rest = tok[1].split(".", 1)
                ln = " ".join([tok[0], sp.upper() + "." + rest] + tok[2:])                  # upper-case species
            block.append(ln)
            if rnd.random() < 0.1:
                block.append("i " + tok[1] + " C 0 C 0")
            if rnd.random() < 0.05:
                block.append("q " + tok[1] + " " + "9" * len(tok[6]))
        elif ln == "":
            flush()
            out.append("")
        else:
            flush()
            out.append(ln)
    flush()
    odd = os.path.join(str(tmp_path), "odd.maf")
    open(odd, "w").write("\n".join(out) + "\n")
    ref = list(MafReader(odd, model.seqid_to_phyloid, model.nl, True, warn=False))
    base = list(MafReader(plain, model.seqid_to_phyloid, model.nl, True, warn=False))
    assert len(ref) == len(base) and all(a.L == b.L and np.array_equal(a.seqs, b.seqs) for a, b in zip(ref, base))
    for threads in (1, 4):          # dump-alignments itself dies if read_chain_into and read_chain disagree on a chain
        _compare(_dump("12flies", odd, True, threads), ref, hash_limit=10 ** 7)
    env = dict(os.environ, PCSF_REQUIRE_DIRECT="1")          # and none of these lines may push a chain onto the slow path
    subprocess.run([BIN, "dump-alignments", "--hash", "0", "--threads", "2", "12flies", odd], check=True, capture_output=True, env=env)


def test_wig_number_formatting_matches_printf():
    """my_fprintf (reference src/common.hpp:48-68) re-implemented without snprintf: identical text on 2 M pseudo-random
    floats, exact rounding ties and special values."""
    _need_bin()
    r = subprocess.run([BIN, "format-selftest", "1000000"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("0 mismatches"), r.stdout[-2000:]


FRAMES6 = ["+1", "+2", "+3", "-1", "-2", "-3"]


def _smooth_inputs(golden_dir, tmp_path, tag):
    """raw-track directory + exon list + genome length of the two smoothing fixtures of tests/golden/ref-generated/."""
    from tests.util import write_synthetic_exons
    R = os.path.join(golden_dir, "ref-generated")
    raw = os.path.join(str(tmp_path), "raw_" + tag)
    os.makedirs(raw, exist_ok=True)
    exons = os.path.join(str(tmp_path), tag + ".exons.txt")
    if tag == "smooth53":
        for k in FRAMES6:
            _gunzip(os.path.join(golden_dir, "build-tracks", f"PhyloCSFRaw{k}.wig.gz"), raw)
        with gzip.open(os.path.join(R, "smooth53.coding_exons.txt.gz"), "rb") as fi, open(exons, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        return raw, exons, "1065365434"
    for k in FRAMES6:
        with gzip.open(os.path.join(R, f"tracks12.PhyloCSFRaw{k}.wig.gz"), "rb") as fi, open(os.path.join(raw, f"PhyloCSFRaw{k}.wig"), "wb") as fo:
            shutil.copyfileobj(fi, fo)
    write_synthetic_exons(exons)
    return raw, exons, "400000000"


def _assert_smoothed(out, golden_dir, tag):
    R = os.path.join(golden_dir, "ref-generated")
    for k in FRAMES6:
        assert open(os.path.join(out, f"PhyloCSF{k}.wig"), "rb").read() == gzip.open(os.path.join(R, f"{tag}.PhyloCSF{k}.wig.gz"), "rb").read(), k
        assert open(os.path.join(out, f"PhyloCSF{k}Regions.bed"), "rb").read() == gzip.open(os.path.join(R, f"{tag}.PhyloCSF{k}Regions.bed.gz"), "rb").read(), k


@pytest.mark.parametrize("tag", ["smooth53", "smooth12"])
def test_hmm_smoothing_matches_reference(golden_dir, tmp_path, tag):
    """PhyloCSF-HMM stage alone (host/hmm.hpp; HMM parameter estimation from coding exons, scaled forward-backward, Viterbi
    regions) on raw tracks written by the reference: smoothed wigs and region BEDs byte-identical to what the reference
    itself wrote (tests/golden/make_ref_fixtures.py).  smooth12 has > 20 000 gaps: the std::shuffle subsampling path."""
    _need_bin()
    raw, exons, glen = _smooth_inputs(golden_dir, tmp_path, tag)
    out = os.path.join(str(tmp_path), "out_" + tag)
    subprocess.run([BIN, "smooth-tracks", "--genome-length", glen, "--coding-exons", exons, raw, out], check=True, capture_output=True)
    _assert_smoothed(out, golden_dir, tag)


@pytest.mark.gpu
def test_build_tracks_cli_smoothing(golden_dir, tmp_path):
    """build-tracks --output-phylo 1 --output-regions 1 --output-raw-phylo 0 end to end (CUDA likelihoods -> raw text -> HMM):
    identical to the reference's files, raw tracks removed afterwards (build_tracks.hpp:343-346)."""
    _need_bin()
    G = os.path.join(golden_dir, "build-tracks")
    maf = _gunzip(os.path.join(G, "galGal6_chr22_25_28_each_30k_bases.maf.gz"), str(tmp_path))
    _, exons, glen = _smooth_inputs(golden_dir, tmp_path, "smooth53")
    out = os.path.join(str(tmp_path), "out_s")
    subprocess.run([BIN, "build-tracks", "--threads", "4", "--output-phylo", "1", "--output-regions", "1", "--output-raw-phylo", "0",
                    "--genome-length", glen, "--coding-exons", exons, "--output", out, os.path.join(G, "53birds"), maf], check=True, capture_output=True)
    _assert_smoothed(out, golden_dir, "smooth53")
    assert not os.path.exists(os.path.join(out, "PhyloCSFRaw+1.wig")) and os.path.exists(os.path.join(out, "PhyloCSFpower.wig"))
    r = subprocess.run([BIN, "build-tracks", "--output-phylo", "1", "--output", out, os.path.join(G, "53birds"), maf], capture_output=True, text=True)
    assert r.returncode != 0 and "--genome-length and --coding-exons" in r.stdout


@pytest.mark.gpu
def test_maf_to_bigwig_to_annotation(golden_dir, tmp_path):
    """The whole chain without an external tool: build-tracks --output-phylo 1 --output-bigwig 1 (CUDA likelihoods -> wig text -> HMM ->
    bigWig) -> annotate-with-tracks.  Checked against the same annotation computed from bigWigs made of the files the REFERENCE wrote
    for this input (smooth53 fixtures + its expected PhyloCSFpower.wig): identical text, and the .bw files hold exactly the wig values."""
    _need_bin()
    from tests.test_host_annotate import CHROMS, SETS, _body, _sizes_file, parse_bigwig, wig_intervals
    G = os.path.join(golden_dir, "build-tracks")
    A = os.path.join(golden_dir, "annotate-with-tracks")
    tmp = str(tmp_path)
    maf = _gunzip(os.path.join(G, "galGal6_chr22_25_28_each_30k_bases.maf.gz"), tmp)
    _, exons, glen = _smooth_inputs(golden_dir, tmp_path, "smooth53")
    out = os.path.join(tmp, "out_bw")
    r = subprocess.run([BIN, "build-tracks", "--threads", "4", "--output-phylo", "1", "--output-bigwig", "1", "--genome-length", glen, "--coding-exons", exons,
                        "--output", out, os.path.join(G, "53birds"), maf], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    for name in ["PhyloCSFpower"] + [f"PhyloCSF{k}" for k in FRAMES6] + [f"PhyloCSFRaw{k}" for k in FRAMES6]:
        bw = parse_bigwig(os.path.join(out, name + ".bw"))
        assert {v[0]: v[1] for v in bw["chroms"].values()} == CHROMS          # lengths taken from the MAF's srcSize fields
        assert bw["intervals"] == sorted(wig_intervals(os.path.join(out, name + ".wig"))), name
    # the same tracks made of the reference-written wig files
    refdir = os.path.join(tmp, "ref_bw")
    os.makedirs(refdir)
    sizes = _sizes_file(tmp)
    wigs = {"PhyloCSFpower": os.path.join(G, "PhyloCSFpower.wig.gz")}
    wigs.update({f"PhyloCSF{k}": os.path.join(golden_dir, "ref-generated", f"smooth53.PhyloCSF{k}.wig.gz") for k in FRAMES6})
    for name, gz in wigs.items():
        wig = os.path.join(refdir, name + ".wig")
        with gzip.open(gz, "rb") as fi, open(wig, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        subprocess.run([BIN, "wig-to-bigwig", wig, sizes, os.path.join(refdir, name + ".bw")], check=True, capture_output=True)
    gtfs = [_gunzip(os.path.join(A, f"galGal6_chr22_25_28_subset_{n}.gtf.gz"), tmp) for n in SETS]
    for d, tracks in (("ann_gpu", out), ("ann_ref", refdir)):
        subprocess.run([BIN, "annotate-with-tracks", "--output", os.path.join(tmp, d), os.path.join(tracks, "PhyloCSF+1.bw")] + gtfs, check=True, capture_output=True)
    n_scored = 0
    for n in SETS:
        a = _body(os.path.join(tmp, "ann_gpu", f"galGal6_chr22_25_28_subset_{n}.PhyloCSF++.gtf"))
        b = _body(os.path.join(tmp, "ann_ref", f"galGal6_chr22_25_28_subset_{n}.PhyloCSF++.gtf"))
        assert a == b, n
        n_scored += sum("phylocsf_score_weighted_mean=" in ln and "mean=nan" not in ln for ln in a)
    assert n_scored > 100


@pytest.mark.gpu
def test_build_tracks_cli_golden(golden_dir, tmp_path):
    """Config 1 through the command line tool: the reference's seven expected wig files, byte for byte (FP64 path), with
    one and with several threads; the tcgen05 path within 1e-3 decibans of them."""
    _need_bin()
    G = os.path.join(golden_dir, "build-tracks")
    maf = _gunzip(os.path.join(G, "galGal6_chr22_25_28_each_30k_bases.maf.gz"), str(tmp_path))
    names = ["PhyloCSFpower.wig"] + [f"PhyloCSFRaw{s}{f}.wig" for s in "+-" for f in (1, 2, 3)]
    for threads in (1, 4):
        out = os.path.join(str(tmp_path), f"out{threads}")
        subprocess.run([BIN, "build-tracks", "--threads", str(threads), "--output", out, os.path.join(G, "53birds"), maf], check=True,
                       capture_output=True)
        for n in names:
            assert open(os.path.join(out, n), "rb").read() == gzip.open(os.path.join(G, n + ".gz"), "rb").read(), n
    out = os.path.join(str(tmp_path), "out_tc5")
    subprocess.run([BIN, "build-tracks", "--threads", "2", "--precision", "tc5", "--output", out, os.path.join(G, "53birds"), maf], check=True,
                   capture_output=True)
    for n in names:
        a = open(os.path.join(out, n)).read().split("\n")
        b = gzip.open(os.path.join(G, n + ".gz"), "rt").read().split("\n")
        assert len(a) == len(b)
        for x, y in zip(a, b):
            if x != y:
                assert not x.startswith("fixedStep") and abs(float(x) - float(y)) <= 0.0011, (n, x, y)


@pytest.mark.gpu
def test_score_msa_cli_golden(golden_dir, tmp_path):
    _need_bin()
    S = os.path.join(golden_dir, "score-msa")
    maf = os.path.join(str(tmp_path), "chr22.50alignments.maf")
    shutil.copy(os.path.join(S, "chr22.50alignments.maf"), maf)
    out = os.path.join(str(tmp_path), "o")
    subprocess.run([BIN, "score-msa", "--strategy", "fixed", "--comp-anc", "1", "--output", out, "100vertebrates", maf], check=True, capture_output=True)
    ours = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(out, "chr22.50alignments.maf.scores"))][2:]
    gold = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(S, "chr22.50alignments.fixed.scores"))][1:]
    gold = [g for g in gold if not g[0].startswith("seq")]
    assert len(ours) == len(gold)
    for o, g in zip(ours, gold):
        assert o[:4] == g[:4] and o[6] == g[6]
        assert abs(float(o[4]) - float(g[4])) <= 1e-3 and abs(float(o[5]) - float(g[5])) <= 1e-3


@pytest.mark.gpu
def test_build_tracks_cli_multiple_files(golden_dir, tmp_path):
    """Several alignment files in one call (one work queue across the files, the writer emits file after file): every output is the
    concatenation of the single-file outputs, as with the reference's append mode for file_id > 1 (build_tracks.hpp:245-259)."""
    _need_bin()
    R = os.path.join(golden_dir, "ref-generated")
    m1 = _gunzip(os.path.join(R, "tracks12.maf.gz"), str(tmp_path))
    m2 = os.path.join(str(tmp_path), "second.maf")
    with open(m1) as fi, open(m2, "w") as fo:
        fo.write(fi.read().replace(".chr1 ", ".chr2 "))
    names = ["PhyloCSFpower.wig"] + [f"PhyloCSFRaw{s}{f}.wig" for s in "+-" for f in (1, 2, 3)]
    outs = {}
    for tag, files in (("a", [m1]), ("b", [m2]), ("ab", [m1, m2])):
        out = os.path.join(str(tmp_path), "o_" + tag)
        subprocess.run([BIN, "build-tracks", "--threads", "5", "--output", out, "12flies"] + files, check=True, capture_output=True)
        outs[tag] = out
    for n in names:
        a = open(os.path.join(outs["a"], n), "rb").read()
        b = open(os.path.join(outs["b"], n), "rb").read()
        ab = open(os.path.join(outs["ab"], n), "rb").read()
        assert ab == a + b, n
        assert a == gzip.open(os.path.join(R, "tracks12." + n + ".gz"), "rb").read(), n


@pytest.mark.gpu
def test_build_tracks_cli_sharded_over_gpus(tmp_path):
    """The sharded product path (north star: MAF blocks dealt to the GPUs, host-side ordered gather, no collective; reference analogue
    build_tracks.hpp:88 job dealing and :27-53,245-259 ordered merge): --gpus 2 writes byte for byte what --gpus 1 writes, on both
    precisions, and both devices did score columns.  Skipped on a single-device box."""
    _need_bin()
    import json
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_synth_maf import write_synth_maf
    from phylocsfpp_b200.models import load_model
    maf = os.path.join(str(tmp_path), "shard.maf")
    write_synth_maf(maf, load_model("58mammals"), 400000, seed=77, mean_block=150, hole_p=1 / 60.0, ref_gap=0.01, alien_p=0.02, start0=800000)
    # small groups so that there is something to deal; PCSF_HOST_USE_ALL_GPUS: the host would not bring a second device up for 400 k columns
    env = dict(os.environ, PCSF_HOST_GROUP_COLS="30000", PCSF_HOST_STATS="1", PCSF_HOST_USE_ALL_GPUS="1")
    for prec in ("f64", "tc5"):
        outs = []
        for g in (1, 2):
            out = os.path.join(str(tmp_path), f"{prec}_g{g}")
            r = subprocess.run([BIN, "build-tracks", "--threads", "6", "--gpus", str(g), "--precision", prec, "--output", out, "58mammals", maf],
                               check=True, capture_output=True, text=True, env=env)
            st = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
            assert st["gpus"] == g and len(st["columns_per_gpu"]) == g and all(c > 0 for c in st["columns_per_gpu"]), st
            outs.append(out)
        for n in ["PhyloCSFpower.wig"] + [f"PhyloCSFRaw{s}{f}.wig" for s in "+-" for f in (1, 2, 3)]:
            assert open(os.path.join(outs[0], n), "rb").read() == open(os.path.join(outs[1], n), "rb").read(), (prec, n)
    # without --gpus the host uses every visible device
    r = subprocess.run([BIN, "build-tracks", "--threads", "6", "--output", os.path.join(str(tmp_path), "dflt"), "58mammals", maf],
                       check=True, capture_output=True, text=True, env=env)
    st = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert st["gpus"] == torch.cuda.device_count()


def _py_hmm_smooth(params, runs):
    """Independent restatement of the smoothing stage in plain Python (reference src/create_tracks.hpp:86-158, 162-200, 226-235):
    params = (coding_prior, coding_codons, [w0, w1, w2], [n0, n1, n2]); runs = [(chrom, start, [scores])] -> wig text."""
    import math
    from phylocsfpp_b200.tracks import my_format
    prior, ccod, w, n = params
    unnorm = [w[i] * n[i] for i in range(3)]
    c2nc = [w[i] / ccod for i in range(3)]
    nc2c = [1.0 / n[i] for i in range(3)]
    init = [prior] + [(1 - prior) * unnorm[i] / (unnorm[0] + unnorm[1] + unnorm[2]) for i in range(3)]
    T = [[0.0] * 4 for _ in range(4)]
    T[0][0] = 1.0 - (c2nc[0] + c2nc[1] + c2nc[2])
    for j in range(3):
        T[0][j + 1] = c2nc[j]
    for i in range(1, 4):
        T[i][0] = nc2c[i - 1]
        T[i][i] = 1.0 - nc2c[i - 1]
    emit = lambda s, x: math.pow(10, x / 10) if s == 0 else 1
    out = []
    for chrom, start, obs in runs:
        N = len(obs)
        f = [[init[s] * emit(s, obs[0]) for s in range(4)]]
        for p in range(1, N):
            row = []
            for s in range(4):
                acc = 0.0
                for q in range(4):
                    acc += f[p - 1][q] * T[q][s]
                row.append(acc * emit(s, obs[p]))
            mx = max(0.0, *row)
            f.append([v / mx for v in row])
        b = [[1.0] * 4 for _ in range(N)]
        for p in range(N - 2, -1, -1):
            row = []
            for s in range(4):
                acc = 0.0
                for q in range(4):
                    acc += (T[s][q] * emit(q, obs[p + 1]) * b[p + 1][q])
                row.append(acc)
            mx = max(0.0, *row)
            b[p] = [v / mx for v in row]
        out.append(f"fixedStep chrom={chrom} start={start} step=3 span=3")
        for p in range(N):
            tot = 0.0
            for s in range(4):
                tot += f[p][s] * b[p][s]
            pr = (f[p][0] * b[p][0]) / tot
            if pr < math.pow(10, -15.0):
                lo = -15.0
            elif pr > 1 - math.pow(10, -15.0):
                lo = 15.0
            else:
                lo = math.log10(pr / (1 - pr))
            out.append(my_format("%.3f", np.float32(lo)))
    return out


def test_hmm_smoothing_against_plain_python(golden_dir, tmp_path):
    """hmm.hpp's forward-backward against an independent plain-Python restatement on hand-made raw tracks: two runs that continue
    each other (joined, wig_file_reader.hpp:111-127), a gap, another chromosome, a single-codon run, extreme scores (log-odds clipped
    at +-15).  HMM parameters as estimated by the tool itself from the smooth53 exon list (--print-hmm)."""
    _need_bin()
    _, exons, glen = _smooth_inputs(golden_dir, tmp_path, "smooth53")
    raw = os.path.join(str(tmp_path), "handmade")
    os.makedirs(raw)
    rng = np.random.default_rng(3)
    sc = lambda n, mu: [float(np.float32(x)) for x in np.round(rng.normal(mu, 8.0, n), 3)]
    pieces = [("chr1", 100, sc(40, -5)), ("chr1", 100 + 3 * 40, sc(25, 12)),          # continues -> one run of 65
              ("chr1", 1000, sc(1, 3)), ("chr2", 7, [300.0, -300.0, 250.0] + sc(30, 0)), ("chr2", 400, sc(12, 40))]
    from phylocsfpp_b200.tracks import my_format
    for k in FRAMES6:
        with open(os.path.join(raw, f"PhyloCSFRaw{k}.wig"), "w") as fh:
            for chrom, start, vals in pieces:
                fh.write(f"fixedStep chrom={chrom} start={start} step=3 span=3\n")
                for v in vals:
                    fh.write(my_format("%.3f", np.float32(v)) + "\n")
    out = os.path.join(str(tmp_path), "handmade_out")
    r = subprocess.run([BIN, "smooth-tracks", "--genome-length", glen, "--coding-exons", exons, "--output-regions", "0", "--print-hmm", "1", raw, out],
                       check=True, capture_output=True, text=True)
    vals = {ln.split()[0] if not ln.startswith("nc") else "nc" + ln.split()[1]: ln.split() for ln in r.stdout.splitlines() if ln.strip()}
    params = (float(vals["coding_prior"][1]), float(vals["coding_codons"][1]), [float(vals[f"nc{i}"][3]) for i in range(3)],
              [float(vals[f"nc{i}"][5]) for i in range(3)])
    runs = [("chr1", 100, [float(my_format("%.3f", np.float32(v))) for v in pieces[0][2] + pieces[1][2]]),
            ("chr1", 1000, [float(my_format("%.3f", np.float32(v))) for v in pieces[2][2]]),
            ("chr2", 7, [float(my_format("%.3f", np.float32(v))) for v in pieces[3][2]]),
            ("chr2", 400, [float(my_format("%.3f", np.float32(v))) for v in pieces[4][2]])]
    expect = _py_hmm_smooth(params, runs)
    got = open(os.path.join(out, "PhyloCSF+1.wig")).read().split("\n")
    assert got[-1] == "" and got[:-1] == expect
    assert "15.0" in expect or "-15.0" in expect          # the clipping branch was exercised


def test_fixed_mean_against_reference_on_cpu(golden_dir, tmp_path):
    """score-msa --strategy fixed_mean restated on the CPU: per-codon decibans from the oracle (run_tracks), the plain-Python HMM above
    with the parameters the host estimates from the synthetic exon list (> 20 000 gaps: the std::shuffle path), mean log-odds summed in a
    float (score_msa.hpp:152-213) — against the reference's own output for msa29."""
    _need_bin()
    import math
    from oracle import oracle as orc
    from tests.util import write_synthetic_exons
    SPECIES29 = "Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat"
    R = os.path.join(golden_dir, "ref-generated")
    exons = os.path.join(str(tmp_path), "exons.txt")
    write_synthetic_exons(exons)
    raw = os.path.join(str(tmp_path), "empty_raw")
    os.makedirs(raw)
    for k in FRAMES6:
        open(os.path.join(raw, f"PhyloCSFRaw{k}.wig"), "w").close()
    r = subprocess.run([BIN, "smooth-tracks", "--genome-length", "400000000", "--coding-exons", exons, "--print-hmm", "1", raw,
                        os.path.join(str(tmp_path), "o")], check=True, capture_output=True, text=True)
    vals = {ln.split()[0] if not ln.startswith("nc") else "nc" + ln.split()[1]: ln.split() for ln in r.stdout.splitlines() if ln.strip()}
    params = (float(vals["coding_prior"][1]), float(vals["coding_codons"][1]), [float(vals[f"nc{i}"][3]) for i in range(3)],
              [float(vals[f"nc{i}"][5]) for i in range(3)])
    m = load_model("29mammals", SPECIES29)
    mc, mnc = orc.OracleModel(m.tree, m.S_c, m.f_c), orc.OracleModel(m.tree, m.S_nc, m.f_nc)
    alns = list(MafReader(os.path.join(R, "msa29.maf.gz"), m.seqid_to_phyloid, m.nl, False, warn=False))
    gold = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(R, "msa29.fixed_mean.scores")) if not ln.startswith("#") and not ln.startswith("seq\t")]
    assert len(alns) == len(gold)
    for a, g in list(zip(alns, gold))[::3]:
        scores = [float(x) for x in orc.run_tracks(mc, mnc, orc.translate(a.seqs))]
        if not scores:
            continue
        lines = _py_hmm_smooth_values(params, scores)
        acc = np.float32(0.0)
        for v in lines:
            acc = np.float32(float(acc) + v)
        mean = np.float32(acc) / np.float32(len(lines))
        assert abs(float(mean) - float(g[4])) <= 2e-5, (a.start_pos, float(mean), g[4])


def _py_hmm_smooth_values(params, obs):
    """Posterior log-odds (doubles, unformatted) of one run: the numeric core of _py_hmm_smooth."""
    import math
    prior, ccod, w, n = params
    unnorm = [w[i] * n[i] for i in range(3)]
    c2nc = [w[i] / ccod for i in range(3)]
    nc2c = [1.0 / n[i] for i in range(3)]
    init = [prior] + [(1 - prior) * unnorm[i] / (unnorm[0] + unnorm[1] + unnorm[2]) for i in range(3)]
    T = [[0.0] * 4 for _ in range(4)]
    T[0][0] = 1.0 - (c2nc[0] + c2nc[1] + c2nc[2])
    for j in range(3):
        T[0][j + 1] = c2nc[j]
    for i in range(1, 4):
        T[i][0] = nc2c[i - 1]
        T[i][i] = 1.0 - nc2c[i - 1]
    emit = lambda s, x: math.pow(10, x / 10) if s == 0 else 1
    N = len(obs)
    f = [[init[s] * emit(s, obs[0]) for s in range(4)]]
    for p in range(1, N):
        row = [sum_in_order([f[p - 1][q] * T[q][s] for q in range(4)]) * emit(s, obs[p]) for s in range(4)]
        mx = max(0.0, *row)
        f.append([v / mx for v in row])
    b = [[1.0] * 4 for _ in range(N)]
    for p in range(N - 2, -1, -1):
        row = [sum_in_order([(T[s][q] * emit(q, obs[p + 1]) * b[p + 1][q]) for q in range(4)]) for s in range(4)]
        mx = max(0.0, *row)
        b[p] = [v / mx for v in row]
    out = []
    for p in range(N):
        tot = sum_in_order([f[p][s] * b[p][s] for s in range(4)])
        pr = (f[p][0] * b[p][0]) / tot
        out.append(-15.0 if pr < math.pow(10, -15.0) else 15.0 if pr > 1 - math.pow(10, -15.0) else math.log10(pr / (1 - pr)))
    return out


def sum_in_order(xs):
    acc = 0.0
    for x in xs:
        acc += x
    return acc
