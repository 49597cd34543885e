"""Host-side mirrors: Newick/ECM/model loading, frame arithmetic, wig formatting, MAF reader semantics."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from phylocsfpp_b200 import newick, tracks
from phylocsfpp_b200.maf import MafReader
from phylocsfpp_b200.models import builtin_models, load_model
from tests.util import random_alignment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_builtin_models_load():
    names = builtin_models()
    assert len(names) == 11 and {"53birds", "58mammals", "100vertebrates", "29mammals", "23flies"} <= set(names)
    expect = {"53birds": 53, "58mammals": 58, "100vertebrates": 100, "29mammals": 29, "7yeast": 7}
    for n, nl in expect.items():
        m = load_model(n)
        assert m.nl == nl and m.tree.n == 2 * nl - 1
        assert np.allclose(m.S_c, m.S_c.T) and np.all(np.diag(m.S_c) == 0) and abs(m.f_c.sum() - 1) < 1e-6
        assert m.tree.branch_len[-1] == 0 and m.tree.parent[-1] == -1
    m = load_model("100vertebrates")
    assert m.seqid_to_phyloid["human"] == m.seqid_to_phyloid["hg38"] == m.seqid_to_phyloid["hg19"] == 0


def test_flatten_order_and_float_branches():
    root = newick.parse("((A:0.1,b:0.2):0.05,(C:.3,(D:0.4,E:0.5):0.6):0.7);")
    t = newick.flatten(root)
    assert t.labels[:5] == ["a", "b", "c", "d", "e"] and t.nl == 5 and t.n == 9
    assert list(t.child1[5:]) == [0, 3, 2, 5] and list(t.child2[5:]) == [1, 4, 6, 7]
    assert t.branch_len.dtype == np.float32 and t.branch_len[0] == np.float32(0.1) and t.branch_len_f64[0] == 0.1
    assert list(t.parent) == [5, 5, 7, 6, 6, 8, 7, 8, -1]


def test_species_reduction_adds_branch_lengths_in_double():
    m = load_model("29mammals", "Human,Chimp,Mouse,Dog")
    assert m.tree.labels[:4] == ["human", "chimp", "mouse", "dog"] and m.tree.n == 7
    full = load_model("29mammals")
    # mouse's reduced branch = sum of the collapsed chain above it
    i = full.tree.labels.index("mouse")
    chain, s = [], 0.0
    while full.tree.parent[i] >= 0:
        chain.append(full.tree.branch_len_f64[i])
        i = int(full.tree.parent[i])
    assert m.tree.branch_len_f64[2] <= sum(chain) + 1e-12
    with pytest.raises(ValueError):
        load_model("29mammals", "Human,Unicorn")
    # assembly aliases are accepted (models.hpp:1808-1822)
    m2 = load_model("29mammals", "hg19,panTro4,mm9,canFam2")
    assert m2.tree.labels[:4] == ["human", "chimp", "mouse", "dog"]


def test_my_format():
    f = tracks.my_format
    assert f("%.3f", 24.834) == "24.834" and f("%.3f", 3.54) == "3.54" and f("%.3f", 2.0) == "2.0"
    assert f("%.3f", -0.0001) == "-0.0" and f("%.4f", 0.89410) == "0.8941" and f("%.3f", 100.0) == "100.0"
    assert f("%.3f", float("nan")) == "nan"


@pytest.mark.parametrize("start,chrom_len,L", [(200001, 4_000_000, 100), (16050002, 51304566, 101), (7, 1000, 5), (9, 50, 2), (3, 30, 0)])
def test_window_formulation_equals_update_seqs(start, chrom_len, L):
    """frame tracks are slices of the per-offset '+'/'-' window arrays."""
    seqs = random_alignment(4, L, seed=start % 1000 + L)
    plus, minus = orc.window_codons(seqs)
    rc = orc.reverse_complement(seqs)
    for strand, frame in tracks.FRAMES:
        src = seqs if strand == "+" else rc
        skip, new_start = orc.skip_bases(start, chrom_len, L, strand, frame)
        pep = orc.translate(src, skip)                      # what the reference scores (update_seqs)
        o0, K = tracks.frame_offsets(start, chrom_len, L, strand, frame)
        assert K == pep.shape[1]
        if strand == "+":
            assert o0 == skip and new_start == start + skip
            assert np.array_equal(pep, plus[:, o0:o0 + 3 * K:3])
        else:
            assert o0 == (L - skip) % 3
            assert np.array_equal(pep[:, ::-1], minus[:, o0:o0 + 3 * K:3])   # build_tracks.hpp:175 reverses the scores


MAF_HEAD = "##maf version=1\n"


def _block(start, rows):
    out = "a score=1.0\n"
    for sp, seq, extra in rows:
        size = len(seq.replace("-", ""))
        out += f"s {sp}.chr1 {start if sp == 'hg38' else 777} {size} + {extra} {seq}\n"
    return out + "\n"


def test_maf_reader_concatenation_gaps_and_unknown_species(tmp_path):
    m = load_model("29mammals", "Human,Chimp,Mouse")
    txt = MAF_HEAD
    txt += _block(99, [("hg38", "ACG-T", 5000), ("panTro4", "AC-GT", 4000), ("xenTro9", "TTTTT", 9)])   # 4 ref bases: 100..103
    txt += _block(103, [("hg38", "GGCC", 5000), ("mm10", "G-CC", 3000)])                               # contiguous -> appended
    txt += _block(200, [("hg38", "AAAT", 5000)])                                                        # hole -> new alignment
    p = tmp_path / "t.maf"
    p.write_text(txt)
    rd = MafReader(str(p), m.seqid_to_phyloid, m.nl, True, warn=False)
    a, b = list(rd)
    assert (a.chrom, a.start_pos, a.chrom_len, a.L) == ("chr1", 100, 5000, 8)
    assert bytes(a.seqs[0]) == b"ACGTGGCC" and bytes(a.seqs[1]) == b"AC-TNNNN" and bytes(a.seqs[2]) == b"NNNNG-CC"
    assert (b.start_pos, b.L) == (201, 4) and "xentro9" in rd.unresolved
    # score-msa mode: one block = one alignment
    rd = MafReader(str(p), m.seqid_to_phyloid, m.nl, False, warn=False)
    assert [x.L for x in rd] == [4, 4, 4]


def test_maf_reader_breakpoint_overlap(tmp_path):
    """Chains are cut after the block crossing a multiple of 1,000,000 plus exactly 2 more reference bases, and
    the next chain starts at the block after the crossing block (parallel_file_reader.hpp:456-473,616-629,671-679)."""
    m = load_model("29mammals", "Human,Chimp")
    txt = MAF_HEAD
    txt += _block(999_990, [("hg38", "ACGTACGTAC", 9_000_000), ("panTro4", "ACGTACGTAC", 1)])   # 999991..1000000 crosses
    txt += _block(1_000_000, [("hg38", "GGGTT", 9_000_000), ("panTro4", "GGGTA", 1)])
    txt += _block(1_000_005, [("hg38", "CCCC", 9_000_000)])
    p = tmp_path / "bp.maf"
    p.write_text(txt)
    rd = MafReader(str(p), m.seqid_to_phyloid, m.nl, True, warn=False)
    a, b = list(rd)
    assert (a.start_pos, a.L) == (999_991, 12) and bytes(a.seqs[0]) == b"ACGTACGTACGG"
    assert (b.start_pos, b.L) == (1_000_001, 9) and bytes(b.seqs[0]) == b"GGGTTCCCC" and bytes(b.seqs[1]) == b"GGGTANNNN"


def test_bls_restatement_small():
    m = load_model("29mammals", "Human,Chimp,Mouse,Dog")
    t = m.tree
    seqs = np.frombuffer(b"AAN-" b"CANN" b"GNNC" b"TNNN", np.uint8).reshape(4, 4)
    score, per, bad = orc.bls(t, seqs)
    total = t.branch_len_f64[:-1].sum()
    assert not bad and per[2] == 0.0 and per[3] == 0.0           # fewer than two species present
    assert abs(per[0] - 1.0) < 1e-15
    # column 1: human+chimp only -> their two leaf branches
    assert abs(per[1] - (t.branch_len_f64[0] + t.branch_len_f64[1]) / total) < 1e-15
    assert abs(score - per.sum() / 4) < 1e-15


def test_synthetic_workload_has_the_baseline_shape():
    """The bench / fixture generator: deterministic per seed, reference row complete, ~30 % of the other cells
    missing ('-' or 'N'), only characters the reference accepts, both coding-like and non-coding-like stretches."""
    import torch
    from phylocsfpp_b200.models import load_model
    from phylocsfpp_b200.synth import synth_alignment
    m = load_model("58mammals")
    L = 60000
    a = synth_alignment(m, L, seed=5, device="cpu")[:, :L].numpy()
    b = synth_alignment(m, L, seed=5, device="cpu")[:, :L].numpy()
    assert a.shape == (m.nl, L) and (a == b).all()
    assert set(np.unique(a).tolist()) <= set(b"ACGTacgtN-")
    ref_missing = np.isin(a[0], np.frombuffer(b"N-", np.uint8)).mean()
    other_missing = np.isin(a[1:], np.frombuffer(b"N-", np.uint8)).mean()
    assert ref_missing == 0.0
    assert 0.25 <= other_missing <= 0.36, other_missing
    c = synth_alignment(m, L, seed=6, device="cpu")[:, :L].numpy()
    assert (a != c).mean() > 0.3


@pytest.fixture(scope="module")
def tc5_check_exe(tmp_path_factory):
    import subprocess
    exe = os.path.join(str(tmp_path_factory.mktemp("tc5chk")), "tc5_program_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "tc5_program_check.cpp")], check=True)
    return exe


@pytest.mark.parametrize("model,species", [("58mammals", ""), ("100vertebrates", ""), ("53birds", ""), ("7yeast", ""),
                                           ("29mammals", "Human,Chimp,Mouse"), ("29mammals", "Human,Mouse"),
                                           ("29mammals", "Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat")])
def test_tcgen05_step_program_on_the_cpu(tc5_check_exe, model, species):
    """The step program of k_prune_tc5 (chain starts, GEMM steps, leaf sources in ring order, cherry tables in table order,
    pushes / pops) interpreted on the host in FP64 equals the plain post-order Felsenstein recursion (fixed_lik.hpp:125-164) on
    random codon columns; the FP32-class emulation of the same program stays far inside the 1e-3 deciban contract."""
    import re
    import subprocess
    exe = tc5_check_exe
    env = dict(os.environ, PHYLOCSF_B200_DATA=os.path.join(ROOT, "phylocsfpp_b200", "data"))
    r = subprocess.run([exe, model, species, "60"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    d64 = float(re.search(r"max \|d log z\| = (\S+)", r.stdout).group(1))
    d32 = float(re.search(r"max \|d\| = (\S+) decibans", r.stdout).group(1))
    assert d64 < 1e-11 and d32 < 2e-4, r.stdout
    # k_bls's program with tabulated subtrees: prepare_model itself compares it bit for bit with the node-by-node program on 2000
    # presence masks (a mismatch is an error above); here: it really is shorter
    mb = re.search(r"BLS program: (\d+) entries \((\d+) tables, (\d+) doubles\) for (\d+) inner nodes", r.stdout)
    assert mb and int(mb.group(2)) >= 1 and int(mb.group(1)) <= max(1, (int(mb.group(4)) + 1) // 2), r.stdout
