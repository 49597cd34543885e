"""Parity of the CUDA build-tracks path (pcsf_tracks, through the C-ABI) against the CPU oracle.

Tolerance (BASELINE.json north_star): |delta| <= 1e-3 decibans on scores; the FP64 path is asserted at
1e-6 here.  BLS/power, pattern indices and wig positions are bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from phylocsfpp_b200 import capi, tracks
from phylocsfpp_b200.maf import MafReader
from phylocsfpp_b200.models import load_model
from tests.util import pattern_index_reference, prune_longdouble, random_alignment, read_lines

pytestmark = pytest.mark.gpu

TOL = 1e-6  # decibans; the contract is 1e-3


def oracle_tracks(model, seqs):
    mc = orc.OracleModel(model.tree, model.S_c, model.f_c)
    mnc = orc.OracleModel(model.tree, model.S_nc, model.f_nc)
    plus, minus = orc.window_codons(seqs)
    return orc.run_tracks(mc, mnc, plus), orc.run_tracks(mc, mnc, minus), orc.bls(model.tree, seqs)[1], plus, minus


@pytest.mark.parametrize("name,L", [("7yeast", 64), ("12flies", 333), ("20flies", 200), ("29mammals", 400),
                                    ("58mammals", 300), ("100vertebrates", 200), ("49birds", 150)])
def test_tracks_vs_oracle(name, L):
    model = load_model(name)
    seqs = random_alignment(model.nl, L, seed=len(name) * 1000 + L)
    dm = capi.DeviceModel(model)
    res = dm.tracks(seqs, want_patterns=True)
    ref_p, ref_m, ref_b, plus, minus = oracle_tracks(model, seqs)
    assert np.abs(res["plus"] - ref_p).max() <= TOL
    assert np.abs(res["minus"] - ref_m).max() <= TOL
    assert np.array_equal(res["bls"], ref_b), "BLS must be bit-exact"
    assert np.array_equal(res["pattern_index"], pattern_index_reference(plus, minus, 1 << 22))
    assert res["stats"]["n_windows"] == 2 * (L - 2)
    assert res["stats"]["n_unique"] == int(res["pattern_index"].max()) + 1
    dm.close()


def test_model_matrices_match_oracle():
    model = load_model("58mammals")
    dm = capi.DeviceModel(model)
    for which, (S, f) in enumerate([(model.S_c, model.f_c), (model.S_nc, model.f_nc)]):
        lam, pi, P = dm.get(which)
        om = orc.OracleModel(model.tree, S, f)
        _, _, _, opi = om.eigen()
        assert np.abs(pi - opi).max() < 1e-14
        assert np.abs(P - om.pmatrices()).max() < 1e-12
    dm.close()


def test_reduced_tree_and_external_model(golden_dir):
    model = load_model("29mammals", "Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat")
    assert model.nl == 12
    seqs = random_alignment(model.nl, 500, seed=5)
    dm = capi.DeviceModel(model)
    res = dm.tracks(seqs)
    ref_p, ref_m, ref_b, _, _ = oracle_tracks(model, seqs)
    assert np.abs(res["plus"] - ref_p).max() <= TOL and np.abs(res["minus"] - ref_m).max() <= TOL
    assert np.array_equal(res["bls"], ref_b)
    dm.close()
    ext = load_model(os.path.join(golden_dir, "build-tracks", "53birds"))
    builtin = load_model("53birds")
    assert np.array_equal(ext.S_c, builtin.S_c) and np.array_equal(ext.tree.branch_len, builtin.tree.branch_len)


def test_dedup_and_chunking_do_not_change_results():
    model = load_model("29mammals")
    seqs = random_alignment(model.nl, 5000, seed=11, gap=0.5, conserve=0.97)
    dm = capi.DeviceModel(model)
    a = dm.tracks(seqs, want_patterns=True)
    b = dm.tracks(seqs, dedup=False)
    assert a["stats"]["n_unique"] < a["stats"]["n_windows"], "this input must contain repeated site patterns"
    assert b["stats"]["n_unique"] == b["stats"]["n_windows"]
    assert np.array_equal(a["plus"], b["plus"]) and np.array_equal(a["minus"], b["minus"])
    dm.set_chunk_columns(700)
    c = dm.tracks(seqs, want_patterns=True)
    assert c["stats"]["n_chunks"] == -(-(5000 - 2) // 700)
    assert np.array_equal(a["plus"], c["plus"]) and np.array_equal(a["minus"], c["minus"])
    plus, minus = orc.window_codons(seqs)
    assert np.array_equal(c["pattern_index"], pattern_index_reference(plus, minus, 700))
    assert np.array_equal(a["pattern_index"], pattern_index_reference(plus, minus, 1 << 22))
    dm.close()


def test_segmented_input_with_a_tiny_last_chunk():
    """pcsf_tracks with two or more dedup chunks copies / packs the input segment by segment (a chunk's columns + 2 of halo,
    rounded up to 16).  When the last chunk has <= 13 windows the previous segment already reached L, the last one is empty and
    must not be packed again (it used to start a 16-byte store at an unaligned column L): L = 64 k + 2 + r, r = 1..13."""
    model = load_model("12flies")
    dm = capi.DeviceModel(model)
    for r in (1, 2, 7, 13, 14, 15):
        L = 64 * 5 + 2 + r
        seqs = random_alignment(model.nl, L, seed=900 + r)
        dm.set_chunk_columns(0)
        whole = dm.tracks(seqs)
        dm.set_chunk_columns(64)
        seg = dm.tracks(seqs)
        assert seg["stats"]["n_chunks"] == -(-(L - 2) // 64)
        assert np.array_equal(whole["plus"], seg["plus"]) and np.array_equal(whole["minus"], seg["minus"])
        assert np.array_equal(whole["bls"], seg["bls"])
        again = dm.tracks(seqs)           # a sticky CUDA error from a misaligned store would surface here
        assert np.array_equal(again["plus"], seg["plus"])
    dm.close()


def test_edge_cases():
    model = load_model("12flies")
    dm = capi.DeviceModel(model)
    for L in (0, 1, 2, 3, 4, 63, 64, 65):
        seqs = random_alignment(model.nl, L, seed=L)
        res = dm.tracks(seqs)
        assert len(res["plus"]) == max(L - 2, 0) and len(res["bls"]) == L
        if L >= 3:
            ref_p, ref_m, ref_b, _, _ = oracle_tracks(model, seqs)
            assert np.abs(res["plus"] - ref_p).max() <= TOL and np.abs(res["minus"] - ref_m).max() <= TOL
            assert np.array_equal(res["bls"], ref_b)
    # all-gap / all-N columns: every leaf marginalised, score = 10*(log zC - log zNC)/ln10 of the row-sum products
    seqs = np.full((model.nl, 30), ord("-"), np.uint8)
    seqs[:, 10:20] = ord("N")
    res = dm.tracks(seqs)
    ref_p, ref_m, ref_b, _, _ = oracle_tracks(model, seqs)
    assert np.abs(res["plus"] - ref_p).max() <= TOL and np.all(res["bls"] == 0.0)
    assert res["stats"]["n_unique"] == 1
    # only one species present -> BLS 0 (additional_scores.hpp:66-79)
    seqs[0, :] = ord("A")
    assert np.all(dm.tracks(seqs)["bls"] == 0.0)
    # a character outside ACGTacgt.-Nn: the reference exit(37)s
    seqs[3, 7] = ord("R")
    with pytest.raises(capi.PcsfError) as ei:
        dm.tracks(seqs)
    assert ei.value.status == capi.PCSF_ERR_BAD_CHAR
    dm.close()


def test_reverse_complement_symmetry_large():
    """Size-independent property at a BASELINE-scale model: the '-' track of S equals the '+' track of
    revcomp(S) read backwards (build_tracks.hpp:175-226), and BLS reverses."""
    model = load_model("58mammals")
    L = 200_000
    seqs = random_alignment(model.nl, L, seed=3, gap=0.3, conserve=0.9)
    dm = capi.DeviceModel(model)
    a = dm.tracks(seqs)
    b = dm.tracks(orc.reverse_complement(seqs))
    assert np.array_equal(a["minus"], b["plus"][::-1])
    assert np.array_equal(a["plus"], b["minus"][::-1])
    assert np.array_equal(a["bls"], b["bls"][::-1])
    # spot-check 64 windows against the oracle
    idx = np.random.default_rng(0).integers(0, L - 2, 64)
    mc = orc.OracleModel(model.tree, model.S_c, model.f_c)
    mnc = orc.OracleModel(model.tree, model.S_nc, model.f_nc)
    sub = np.stack([seqs[:, i:i + 3] for i in idx], axis=1).reshape(model.nl, -1)
    ref = orc.run_tracks(mc, mnc, orc.translate(sub))
    assert np.abs(a["plus"][idx] - ref).max() <= TOL
    dm.close()


def test_build_tracks_golden(golden_dir):
    """Config 1: the reference's own expected wig files (test/tests.sh:15-19), external 53birds model."""
    G = os.path.join(golden_dir, "build-tracks")
    model = load_model(os.path.join(G, "53birds"))
    dm = capi.DeviceModel(model)
    out = {k: [] for k in tracks.FRAMES}
    power = []
    for aln in MafReader(os.path.join(G, "galGal6_chr22_25_28_each_30k_bases.maf.gz"), model.seqid_to_phyloid, model.nl,
                         True, warn=False):
        res = dm.tracks(aln.seqs)
        power += tracks.power_wig(aln.chrom, aln.start_pos, res["bls"])
        r = tracks.raw_wigs(aln.chrom, aln.start_pos, aln.chrom_len, res["plus"], res["minus"], res["bls"])
        for k in out:
            out[k] += r[k]
    assert power == read_lines(os.path.join(G, "PhyloCSFpower.wig.gz")), "power track must be byte-identical"
    flips = 0
    for (s, f), lines in out.items():
        gold = read_lines(os.path.join(G, tracks.wig_filename(s, f) + ".gz"))
        assert len(lines) == len(gold)
        for x, y in zip(lines, gold):
            if x == y:
                continue
            assert not y.startswith("fixedStep") and not x.startswith("fixedStep"), "wig positions must be exact"
            assert abs(float(x) - float(y)) <= 1e-3 + 1e-9
            flips += 1
    print(f"build-tracks golden: {flips} printed digits differ (FP64 path expects 0)")
    assert flips == 0
    dm.close()


@pytest.mark.parametrize("name,L", [("12flies", 400), ("58mammals", 700), ("100vertebrates", 300), ("53birds", 500), ("7yeast", 130)])
def test_tcgen05_path_within_contract(name, L):
    """tcgen05/TMEM path (kind::tf32 MMA with hi/lo split operands, leaf gathers as one-hot GEMMs, per-window
    log-scaling): |delta| <= 1e-3 decibans is the contract; asserted at 2e-4 against the oracle and the FP64 path."""
    model = load_model(name)
    seqs = random_alignment(model.nl, L, seed=177 + L, gap=0.35, conserve=0.8)
    dm = capi.DeviceModel(model)
    f64 = dm.tracks(seqs, want_patterns=True)
    t5 = dm.tracks(seqs, want_patterns=True, tc5=True)
    ref_p, ref_m, ref_b, _, _ = oracle_tracks(model, seqs)
    d = max(np.abs(t5["plus"] - ref_p).max(), np.abs(t5["minus"] - ref_m).max())
    print(f"{name}: tcgen05 path max |delta| vs oracle = {d:.3e} decibans")
    assert d <= 2e-4
    assert max(np.abs(t5["plus"] - f64["plus"]).max(), np.abs(t5["minus"] - f64["minus"]).max()) <= 2e-4
    assert np.array_equal(t5["bls"], ref_b) and np.array_equal(t5["pattern_index"], f64["pattern_index"])
    dm.close()


def test_tcgen05_path_many_tiles_and_no_dedup():
    """More window pairs than SMs (persistent loop, ring wrap-around, both phases of every barrier) and the
    no-dedup route; compared with the FP64 path on the same input."""
    model = load_model("29mammals")
    seqs = random_alignment(model.nl, 60000, seed=4242, gap=0.3, conserve=0.7)
    dm = capi.DeviceModel(model)
    f64 = dm.tracks(seqs, bls=False)
    t5 = dm.tracks(seqs, bls=False, tc5=True)
    t5n = dm.tracks(seqs, bls=False, tc5=True, dedup=False)
    d = max(np.abs(t5["plus"] - f64["plus"]).max(), np.abs(t5["minus"] - f64["minus"]).max())
    assert d <= 2e-4, d
    assert np.array_equal(t5["plus"], t5n["plus"]) and np.array_equal(t5["minus"], t5n["minus"])
    dm.close()


def test_tcgen05_path_extreme_columns_do_not_underflow():
    """All-certain, maximally diverged columns: with 58 leaves the raw likelihood is ~1e-150 (far below FP32's range, fine
    for FP64) -> the log-scaled tcgen05 path must still agree with the oracle; with 100 leaves even the reference's
    unscaled FP64 product underflows (z = 0 -> log 0 = -inf, fixed_lik.hpp:431) while the scaled path stays finite."""
    rng = np.random.default_rng(5)
    model = load_model("58mammals")
    seqs = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(model.nl, 300))]
    dm = capi.DeviceModel(model)
    t5 = dm.tracks(seqs, tc5=True)
    ref_p, ref_m, _, _, _ = oracle_tracks(model, seqs)
    assert np.isfinite(ref_p).all() and np.isfinite(ref_m).all()
    assert max(np.abs(t5["plus"] - ref_p).max(), np.abs(t5["minus"] - ref_m).max()) <= 1e-3
    dm.close()
    rng = np.random.default_rng(6)
    model = load_model("100vertebrates")
    seqs = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(model.nl, 300))]
    dm = capi.DeviceModel(model)
    t5, f64 = dm.tracks(seqs, tc5=True), dm.tracks(seqs)
    ref_p, ref_m, _, plus, minus = oracle_tracks(model, seqs)
    ref = np.concatenate([ref_p, ref_m])
    assert (~np.isfinite(ref)).sum() > 0, "this input is meant to underflow the reference's FP64 product"
    got64 = np.concatenate([f64["plus"], f64["minus"]])
    assert np.array_equal(np.isfinite(got64), np.isfinite(ref))
    got5 = np.concatenate([t5["plus"], t5["minus"]])
    assert np.isfinite(got5).all()
    # the yardstick for this input is extended precision: where the reference's doubles are denormal (just above the underflow)
    # its own scores are off by whole decibans, so it cannot judge the scaled path there
    lz = []
    for which in (0, 1):
        _, pi, P = dm.get(which)
        lz.append(prune_longdouble(model.tree, P, pi, np.concatenate([plus, minus], axis=1)))
    exact = (10.0 * (lz[0] - lz[1]) / np.log(np.longdouble(10.0))).astype(np.float64)
    assert np.abs(got5 - exact).max() <= 1e-3
    ok = np.isfinite(ref) & (np.abs(ref - exact) < 1e-6)           # windows the reference itself still resolves
    assert ok.sum() > 0 and np.abs(got5[ok] - ref[ok]).max() <= 1e-3
    dm.close()


@pytest.mark.parametrize("name,L", [("58mammals", 1 << 21), ("100vertebrates", 1 << 20)])
def test_tcgen05_path_properties_at_scale(name, L):
    """BASELINE-scale properties of the tcgen05 path (2 Mi columns of the config-3 shape, 58mammals; 1 Mi of config 4's 100vertebrates): every window is
    computed independently and deterministically, so (a) two runs are bit-identical, (b) the '-' track of S equals the
    '+' track of revcomp(S) read backwards, bit for bit, although the windows land in other tiles, lanes and chains,
    (c) all 4.2 M windows agree with the FP64 path within the 1e-3 deciban contract (asserted at 3e-4)."""
    import torch
    from phylocsfpp_b200.synth import synth_alignment
    model = load_model(name)
    seqs = synth_alignment(model, L, seed=99, device="cpu")[:, :L].numpy()
    dm = capi.DeviceModel(model)
    a = dm.tracks(seqs, bls=False, tc5=True)
    a2 = dm.tracks(seqs, bls=False, tc5=True)
    assert np.array_equal(a["plus"], a2["plus"]) and np.array_equal(a["minus"], a2["minus"])
    b = dm.tracks(orc.reverse_complement(seqs), bls=False, tc5=True)
    assert np.array_equal(a["minus"], b["plus"][::-1]) and np.array_equal(a["plus"], b["minus"][::-1])
    # (c) EVERY one of the 4.2 M windows against the FP64 DMMA path (the parity anchor, itself within 1e-6 of the oracle): contract
    # 1e-3 decibans, asserted at 3e-4
    f64 = dm.tracks(seqs, bls=False)
    dp, dmn = np.abs(f64["plus"] - a["plus"]).max(), np.abs(f64["minus"] - a["minus"]).max()
    print(f"{name}, {2 * (L - 2)} windows: tcgen05 path max |delta| vs FP64 path = {max(dp, dmn):.3e} decibans")
    assert dp <= 3e-4 and dmn <= 3e-4
    assert np.isfinite(a["plus"]).all() and np.isfinite(a["minus"]).all()
    dm.close()


@pytest.mark.parametrize("species", ["dmel,dsim", "dmel,dsim,dyak", "dmel,dsim,dyak,dpse,dvir"])
def test_tcgen05_path_tiny_trees(species):
    """--species reductions down to two leaves (the whole tree is one cherry: no GEMM step at all), three (one step)
    and five; also alignments shorter than one codon window."""
    model = load_model("12flies", species)
    dm = capi.DeviceModel(model)
    for L in (0, 1, 2, 3, 4, 700):
        seqs = random_alignment(model.nl, L, seed=31 + L, gap=0.3, conserve=0.6)
        f64 = dm.tracks(seqs)
        t5 = dm.tracks(seqs, tc5=True)
        assert t5["plus"].shape == f64["plus"].shape
        if L > 2:
            assert max(np.abs(t5["plus"] - f64["plus"]).max(), np.abs(t5["minus"] - f64["minus"]).max()) <= 2e-4
        assert np.array_equal(t5["bls"], f64["bls"])
    dm.close()
