"""Shared helpers for the parity tests."""
from __future__ import annotations

import gzip
import os

import numpy as np

ALPHABET = np.frombuffer(b"ACGTacgtN-.n", np.uint8)


def random_alignment(nl: int, L: int, seed: int, gap: float = 0.3, conserve: float = 0.7) -> np.ndarray:
    """Random ASCII alignment [nl, L]: a random ancestral row copied to every species with per-cell mutation
    (1 - conserve), upper/lower case mix, and `gap` of the cells replaced by one of N - . n."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, size=L)
    cells = np.where(rng.random((nl, L)) < conserve, anc[None, :], rng.integers(0, 4, size=(nl, L)))
    lower = rng.random((nl, L)) < 0.3
    out = ALPHABET[cells + 4 * lower]
    miss = rng.random((nl, L)) < gap
    out = np.where(miss, ALPHABET[8 + rng.integers(0, 4, size=(nl, L))], out)
    return np.ascontiguousarray(out, np.uint8)


def read_lines(path: str):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        lines = fh.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def first_occurrence_ranks(keys) -> np.ndarray:
    """CPU restatement of the site-pattern index rule: rank of the first occurrence of each key, in order."""
    seen = {}
    out = np.zeros(len(keys), np.uint32)
    for i, k in enumerate(keys):
        r = seen.get(k)
        if r is None:
            r = len(seen)
            seen[k] = r
        out[i] = r
    return out


def pattern_index_reference(plus: np.ndarray, minus: np.ndarray, chunk_cols: int) -> np.ndarray:
    """pattern_index[2*o + strand] for codon-id matrices plus/minus [nl, W], dedup domain = chunk of columns."""
    nl, W = plus.shape
    out = np.zeros(2 * W, np.uint32)
    for c0 in range(0, W, chunk_cols):
        c1 = min(W, c0 + chunk_cols)
        keys = []
        for o in range(c0, c1):
            keys.append(plus[:, o].tobytes())
            keys.append(minus[:, o].tobytes())
        out[2 * c0:2 * c1] = first_occurrence_ranks(keys)
    return out


def write_synthetic_exons(path: str, n: int = 46000, seed: int = 99) -> None:
    """A deterministic BED-like coding-exon list (chrom, strand, phase, start, end) with overlaps, for the HMM
    parameter estimation (reference src/estimate_hmm_parameter.hpp:243-340): more than 2 x 20 000 inter-exon gaps in a
    few chrom:strand:phase classes so that the gap subsampling (std::shuffle, default_random_engine(0)) is exercised."""
    x = seed
    pos = {}
    with open(path, "w") as fh:
        for _ in range(n):
            x = (x * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            chrom = "chr%d" % (1 + (x >> 60) % 2)
            strand = "+-"[(x >> 58) & 1]
            phase = (x >> 55) % 3
            key = (chrom, strand, phase)
            gap_class = (x >> 40) % 10
            gap = 30 + (x >> 20) % (200 if gap_class < 3 else 6000 if gap_class < 9 else 150000)
            length = 20 + (x >> 8) % 400
            start = pos.get(key, 1000) + gap - (120 if (x >> 5) % 17 == 0 else 0)      # now and then an overlap
            start = max(1, start)
            end = start + length
            pos[key] = end
            fh.write("%s\t%s\t%d\t%d\t%d\n" % (chrom, strand, phase, start, end))


def prune_longdouble(tree, P: np.ndarray, pi: np.ndarray, codons: np.ndarray) -> np.ndarray:
    """log z per window by Felsenstein pruning (fixed_lik.hpp:125-164) in x87 extended precision: 15 exponent bits, so the
    unscaled products of 100 certain leaves (~1e-400) neither underflow nor go denormal as the reference's doubles do.
    tree: flattened tree (child1, child2, nl, n); P [(n-1), 64, 64] row-major P_b[a][c]; codons uint8 [nl, W] (64 = gap/N)."""
    LD = np.longdouble
    nl, n = tree.nl, tree.n
    W = codons.shape[1]
    Pl = P.astype(LD)
    alpha = [None] * n
    msg = [None] * n
    for v in range(n):
        if v < nl:
            x = codons[v].astype(np.int64)
            col = np.ones((W, 64), LD)
            certain = x < 64
            col[certain] = Pl[v][:, x[certain]].T          # message of a leaf with codon x: P_v[:, x]; all ones for a gap
            msg[v] = col
        else:
            a = msg[tree.child1[v]] * msg[tree.child2[v]]
            alpha[v] = a
            if v < n - 1:
                msg[v] = a @ Pl[v].T                        # msg[w][a] = sum_b P_v[a][b] alpha[w][b]
    z = alpha[n - 1] @ pi.astype(LD)
    return np.log(z)
