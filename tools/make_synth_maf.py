#!/usr/bin/env python3
"""Writes a synthetic MAF file of the BASELINE shape for the end-to-end runs and the
reader tests: the [nl, N] matrix of phylocsfpp_b200.synth.synth_alignment cut into blocks of geometric length
(mean 120), contiguous except for a hole of 1..300 bases with probability 1/400; species whose row is all 'N' in a
block are omitted (the reader pads them); with probability `ref_gap` a column with '-' in the reference row is
inserted (the reader deletes it); now and then a row of a species unknown to the model is added.
usage: tools/make_synth_maf.py <model> <columns> <out.maf> [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_synth_maf(path, model, ncols, seed=1, start0=10000, mean_block=120, hole_p=1 / 400.0, ref_gap=0.01, alien_p=0.02,
                    loguniform_blocks=None, mat=None, chain_cols=None):
    """loguniform_blocks=(lo, hi): block lengths log-uniform in [lo, hi] and a hole after every block (BASELINE config 5:
    every block is its own alignment).  mat: write these [nl, >=ncols] ASCII columns instead of generating them.
    chain_cols: additionally force a block boundary and a hole every chain_cols columns (bounded chains = units of work for
    the reference's job-parallel reader)."""
    import torch
    from phylocsfpp_b200.models import sequence_name_mapping
    from phylocsfpp_b200.synth import synth_alignment
    rng = np.random.default_rng(seed)
    if mat is None:
        dev = "cuda" if torch.cuda.is_available() and os.environ.get("PCSF_SYNTH_CPU") is None else "cpu"
        mat = synth_alignment(model, ncols, seed=seed, device=dev)[:, :ncols].cpu().numpy()
    else:
        mat = np.ascontiguousarray(mat[:, :ncols])
    nl = model.nl
    names = []
    for i in range(nl):
        label = model.tree.labels[i]
        alts = sequence_name_mapping.get(label, [])
        names.append(alts[0] if alts else label)
    src_size = start0 + ncols + ncols // 100 + 400000
    if loguniform_blocks:
        lo, hi = loguniform_blocks
        lens = np.exp(rng.uniform(np.log(lo), np.log(hi), size=int(ncols / lo) + 1)).astype(np.int64)
        cuts = np.concatenate([[0], np.cumsum(lens)])
        cuts = np.unique(np.concatenate([cuts[cuts < ncols], [ncols]]))
        hole_p = 1.1
    else:
        cuts = np.flatnonzero(rng.random(ncols) < 1.0 / mean_block)
        cuts = np.unique(np.concatenate([[0], cuts, [ncols]]))
    forced = set()
    if chain_cols:
        forced = set(range(chain_cols, ncols, chain_cols))
        cuts = np.unique(np.concatenate([cuts, np.fromiter(forced, np.int64, len(forced))]))
    pos = start0
    n_blocks = 0
    with open(path, "wb") as fh:
        fh.write(b"##maf version=1 scoring=synthetic\n")
        for a, b in zip(cuts[:-1], cuts[1:]):
            blk = mat[:, a:b]
            size = int(b - a)
            gaps = np.flatnonzero(rng.random(size) < ref_gap) if ref_gap > 0 else np.zeros(0, np.int64)
            if gaps.size:
                cols = rng.choice(np.frombuffer(b"ACGTacgt-", np.uint8), size=(nl, gaps.size))
                cols[0, :] = ord("-")
                blk = np.insert(blk, gaps, cols, axis=1)
            fh.write(b"a score=0.0\n")
            lines = []
            for s in range(nl):
                row = blk[s]
                if s != 0 and (row == ord("N")).all():
                    continue
                if s == 0:
                    lines.append(b"s %s.chr1 %d %d + %d %s\n" % (names[0].encode(), pos, size, src_size, row.tobytes()))
                else:
                    nb = int(((row != ord("-"))).sum())
                    lines.append(b"s %s.scaffold_%d %d %d %s %d %s\n" % (names[s].encode(), s, 1000 + a, nb, b"+-"[s % 2:s % 2 + 1], 50000000, row.tobytes()))
                if s == 0 and rng.random() < alien_p:
                    # an unknown species right after the reference row
                    lines.append(b"s alien9.chrZ %d %d + 1000000 %s\n" % (int(a), size, blk[0].tobytes().replace(b"-", b"A")))
            fh.write(b"".join(lines))
            fh.write(b"\n")
            pos += size
            n_blocks += 1
            if rng.random() < hole_p or int(b) in forced:
                pos += int(rng.integers(1, 301))
    return dict(columns=int(ncols), blocks=n_blocks, bytes=os.path.getsize(path))


if __name__ == "__main__":
    from phylocsfpp_b200.models import load_model
    m = load_model(sys.argv[1])
    kw = dict(loguniform_blocks=(30, 600)) if os.environ.get("PCSF_SYNTH_SINGLE_BLOCKS") else {}
    print(write_synth_maf(sys.argv[3], m, int(sys.argv[2]), seed=int(sys.argv[4]) if len(sys.argv) > 4 else 1, **kw))
