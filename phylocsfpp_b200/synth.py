"""Synthetic alignment generator for the throughput configs (BASELINE.json configs 3-5).

Produces what the MAF reader hands to the hot path — the [nl, L] ASCII matrix of one concatenated
alignment chain (reference gaps already removed) — directly on the device with torch ops, seeded:
  * block structure: a new MAF block starts at each column with probability 1/120;
  * root codons drawn from the coding codon frequencies in 30 % of the blocks, non-coding in 70 %;
  * sequences evolve down the model tree, Jukes-Cantor, substitution probability 3/4 (1 - exp(-4t/3)) per branch;
  * missingness (30 % of species x column cells by default): non-reference species absent from a block with
    p = 0.18 (reader fills 'N'), '-' runs of geometric mean length 9, 2 % 'N' cells; 40 % of (block, species)
    pairs are soft-masked (lower case).  The reference row (species 0) has no gaps.
"""
from __future__ import annotations

import math

import torch


def _preorder(tree):
    out, stack = [], [tree.n - 1]
    while stack:
        i = stack.pop()
        out.append(i)
        if tree.child1[i] >= 0:
            stack.append(int(tree.child2[i]))
            stack.append(int(tree.child1[i]))
    return out


@torch.no_grad()
def synth_alignment(model, L: int, seed: int, device="cuda", gap: float = 0.30, ld: int | None = None) -> torch.Tensor:
    """Returns a uint8 ASCII tensor [nl, ld] (ld = L rounded up to 16, padding = 'N'); columns [0, L) are data."""
    tree = model.tree
    nl = tree.nl
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    if ld is None:
        ld = (L + 15) // 16 * 16
    out = torch.full((nl, ld), ord("N"), dtype=torch.uint8, device=dev)
    if L == 0:
        return out

    def rand(*shape):
        return torch.rand(*shape, generator=g, device=dev)

    # blocks
    new_block = rand(L) < (1.0 / 120.0)
    new_block[0] = True
    block = torch.cumsum(new_block.to(torch.int32), 0) - 1
    nblocks = int(block[-1].item()) + 1
    coding_block = rand(nblocks) < 0.30
    # root codons
    K = (L + 2) // 3
    pc = torch.as_tensor(model.f_c / model.f_c.sum(), dtype=torch.float32, device=dev)
    pn = torch.as_tensor(model.f_nc / model.f_nc.sum(), dtype=torch.float32, device=dev)
    cod_c = torch.multinomial(pc, K, replacement=True, generator=g)
    cod_n = torch.multinomial(pn, K, replacement=True, generator=g)
    is_c = coding_block[block[torch.arange(K, device=dev).clamp_(max=(L - 1) // 3) * 3]]
    cod = torch.where(is_c, cod_c, cod_n)
    root = torch.stack([cod // 16, (cod // 4) % 4, cod % 4], dim=1).reshape(-1)[:L].to(torch.uint8)
    del cod, cod_c, cod_n, is_c

    letters_up = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    letters_lo = torch.tensor(list(b"acgt"), dtype=torch.uint8, device=dev)
    p_absent, p_n, mean_run = 0.18, 0.02, 9.0
    r = max(0.0, (gap - p_absent - p_n * (1 - p_absent)) / (1 - p_absent))   # '-' fraction among present cells
    p_start = (r / mean_run) / max(1e-9, (1 - r))
    ar = torch.arange(L, device=dev, dtype=torch.int32)

    seqs = {tree.n - 1: root}
    for i in _preorder(tree):
        if i == tree.n - 1:
            continue
        parent = int(tree.parent[i])
        t = float(tree.branch_len[i])
        p_sub = 0.75 * (1.0 - math.exp(-4.0 * t / 3.0))
        par = seqs[parent]
        mut = rand(L) < p_sub
        shift = (torch.rand(L, generator=g, device=dev) * 3).to(torch.uint8).clamp_(max=2) + 1
        cur = torch.where(mut, (par + shift) % 4, par)
        if tree.child1[i] >= 0:
            seqs[i] = cur
        else:
            soft = (rand(nblocks) < 0.40)[block]
            row = torch.where(soft, letters_lo[cur.long()], letters_up[cur.long()])
            if i != 0:
                absent = (rand(nblocks) < p_absent)[block]
                starts = rand(L) < p_start
                lens = torch.clamp((torch.log(rand(L).clamp_(min=1e-12)) / math.log(1 - 1 / mean_run)).to(torch.int32) + 1, max=10000)
                last = torch.cummax(torch.where(starts, ar, torch.full_like(ar, -1)), 0).values
                in_run = (last >= 0) & ((ar - last) < lens[last.clamp(min=0).long()])
                row = torch.where(in_run, torch.full_like(row, ord("-")), row)
                row = torch.where(rand(L) < p_n, torch.full_like(row, ord("N")), row)
                row = torch.where(absent, torch.full_like(row, ord("N")), row)
            out[i, :L] = row
        # free the parent once both children are done (child2 is visited last in this pre-order)
        if int(tree.child2[parent]) == i:
            del seqs[parent]
    return out
