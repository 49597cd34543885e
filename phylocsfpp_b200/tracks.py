"""build-tracks post-processing: window scores -> six frame tracks + power track, wig text.

Host-side mirror of the emission half of the reference's `run_tracks(file…)`
(src/phylocsf++build_tracks.hpp:128-228) and `my_fprintf` (src/common.hpp:48-68).

The device returns, per alignment of L reference columns,
    plus[o], minus[o]   o in [0, L-3]   decibans of the '+' codon n[o..o+2] and of the '-' codon
                                        comp(n[o+2]) comp(n[o+1]) comp(n[o])
    bls[i]              i in [0, L-1]   per-base branch-length score
and this module maps them onto the reference's frame tracks.  Derivation:
  '+', frame f: skip = (f - start_pos) mod 3 (clipped to L); codon xx covers offset o = skip + 3 xx
                (parallel_file_reader.hpp:77-85), emitted at position start_pos + o.
  '-', frame f: on the reverse-complemented rows skip_r = (f - (chrom_len - (start_pos + L) + 2)) mod 3
                (:92-97); after the score vector is reversed back (build_tracks.hpp:175-182) entry xx covers
                forward offset o = ((L - skip_r) mod 3) + 3 xx and is emitted at start_pos + o.
Thresholding (:199-205) uses float32 arithmetic: skip iff float(b0+b1+b2) < threshold_f32 * 3.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

FRAMES: List[Tuple[str, int]] = [("+", 1), ("+", 2), ("+", 3), ("-", 1), ("-", 2), ("-", 3)]


def my_format(fmt: str, value) -> str:
    """common.hpp:48-68 my_fprintf: the value is narrowed to float, printed with fmt, trailing zeros are
    stripped but one decimal is kept ("24.834", "3.54", "2.0", "-0.0")."""
    buf = list(fmt % float(np.float32(value)))
    i = len(buf) - 1
    while i >= 0:
        ch = buf[i]
        if ch == ".":
            buf = buf[:i + 1] + ["0"]
            break
        if ch.isdigit():
            if ch != "0":
                break
            buf = buf[:i]
        i -= 1
    return "".join(buf)


def _mod3(x: int) -> int:
    # C++: int64 % 3 then += 3 if negative == mathematical mod for these magnitudes
    return x % 3


def frame_offsets(start_pos: int, chrom_len: int, L: int, strand: str, frame: int) -> Tuple[int, int]:
    """(o0, K): the frame's codons start at forward offsets o0, o0+3, ... (K of them)."""
    if strand == "+":
        skip = min(_mod3(frame - start_pos), L)
        return skip, (L - skip) // 3
    skip_r = min(_mod3(frame - (chrom_len - (start_pos + L) + 2)), L)
    return (L - skip_r) % 3, (L - skip_r) // 3


def power_wig(chrom: str, start_pos: int, bls: np.ndarray) -> List[str]:
    """build_tracks.hpp:139-158."""
    L = len(bls)
    out: List[str] = []
    skip = _mod3(3 - start_pos)
    if skip + 2 < L:
        out.append(f"fixedStep chrom={chrom} start={start_pos + skip} step=3 span=3")
    for pos in range(skip, L - 2, 3):
        avg = np.float32((bls[pos] + bls[pos + 1] + bls[pos + 2]) / 3.0)
        out.append(my_format("%.4f", avg))
    return out


def raw_wigs(chrom: str, start_pos: int, chrom_len: int, plus: np.ndarray, minus: np.ndarray, bls: np.ndarray,
             threshold: float = 0.1) -> Dict[Tuple[str, int], List[str]]:
    """build_tracks.hpp:160-216 for one alignment: wig lines per (strand, frame)."""
    L = len(bls)
    thr3 = np.float32(np.float32(threshold) * np.float32(3))
    out: Dict[Tuple[str, int], List[str]] = {}
    for strand, frame in FRAMES:
        lines: List[str] = []
        o0, K = frame_offsets(start_pos, chrom_len, L, strand, frame)
        src = plus if strand == "+" else minus
        prev = -4
        for xx in range(K):
            o = o0 + 3 * xx
            bsum = np.float32(bls[o] + bls[o + 1] + bls[o + 2])
            if bsum < thr3:
                continue
            new = start_pos + o
            if prev + 3 != new:
                lines.append(f"fixedStep chrom={chrom} start={new} step=3 span=3")
            prev = new
            lines.append(my_format("%.3f", src[o]))
        out[(strand, frame)] = lines
    return out


def wig_filename(strand: str, frame: int) -> str:
    return f"PhyloCSFRaw{strand}{frame}.wig"
