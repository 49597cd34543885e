"""Empirical codon model (ECM) files: 64x64 symmetric exchangeabilities + 64 codon frequencies.

Mirror of the reference's src/ecm.hpp:21-70 (`empirical_codon_model::open`): text line i (1-based,
i = 1..63) holds the i lower-triangle entries of matrix row i; line 65 holds the 64 codon frequencies;
everything else (blank line 64, the codon legend) is ignored.  Codon order is AAA, AAC, AAG, AAT, ACA ...
i.e. id = 16*n1 + 4*n2 + n3 with A,C,G,T = 0..3 (src/translation.hpp:80-88).
"""
from __future__ import annotations

import numpy as np


def load_ecm(path: str):
    """Returns (S float64[64,64] symmetric with zero diagonal, f float64[64])."""
    S = np.zeros((64, 64), np.float64)
    f = np.zeros(64, np.float64)
    with open(path, "r") as fh:
        for line_id, line in enumerate(fh, start=1):
            if line_id <= 63:
                vals = [float(tok) for tok in line.split()]
                if len(vals) != line_id:
                    raise ValueError(f"{path}: line {line_id} has {len(vals)} entries, expected {line_id}")
                for j, v in enumerate(vals):
                    S[j, line_id] = v
                    S[line_id, j] = v
            elif line_id == 65:
                vals = [float(tok) for tok in line.split()]
                if len(vals) != 64:
                    raise ValueError(f"{path}: line 65 has {len(vals)} codon frequencies, expected 64")
                f[:] = vals
    return S, f
