"""ctypes/numpy wrapper around oracle/liboracle.so (phylocsf_oracle.c).

TEST INFRASTRUCTURE ONLY — the parity checker for the CUDA path.  May be imported from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the product
package (phylocsfpp_b200/).  Parity status: pinned against the reference's golden files
(tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "phylocsf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_model_new.restype = C.c_void_p
        L.orc_model_new.argtypes = [C.c_int, _i16p, _i16p, _f32p, _f64p, _f64p]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_get.argtypes = [C.c_void_p, _f64p, _f64p, _f64p, _f64p]
        L.orc_model_set_rho.restype = C.c_int
        L.orc_model_set_rho.argtypes = [C.c_void_p, C.c_double]
        L.orc_model_pmatrices.restype = C.POINTER(C.c_double)
        L.orc_model_pmatrices.argtypes = [C.c_void_p]
        L.orc_lpr_leaves.argtypes = [C.c_void_p, _u8p, C.c_int64, C.c_int64, C.c_int,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.orc_max_lik.restype = C.c_int
        L.orc_max_lik.argtypes = [C.c_void_p, _u8p, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double,
                                  C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_omega.restype = C.c_int
        L.orc_omega.argtypes = [C.c_int, _i16p, _i16p, C.c_void_p, _u8p, C.c_int64, C.c_int64, C.c_uint32,
                                C.POINTER(C.c_double), C.c_void_p]
        L.orc_mt_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_mt_next.restype = C.c_uint32
        L.orc_mt_next.argtypes = [C.c_void_p]
        L.orc_uniform.restype = C.c_double
        L.orc_uniform.argtypes = [C.c_void_p, C.c_double]
        L.orc_bls.restype = C.c_double
        L.orc_bls.argtypes = [C.c_int, _i16p, _i16p, _f64p, _u8p, C.c_int64, C.c_int64, C.c_void_p,
                              C.POINTER(C.c_int)]
        L.orc_window_codons.argtypes = [_u8p, C.c_int, C.c_int64, C.c_int64, _u8p, _u8p]
        L.orc_skip_bases.restype = C.c_int64
        L.orc_skip_bases.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_char, C.c_uint,
                                     C.POINTER(C.c_uint64)]
        L.orc_translate_row.argtypes = [_u8p, C.c_uint64, C.c_uint64, _u8p]
        L.orc_run_tracks.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int64, C.c_int64, _f64p]
        L.orc_branch_time.restype = C.c_double
        L.orc_branch_time.argtypes = [C.c_float, C.c_double]
        L.orc_build_q.argtypes = [_f64p, _f64p, _f64p]
        _LIB = L
    return _LIB


class MT19937:
    """std::mt19937 as restated in the oracle (state lives in a ctypes buffer)."""

    def __init__(self, seed: int = 42):
        self.buf = C.create_string_buffer(4 * 624 + 8)
        self.seed(seed)

    def seed(self, seed: int):
        lib().orc_mt_seed(self.buf, seed)

    def next_u32(self) -> int:
        return lib().orc_mt_next(self.buf)

    def uniform(self, width: float) -> float:
        return lib().orc_uniform(self.buf, width)


class OracleModel:
    """One (tree, ECM) instance: run.hpp:38-39 PhyloCSFModel_make, at tree scale rho."""

    def __init__(self, tree, S, f):
        self._keep = (np.ascontiguousarray(tree.child1, np.int16), np.ascontiguousarray(tree.child2, np.int16),
                      np.ascontiguousarray(tree.branch_len, np.float32))
        self.nl, self.n = tree.nl, tree.n
        self.h = lib().orc_model_new(tree.nl, self._keep[0], self._keep[1], self._keep[2],
                                     np.ascontiguousarray(S, np.float64), np.ascontiguousarray(f, np.float64))
        self.rho = None
        self.set_rho(1.0)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_free(self.h)
            self.h = None

    def set_rho(self, rho: float) -> int:
        self.rho = rho
        return lib().orc_model_set_rho(self.h, rho)

    def eigen(self):
        lam, SR, SRi, pi = np.zeros(64), np.zeros((64, 64)), np.zeros((64, 64)), np.zeros(64)
        lib().orc_model_get(self.h, lam, SR, SRi, pi)
        return lam, SR, SRi, pi

    def pmatrices(self) -> np.ndarray:
        p = lib().orc_model_pmatrices(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n - 1, 64, 64)).copy()

    def lpr_leaves(self, peptides: np.ndarray, compute_anc: bool = False):
        """fixed_lik.hpp:362 at the current rho.  peptides: uint8 [nl, K].
        Returns (lpr, elpr_anc, lpr_per_codon[K], anc_per_codon[K])."""
        pep = np.ascontiguousarray(peptides, np.uint8)
        nl, K = pep.shape
        assert nl == self.nl
        per = np.zeros(max(K, 1), np.float64)
        anc = np.zeros(max(K, 1), np.float64)
        lpr, el = C.c_double(), C.c_double()
        lib().orc_lpr_leaves(self.h, pep, K, K, int(compute_anc), C.byref(lpr), C.byref(el),
                             per.ctypes.data, anc.ctypes.data)
        return lpr.value, el.value, per[:K], anc[:K]

    def max_lik(self, peptides: np.ndarray, gen: MT19937, compute_anc: bool = False,
                init: float = 1.0, lo: float = 1e-2, hi: float = 10.0):
        """fixed_lik.hpp:511 max_lik_lpr_leaves.  Returns (status, lpr, elpr_anc, x_final, n_evals)."""
        pep = np.ascontiguousarray(peptides, np.uint8)
        nl, K = pep.shape
        lpr, el, x = C.c_double(), C.c_double(), C.c_double()
        ne = C.c_int()
        st = lib().orc_max_lik(self.h, pep, K, K, int(compute_anc), init, lo, hi, gen.buf, C.byref(lpr),
                               C.byref(el), C.byref(x), C.byref(ne))
        return st, lpr.value, el.value, x.value, ne.value


def window_codons(seqs: np.ndarray):
    """All '+' and '-' codon windows of a [nl, L] ASCII matrix -> two uint8 [nl, L-2] matrices."""
    seqs = np.ascontiguousarray(seqs, np.uint8)
    nl, L = seqs.shape
    W = max(L - 2, 0)
    plus = np.zeros((nl, W), np.uint8)
    minus = np.zeros((nl, W), np.uint8)
    if W > 0:
        lib().orc_window_codons(seqs, nl, L, L, plus, minus)
    return plus, minus


def translate(seqs: np.ndarray, skip: int = 0) -> np.ndarray:
    """alignment_t::translate / update_seqs codon ids from offset `skip`: uint8 [nl, (L-skip)//3]."""
    seqs = np.ascontiguousarray(seqs, np.uint8)
    nl, L = seqs.shape
    K = (L - skip) // 3
    out = np.zeros((nl, K), np.uint8)
    for s in range(nl):
        lib().orc_translate_row(seqs[s], L, skip, out[s])
    return out


def skip_bases(orig_start_pos: int, chrom_len: int, L: int, strand: str, frame: int):
    ns = C.c_uint64()
    sk = lib().orc_skip_bases(orig_start_pos, chrom_len, L, strand.encode(), frame, C.byref(ns))
    return int(sk), int(ns.value)


_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_a] = _b


def reverse_complement(seqs: np.ndarray) -> np.ndarray:
    """build_tracks.hpp:219-226."""
    return np.ascontiguousarray(_COMP[seqs[:, ::-1]])


def bls(tree, seqs: np.ndarray, per_base: bool = True):
    """additional_scores.hpp:43 compute_bls_score.  Returns (score, per_base[L] or None, bad_char)."""
    seqs = np.ascontiguousarray(seqs, np.uint8)
    nl, L = seqs.shape
    pb = np.zeros(max(L, 1), np.float64) if per_base else None
    bad = C.c_int(0)
    sc = lib().orc_bls(nl, np.ascontiguousarray(tree.child1, np.int16), np.ascontiguousarray(tree.child2, np.int16),
                       np.ascontiguousarray(tree.branch_len_f64, np.float64), seqs, L, L,
                       pb.ctypes.data if per_base else None, C.byref(bad))
    return sc, (pb[:L] if per_base else None), bool(bad.value)


def run_tracks(mc: OracleModel, mnc: OracleModel, peptides: np.ndarray) -> np.ndarray:
    """run.hpp:35 run_tracks: per-codon decibans."""
    pep = np.ascontiguousarray(peptides, np.uint8)
    nl, K = pep.shape
    out = np.zeros(max(K, 1), np.float64)
    lib().orc_run_tracks(mc.h, mnc.h, pep, K, K, out)
    return out[:K]


def run_fixed(mc: OracleModel, mnc: OracleModel, peptides: np.ndarray, compute_anc: bool):
    """run.hpp:196-209 FIXED: (float32 score, float32 anc)."""
    lc, ac, _, _ = mc.lpr_leaves(peptides, compute_anc)
    lnc, anc, _, _ = mnc.lpr_leaves(peptides, compute_anc)
    ln10 = np.log(10.0)
    return np.float32(10.0 * (lc - lnc) / ln10), np.float32(10.0 * (ac - anc) / ln10)


def run_mle(mc: OracleModel, mnc: OracleModel, peptides: np.ndarray, compute_anc: bool, seed: int = 42):
    """run.hpp:191-195 + 206-209 MLE: one mt19937(42) shared by the coding then the non-coding fit.
    Returns (float32 score, float32 anc, info dict); NaNs if the reference would have thrown."""
    gen = MT19937(seed)
    sc, lc, ac, xc, nc_ = mc.max_lik(peptides, gen, compute_anc)
    sn, lnc, anc, xn, nn = mnc.max_lik(peptides, gen, compute_anc)
    mc.set_rho(1.0)
    mnc.set_rho(1.0)
    info = dict(rho_c=xc, rho_nc=xn, evals_c=nc_, evals_nc=nn, status=(sc, sn))
    if sc or sn:
        return np.float32(np.nan), np.float32(np.nan), info
    ln10 = np.log(10.0)
    return np.float32(10.0 * (lc - lnc) / ln10), np.float32(10.0 * (ac - anc) / ln10), info


def run_omega(tree, peptides: np.ndarray, seed: int = 42):
    """run.hpp:59-182 OMEGA: (float32 score, info dict); NaN if the reference would have thrown."""
    pep = np.ascontiguousarray(peptides, np.uint8)
    nl, K = pep.shape
    c1 = np.ascontiguousarray(tree.child1, np.int16)
    c2 = np.ascontiguousarray(tree.child2, np.int16)
    bl = np.ascontiguousarray(tree.branch_len, np.float32)
    score = C.c_double()
    info = np.zeros(5)
    st = lib().orc_omega(nl, c1, c2, bl.ctypes.data, pep, K, K, seed, C.byref(score), info.ctypes.data)
    d = dict(rho=info[0], kappa=info[1], lpr_h0=info[2], lpr_h1=info[3], evals=int(info[4]), status=st)
    return (np.float32(np.nan) if st else np.float32(score.value)), d
