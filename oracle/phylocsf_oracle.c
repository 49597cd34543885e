/*
 * phylocsf_oracle.c — CPU restatement of PhyloCSF++'s per-codon-column likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path in
 * phylocsfpp_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against the
 * reference's own golden files (test/expected_results/build-tracks/PhyloCSFRaw{+,-}{1,2,3}.wig and
 * PhyloCSFpower.wig; test/maf-file-small/PhyloCSFpp-results/chr22.50alignments.{fixed,mle}.scores;
 * test/maf-file-medium/chr22.516alignments.maf.{fixed,mle}.scores), copies of which live in tests/golden/.
 *
 * Every function cites the reference file:line (relative to the reference repo root) it follows.
 * One deliberate deviation: the reference diagonalises Q with GSL's gsl_eigen_nonsymmv + a complex LU
 * inverse (src/instance.hpp:324-346).  GSL is a system dependency that is absent here (unpinned in
 * CMakeLists.txt:43).  Q is reversible (q_ij = s_ij f_j with s symmetric), so this file diagonalises
 * the similar symmetric matrix D^{1/2} Q D^{-1/2} with cyclic Jacobi instead; P(t) = exp(Qt) does not
 * depend on the eigensolver (differences ~1e-13, far below every golden file's printed precision).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NS 64 /* codon states */
#ifndef M_PI
#define M_PI 3.14159265358979323846 /* -std=c11 hides it */
#endif

/* ------------------------------------------------------------------------------------------------
 * translation.hpp:29-53 get_dna_id — ACGT/acgt -> 0..3, ".-Nn" -> 4, anything else -> 99 (the
 * reference prints and exit(37)s; the oracle reports 99 so the caller can assert the same condition).
 */
uint8_t orc_dna_id(char c) {
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case '.': case '-': case 'N': case 'n': return 4;
    default: return 99;
    }
}

/* translation.hpp:55-78 complement — case preserving, non-ACGT unchanged. */
char orc_complement(char c) {
    switch (c) {
    case 'A': return 'T'; case 'a': return 't';
    case 'C': return 'G'; case 'c': return 'g';
    case 'G': return 'C'; case 'g': return 'c';
    case 'T': return 'A'; case 't': return 'a';
    default: return c;
    }
}

/* translation.hpp:80-88 get_amino_acid_id — 16a+4b+c, or 64 ("marginalise") if any base is 4. */
uint8_t orc_codon_id(char n1, char n2, char n3) {
    uint8_t a = orc_dna_id(n1), b = orc_dna_id(n2), c = orc_dna_id(n3);
    if (a == 99 || b == 99 || c == 99) return 255;
    if (a == 4 || b == 4 || c == 4) return 64;
    return (uint8_t)(16 * a + 4 * b + c);
}

/* parallel_file_reader.hpp:61-99 alignment_t::update_seqs, the skip_bases rule only.
 * seq_len = seqs[0].size().  For '-' the caller must already have reverse-complemented the rows
 * (build_tracks.hpp:219-226).  Returns skip_bases; *new_start_pos gets the '+' strand's shifted
 * start (for '-' it is left at orig_start_pos, as in the reference). */
int64_t orc_skip_bases(uint64_t orig_start_pos, uint64_t chrom_len, uint64_t seq_len, char strand,
                       unsigned frame, uint64_t *new_start_pos) {
    int64_t skip;
    uint64_t start_pos = orig_start_pos;
    /* length() is seqs[0].size() - skip_bases with skip_bases == 0 at this point (:63). */
    uint64_t length = seq_len;
    if (strand == '+') {
        skip = ((int64_t)((uint64_t)frame - start_pos)) % 3;             /* :77 */
        if (skip < 0) skip += 3;
        if ((uint64_t)skip > length) skip = (int64_t)length;              /* :80-81 */
        start_pos += (uint64_t)skip;                                      /* :84 */
    } else {
        skip = ((int64_t)((uint64_t)frame - (chrom_len - (start_pos + length) + 2))) % 3; /* :92 */
        if (skip < 0) skip += 3;
        if ((uint64_t)skip > length) skip = (int64_t)length;
    }
    if (new_start_pos) *new_start_pos = start_pos;
    return skip;
}

/* parallel_file_reader.hpp:101-112: codon ids of one row from offset skip. out has (L-skip)/3 slots. */
void orc_translate_row(const char *seq, uint64_t L, uint64_t skip, uint8_t *out) {
    uint64_t K = (L - skip) / 3;
    for (uint64_t k = 0; k < K; ++k)
        out[k] = orc_codon_id(seq[skip + 3 * k], seq[skip + 3 * k + 1], seq[skip + 3 * k + 2]);
}

/* ------------------------------------------------------------------------------------------------
 * instance.hpp:648-685 compute_q_p14ns_and_q_scale_p14ns_fixed_mle.
 * S: symmetric exchangeabilities with zero diagonal (row-major 64x64), f: codon frequencies. */
void orc_build_q(const double *S, const double *f, double *Q) {
    double scale = 0.0;
    for (int i = 0; i < NS; ++i) {
        double sum = 0.0;
        for (int j = 0; j < NS; ++j) {
            double v = S[i * NS + j] * f[j];
            Q[i * NS + j] = v;
            sum -= v;
        }
        Q[i * NS + i] = sum;
        scale -= sum * f[i];
    }
    for (int i = 0; i < NS * NS; ++i) Q[i] = Q[i] / scale;
}

/* Cyclic Jacobi for a symmetric 64x64 matrix A (destroyed).  V's columns are eigenvectors. */
static void jacobi_sym(double *A, double *V, double *w) {
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j) V[i * NS + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < NS; ++p)
            for (int q = p + 1; q < NS; ++q) off += A[p * NS + q] * A[p * NS + q];
        if (off < 1e-300) break;
        for (int p = 0; p < NS - 1; ++p) {
            for (int q = p + 1; q < NS; ++q) {
                double apq = A[p * NS + q];
                if (fabs(apq) < 1e-310) continue;
                double app = A[p * NS + p], aqq = A[q * NS + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < NS; ++k) {
                    double akp = A[k * NS + p], akq = A[k * NS + q];
                    A[k * NS + p] = c * akp - s * akq;
                    A[k * NS + q] = s * akp + c * akq;
                }
                for (int k = 0; k < NS; ++k) {
                    double apk = A[p * NS + k], aqk = A[q * NS + k];
                    A[p * NS + k] = c * apk - s * aqk;
                    A[q * NS + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < NS; ++k) {
                    double vkp = V[k * NS + p], vkq = V[k * NS + q];
                    V[k * NS + p] = c * vkp - s * vkq;
                    V[k * NS + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    for (int i = 0; i < NS; ++i) w[i] = A[i * NS + i];
}

/* instance.hpp:309-434 instantiate_qs restated: Q = S_R diag(lambda) S_Rinv (all real; the reference
 * also keeps the real parts when check_real passes, :407-418).  See the header for the eigensolver
 * deviation.  f must be strictly positive. */
void orc_eigen(const double *Q, const double *f, double *lambda, double *SR, double *SRinv) {
    double *A = (double *)malloc(sizeof(double) * NS * NS);
    double *U = (double *)malloc(sizeof(double) * NS * NS);
    double sq[NS];
    for (int i = 0; i < NS; ++i) sq[i] = sqrt(f[i]);
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j) A[i * NS + j] = sq[i] * Q[i * NS + j] / sq[j];
    /* exact symmetrisation of rounding noise */
    for (int i = 0; i < NS; ++i)
        for (int j = i + 1; j < NS; ++j) {
            double m = 0.5 * (A[i * NS + j] + A[j * NS + i]);
            A[i * NS + j] = m;
            A[j * NS + i] = m;
        }
    jacobi_sym(A, U, lambda);
    for (int i = 0; i < NS; ++i)
        for (int k = 0; k < NS; ++k) {
            SR[i * NS + k] = U[i * NS + k] / sq[i];      /* right eigenvectors in columns */
            SRinv[k * NS + i] = U[i * NS + k] * sq[i];   /* left eigenvectors in rows */
        }
    free(A);
    free(U);
}

/* fixed_lik.hpp:281-360 get_prior, equilibrium branch (:323-346): row of S' at argmin |lambda|
 * (first strict minimum), normalised to sum 1. */
void orc_prior(const double *lambda, const double *SRinv, double *pi) {
    double minL = fabs(lambda[0]);
    int minp = 0;
    for (int i = 1; i < NS; ++i) {
        double m = fabs(lambda[i]);
        if (m < minL) { minL = m; minp = i; }
    }
    double mass = 0.0;
    for (int j = 0; j < NS; ++j) mass += SRinv[minp * NS + j];
    for (int j = 0; j < NS; ++j) pi[j] = SRinv[minp * NS + j] / mass;
}

/* instance.hpp:299-307 instantiate_tree: newick_elem::branch_length is a float; `elem.branch_length
 * *= factor` computes float(double(bl) * factor); PhyloModel_make then reads it back as double (:497). */
double orc_branch_time(float bl, double rho) {
    float scaled = (float)((double)bl * rho);
    return (double)scaled;
}

/* instance.hpp:487-640 PhyloModel_make, real-spectrum branch, one branch:
 *   P = S_R * (diag(exp(lambda t)) * S_Rinv)  (:529-549)
 * then per row (:602-640): negative entries with |x| <= 1e-6 -> 0, else error; diagonal := 1 - sum of
 * the (clamped) off-diagonals; error if the unclamped row sum differs from 1 by more than 1e-6.
 * Returns 0, or 1 for the "< 0" throw (:618), 2 for the row-sum throw (:635). */
int orc_pmatrix(const double *lambda, const double *SR, const double *SRinv, double t, double *P) {
    const double tol = 1e-6;
    double e[NS];
    double *B = (double *)malloc(sizeof(double) * NS * NS);
    for (int k = 0; k < NS; ++k) e[k] = exp(lambda[k] * t);
    for (int k = 0; k < NS; ++k)
        for (int j = 0; j < NS; ++j) B[k * NS + j] = SRinv[k * NS + j] * e[k];
    /* i-k-j loop order: every P[i][j] still accumulates its 64 products in ascending k */
    for (int i = 0; i < NS; ++i) {
        double *row = P + i * NS;
        for (int j = 0; j < NS; ++j) row[j] = 0.0;
        for (int k = 0; k < NS; ++k) {
            const double a = SR[i * NS + k];
            const double *b = B + k * NS;
            for (int j = 0; j < NS; ++j) row[j] += a * b[j];
        }
    }
    free(B);
    for (int i = 0; i < NS; ++i) {
        double total = 0.0, smii = 1.0;
        for (int j = 0; j < NS; ++j) {
            double cell = P[i * NS + j];
            total += cell;
            if (cell < 0.0) {
                if (fabs(cell) > tol) return 1;
                P[i * NS + j] = 0.0;
            }
            if (i != j) smii -= P[i * NS + j];
        }
        if (fabs(total - 1.0) > tol) return 2;
        P[i * NS + i] = smii;
    }
    return 0;
}

/* A prepared (tree, model, rho) evaluation context.  Mirrors what instance_t holds after
 * PhyloCSFModel_make + instantiate_tree(rho) + PhyloModel_make (run.hpp:38-39, fixed_lik.hpp:370-372). */
typedef struct {
    int nl, n;
    const int16_t *child1, *child2; /* newick_flatten ids (newick.hpp:218-229) */
    const float *bl;                /* newick_elem::branch_length */
    double lambda[NS], SR[NS * NS], SRinv[NS * NS], pi[NS];
    double *P;                      /* (n-1) x 64 x 64 */
} orc_model;

orc_model *orc_model_new(int nl, const int16_t *child1, const int16_t *child2, const float *bl,
                         const double *S, const double *f) {
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    double *Q = (double *)malloc(sizeof(double) * NS * NS);
    m->nl = nl;
    m->n = 2 * nl - 1;
    m->child1 = child1;
    m->child2 = child2;
    m->bl = bl;
    orc_build_q(S, f, Q);
    orc_eigen(Q, f, m->lambda, m->SR, m->SRinv);
    orc_prior(m->lambda, m->SRinv, m->pi);
    m->P = (double *)malloc(sizeof(double) * (size_t)(m->n - 1) * NS * NS);
    free(Q);
    return m;
}

void orc_model_free(orc_model *m) {
    if (!m) return;
    free(m->P);
    free(m);
}

void orc_model_get(const orc_model *m, double *lambda, double *SR, double *SRinv, double *pi) {
    memcpy(lambda, m->lambda, sizeof m->lambda);
    memcpy(SR, m->SR, sizeof m->SR);
    memcpy(SRinv, m->SRinv, sizeof m->SRinv);
    memcpy(pi, m->pi, sizeof m->pi);
}

/* fixed_lik.hpp:370-372: instantiate_tree(rho) + PhyloModel_make — all n-1 branch matrices. */
int orc_model_set_rho(orc_model *m, double rho) {
    for (int b = 0; b < m->n - 1; ++b) {
        int rc = orc_pmatrix(m->lambda, m->SR, m->SRinv, orc_branch_time(m->bl[b], rho),
                             m->P + (size_t)b * NS * NS);
        if (rc) return rc;
    }
    return 0;
}

const double *orc_model_pmatrices(const orc_model *m) { return m->P; }

/* fixed_lik.hpp:105-123 dot_with_alpha + :125-164 ensure_alpha for ONE codon column.
 * leaves[nl] are codon ids 0..64.  alpha is scratch of (n-nl) x 64 doubles.  Returns z; alpha_root
 * (may be NULL) receives the root partial. */
double orc_prune_column(const orc_model *m, const uint8_t *leaves, double *alpha, double *alpha_root) {
    const int nl = m->nl, n = m->n;
    for (int i = nl; i < n; ++i) {
        const int lc = m->child1[i], rc = m->child2[i];
        const double *ls = m->P + (size_t)lc * NS * NS, *rs = m->P + (size_t)rc * NS * NS;
        double *out = alpha + (size_t)(i - nl) * NS;
        for (int a = 0; a < NS; ++a) {
            double r1, r2;
            if (lc >= nl) {
                const double *al = alpha + (size_t)(lc - nl) * NS;
                r1 = 0.0;
                for (int j = 0; j < NS; ++j) r1 += ls[a * NS + j] * al[j];
            } else if (leaves[lc] == 64) {
                r1 = 0.0;
                for (int j = 0; j < NS; ++j) r1 += ls[a * NS + j];
            } else {
                r1 = ls[a * NS + leaves[lc]];
            }
            if (rc >= nl) {
                const double *ar = alpha + (size_t)(rc - nl) * NS;
                r2 = 0.0;
                for (int j = 0; j < NS; ++j) r2 += rs[a * NS + j] * ar[j];
            } else if (leaves[rc] == 64) {
                r2 = 0.0;
                for (int j = 0; j < NS; ++j) r2 += rs[a * NS + j];
            } else {
                r2 = rs[a * NS + leaves[rc]];
            }
            out[a] = r1 * r2;
        }
    }
    const double *root = alpha + (size_t)(n - 1 - nl) * NS;
    double z = 0.0;
    for (int a = 0; a < NS; ++a) z += m->pi[a] * root[a]; /* :159-161, beta row n-1 = prior (:424-427) */
    if (alpha_root) memcpy(alpha_root, root, sizeof(double) * NS);
    return z;
}

/* fixed_lik.hpp:362-449 lpr_leaves at the model's current rho.
 * peptides: nl rows x K codon ids, row stride `stride`.  lpr_per_codon / anc_per_codon may be NULL.
 * The anc term follows :435-444 + node_posterior :215-246 (root: alpha*prior/z; z == 0 -> zeros). */
void orc_lpr_leaves(const orc_model *m, const uint8_t *peptides, int64_t K, int64_t stride,
                    int compute_anc, double *lpr, double *elpr_anc, double *lpr_per_codon,
                    double *anc_per_codon) {
    const int nl = m->nl, n = m->n;
    double *alpha = (double *)malloc(sizeof(double) * (size_t)(n - nl) * NS);
    uint8_t *col = (uint8_t *)malloc((size_t)nl);
    double lprior[NS], root[NS];
    for (int x = 0; x < NS; ++x) lprior[x] = log(m->pi[x]);
    double s = 0.0, sa = 0.0;
    for (int64_t k = 0; k < K; ++k) {
        for (int sp = 0; sp < nl; ++sp) col[sp] = peptides[(int64_t)sp * stride + k];
        double z = orc_prune_column(m, col, alpha, root);
        double lz = log(z);
        s += lz;
        if (lpr_per_codon) lpr_per_codon[k] = lz;
        if (compute_anc) {
            double e = 0.0;
            if (z != 0.0)
                for (int x = 0; x < NS; ++x) e += lprior[x] * (root[x] * m->pi[x] / z);
            sa += e;
            if (anc_per_codon) anc_per_codon[k] = e;
        }
    }
    *lpr = s;
    *elpr_anc = sa;
    free(alpha);
    free(col);
}

/* ------------------------------------------------------------------------------------------------
 * std::mt19937 + libstdc++ uniform_real_distribution<double>(0, width) as used by fit_find_init
 * (fixed_lik.hpp:484-490, seeded 42 per alignment in score_msa.hpp:115). */
typedef struct { uint32_t mt[624]; int idx; } orc_mt19937;

void orc_mt_seed(orc_mt19937 *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937 *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ generate_canonical<double,53>: two 32-bit draws, sum = d1 + d2*2^32 (rounded to double),
 * / 2^64, clamped below 1; then * width + 0. */
double orc_uniform(orc_mt19937 *g, double width) {
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k) {
        sum += (double)orc_mt_next(g) * tmp;
        tmp *= 4294967296.0;
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret * width + 0.0;
}

/* OMEGA strategy state (run.hpp:59-182): q_settings = kappa, omega, sigma, 9 F3x4 ratios; tree_settings = rho. */
typedef struct {
    double qs[12];
    double rho;
} orc_omega_state;

static int omega_set_q(orc_model *m, const double *qs);
static double omega_lpr_rho(double rho);
static double omega_lpr_kappa(double kappa);

typedef struct {
    orc_model *m;
    const uint8_t *peptides;
    int64_t K, stride;
    int compute_anc;
    double x, lpr, elpr_anc;
    int status;  /* first PhyloModel_make error (the reference throws std::runtime_error) */
    int evals;
    int kind;    /* 0: lpr_leaves(rho) of a fixed Q (MLE); 1: OMEGA rho fit; 2: OMEGA kappa fit */
    orc_omega_state *om;
} orc_fit;

/* fixed_lik.hpp:460-467 minimizer_lpr_leaves (kind 0): returns -lpr, records x/lpr/elpr_anc.
 * omega.hpp:205-233 minimizer_lpr_leaves_rho / _kappa (kinds 1, 2): the same with the half-Cauchy / Gamma log-prior
 * added; the kappa variant rebuilds Q and its eigensystem first and keeps the last evaluated rho. */
static double fit_eval(orc_fit *p, double x) {
    p->x = x;
    p->evals++;
    int rc;
    if (p->kind == 2) {
        p->om->qs[0] = x;
        omega_set_q(p->m, p->om->qs);
        rc = orc_model_set_rho(p->m, p->om->rho);
    } else {
        if (p->kind == 1) p->om->rho = x;
        rc = orc_model_set_rho(p->m, x);
    }
    if (rc && !p->status) p->status = rc;
    orc_lpr_leaves(p->m, p->peptides, p->K, p->stride, p->compute_anc, &p->lpr, &p->elpr_anc, NULL, NULL);
    if (p->kind == 1) p->lpr += omega_lpr_rho(x);
    if (p->kind == 2) p->lpr += omega_lpr_kappa(x);
    return -p->lpr;
}

/* fixed_lik.hpp:469-544 fit_find_init + max_lik_lpr_leaves with gsl_min_fminimizer_brent restated
 * (GSL 2.x min/brent.c + min/fsolver.c).  Returns 0 or the first P-matrix error.
 * On error the reference would have thrown at that evaluation; the caller maps that to NaN. */
static int max_lik_core(orc_fit *pp_, double init, double lo, double hi, orc_mt19937 *gen, double *lpr, double *elpr_anc,
                        double *x_final, int *n_evals);

int orc_max_lik(orc_model *m, const uint8_t *peptides, int64_t K, int64_t stride, int compute_anc,
                double init, double lo, double hi, orc_mt19937 *gen, double *lpr, double *elpr_anc,
                double *x_final, int *n_evals) {
    orc_fit p = {m, peptides, K, stride, compute_anc, 0.0, 0.0, 0.0, 0, 0, 0, NULL};
    return max_lik_core(&p, init, lo, hi, gen, lpr, elpr_anc, x_final, n_evals);
}

static int max_lik_core(orc_fit *pp_, double init, double lo, double hi, orc_mt19937 *gen, double *lpr, double *elpr_anc,
                        double *x_final, int *n_evals) {
#define p (*pp_)
    const double width = log(hi) - log(lo);
    const double flo = -fit_eval(&p, lo);
    const double fhi = -fit_eval(&p, hi);
    double x = init;
    double fx = -fit_eval(&p, init);
    int i = 0;
    while (i < 250 && (fx <= flo || fx <= fhi)) {
        const double r = orc_uniform(gen, width);
        x = exp(log(lo) + r);
        fx = -fit_eval(&p, x);
        ++i;
    }
    if (i == 250) {
        if (flo > fhi) { p.x = lo; p.lpr = flo; } else { p.x = hi; p.lpr = fhi; }
    }
    fit_eval(&p, p.x);
    if (p.status) goto done;

    if (lo < p.x && p.x < hi) {
        /* gsl_min_fminimizer_set: f at x_minimum, x_lower, x_upper */
        double z = p.x, xl = lo, xu = hi;
        double fl = fit_eval(&p, xl), fu = fit_eval(&p, xu), fz = fit_eval(&p, z);
        (void)fl; (void)fu;
        /* brent_init */
        const double golden = 0.3819660;
        double v = xl + golden * (xu - xl), w = v, d = 0.0, e = 0.0;
        double fv = fit_eval(&p, v), fw = fv;
        int max_iter = 250;
        do {
            /* brent_iterate (note: d and e are swapped on load, as in GSL) */
            double dd = e, ee = d;
            const double tol = 1.4901161193847656e-08 * fabs(z);
            const double mid = 0.5 * (xl + xu);
            const double w_lower = z - xl, w_upper = xu - z;
            double pp = 0, q = 0, r = 0, u, f_u;
            if (fabs(ee) > tol) {
                r = (z - w) * (fz - fv);
                q = (z - v) * (fz - fw);
                pp = (z - v) * q - (z - w) * r;
                q = 2 * (q - r);
                if (q > 0) pp = -pp; else q = -q;
                r = ee;
                ee = dd;
            }
            if (fabs(pp) < fabs(0.5 * q * r) && pp < q * w_lower && pp < q * w_upper) {
                double t2 = 2 * tol;
                dd = pp / q;
                u = z + dd;
                if ((u - xl) < t2 || (xu - u) < t2) dd = (z < mid) ? tol : -tol;
            } else {
                ee = (z < mid) ? xu - z : -(z - xl);
                dd = golden * ee;
            }
            if (fabs(dd) >= tol) u = z + dd; else u = z + ((dd > 0) ? tol : -tol);
            e = ee;
            d = dd;
            f_u = fit_eval(&p, u);
            if (f_u <= fz) {
                if (u < z) { xu = z; } else { xl = z; }
                v = w; fv = fw;
                w = z; fw = fz;
                z = u; fz = f_u;
            } else {
                if (u < z) { xl = u; } else { xu = u; }
                if (f_u <= fw || w == z) {
                    v = w; fv = fw;
                    w = u; fw = f_u;
                } else if (f_u <= fv || v == z || v == w) {
                    v = u; fv = f_u;
                }
            }
            if (((xu - xl) / z) <= 0.01) break; /* fixed_lik.hpp:533 */
            --max_iter;
        } while (max_iter > 0);
    }
done:
    *lpr = p.lpr;
    *elpr_anc = p.elpr_anc;
    if (x_final) *x_final = p.x;
    if (n_evals) *n_evals = p.evals;
    return p.status;
#undef p
}

/* ------------------------------------------------------------------------------------------------
 * OMEGA strategy (score-msa --strategy omega): run.hpp:59-182 + omega.hpp.
 * translation.hpp: the standard genetic code indexed by 16a+4b+c over A,C,G,T. */
static char omega_aa(int codon) {
    static const char tcag[] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    static const int to_tcag[4] = {2, 1, 3, 0}; /* A C G T -> position in T C A G */
    return tcag[16 * to_tcag[codon / 16] + 4 * to_tcag[(codon / 4) % 4] + to_tcag[codon % 4]];
}

/* omega.hpp:8-19 pi_expr_sc */
static double omega_pi_sc(const double *qs, int codon) {
    const int i1 = codon / 16, i2 = (codon - 16 * i1) / 4, i3 = codon - 16 * i1 - 4 * i2;
    const double f1 = ((i1 == 3) ? 1.0 : qs[3 + i1]) / (1.0 + qs[3] + qs[4] + qs[5]);
    const double f2 = ((i2 == 3) ? 1.0 : qs[6 + i2]) / (1.0 + qs[6] + qs[7] + qs[8]);
    const double f3 = ((i3 == 3) ? 1.0 : qs[9 + i3]) / (1.0 + qs[9] + qs[10] + qs[11]);
    return f1 * f2 * f3;
}

/* omega.hpp:21-36 pi_expr + :38-95 comp_q_p14n + :97-128 scale; then instantiate_qs (instance.hpp:309-434) and the
 * equilibrium prior (fixed_lik.hpp:323-346).  pi is a positive multiple of the stationary distribution, which is all the
 * symmetrisation in orc_eigen needs. */
static int omega_set_q(orc_model *m, const double *qs) {
    double pi[NS];
    double *Q = (double *)malloc(sizeof(double) * NS * NS);
    const double kappa = qs[0], omega = qs[1], sigma = qs[2];
    const double denom = 1.0 - ((1.0 - sigma) * (omega_pi_sc(qs, 3 * 16 + 0 * 4 + 0) + omega_pi_sc(qs, 3 * 16 + 0 * 4 + 2) +
                                                 omega_pi_sc(qs, 3 * 16 + 2 * 4 + 0)));
    for (int i = 0; i < NS; ++i) pi[i] = omega_pi_sc(qs, i) / denom;
    for (int i = 0; i < NS; ++i) {
        const int i1 = i / 16, i2 = (i / 4) % 4, i3 = i % 4;
        const char iaa = omega_aa(i);
        for (int j = 0; j < NS; ++j) {
            const int j1 = j / 16, j2 = (j / 4) % 4, j3 = j % 4;
            double val = 0.0;
            if ((i1 != j1) + (i2 != j2) + (i3 != j3) == 1) {
                int transition = 0;
                if (i1 != j1 && (i1 + j1 == 2 || i1 + j1 == 4)) transition = 1;
                if (i2 != j2 && (i2 + j2 == 2 || i2 + j2 == 4)) transition = 1;
                if (i3 != j3 && (i3 + j3 == 2 || i3 + j3 == 4)) transition = 1;
                val = transition ? kappa : 1.0;
                const char jaa = omega_aa(j);
                val *= (iaa != '*' && jaa != '*' && iaa != jaa) ? omega : 1.0;
                val *= pi[j];
            }
            Q[i * NS + j] = val;
        }
    }
    for (int i = 0; i < NS; ++i) {
        double val = 0.0;
        for (int j = 0; j < NS; ++j)
            if (i != j) val -= Q[i * NS + j];
        Q[i * NS + i] = val;
    }
    double factor = 0.0;
    for (int i = 0; i < NS; ++i) factor -= pi[i] * Q[i * NS + i];
    for (int i = 0; i < NS * NS; ++i) Q[i] = Q[i] / factor;
    orc_eigen(Q, pi, m->lambda, m->SR, m->SRinv);
    orc_prior(m->lambda, m->SRinv, m->pi);
    free(Q);
    return 0;
}

/* omega.hpp:130-141 get_lpr_rho: half-Cauchy(mode 1, scale 0.5) log-density */
static double omega_lpr_rho(double rho) {
    const double mode = 1.0, scale = 0.5;
    const double numer = 1.0 / (M_PI * scale * (1.0 + pow(((rho - mode) / scale), 2.0)));
    const double cauchy_cdf = atan((0.0 - mode) / scale) / M_PI + 0.5;
    const double denom = 1.0 - cauchy_cdf;
    return log(numer) - log(denom);
}

/* omega.hpp:143-149 get_lpr_kappa: log Gamma(shape 7, scale 0.25) density at kappa - 1 + DBL_EPSILON
 * (gsl_ran_gamma_pdf: exp((a-1) log(x/b) - x/b - lgamma(a)) / b) */
static double omega_lpr_kappa(double kappa) {
    const double k = kappa - 1.0 + 2.2204460492503131e-16;
    const double a = 7.0, b = 0.25;
    double g;
    if (k < 0) g = 0;
    else if (k == 0) g = 0;
    else g = exp((a - 1) * log(k / b) - k / b - lgamma(a)) / b;
    return log(g);
}

/* run.hpp:59-182, the OMEGA branch of run().  peptides: nl x K codon ids of frame +1 from offset 0.
 * Returns 0 or the first PhyloModel_make error (the reference throws -> NaN row); *score = 10 (lpr_H1 - lpr_H0) / ln 10
 * narrowed to float by the caller.  info (may be NULL): [0] rho, [1] kappa at the end, [2] lpr_H0, [3] lpr_H1, [4] evaluations.
 * Not restated: the very first PhyloModel_make on the uniform-frequency Q (run.hpp:94-103), whose only observable effect is
 * a throw when GSL's nonsymmetric solver returns a bad basis for that highly degenerate matrix. */
int orc_omega(int nl, const int16_t *child1, const int16_t *child2, const float *bl, const uint8_t *peptides, int64_t K,
              int64_t stride, uint32_t seed, double *score, double *info) {
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->nl = nl; m->n = 2 * nl - 1; m->child1 = child1; m->child2 = child2; m->bl = bl;
    m->P = (double *)malloc(sizeof(double) * (size_t)(m->n - 1) * NS * NS);
    orc_omega_state om;
    om.qs[0] = 2.5; om.qs[1] = 1.0; om.qs[2] = 1.0;
    om.rho = 1.0;
    /* update_f3x4 (run.hpp:106-134): pseudo-count 1, certain codons only */
    double counts[3][4];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) counts[i][j] = 1.0;
    for (int sp = 0; sp < nl; ++sp)
        for (int64_t k = 0; k < K; ++k) {
            const uint8_t c = peptides[(int64_t)sp * stride + k];
            if (c != 64) { counts[0][c / 16] += 1; counts[1][(c / 4) % 4] += 1; counts[2][c % 4] += 1; }
        }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) om.qs[3 + 3 * i + j] = counts[i][j] / counts[i][3];
    orc_mt19937 gen;
    orc_mt_seed(&gen, seed);
    int status = 0, evals = 0;
    double lpr[2] = {0.0, 0.0}, anc = 0.0;
    for (int h = 0; h < 2 && !status; ++h) {
        if (h == 1) { om.qs[1] = 0.2; om.qs[2] = 0.01; }
        omega_set_q(m, om.qs);
        status = orc_model_set_rho(m, om.rho);          /* PhyloModel_make(inst, NULL, true) at run.hpp:139 / :167 */
        for (int r = 0; r < 3 && !status; ++r) {
            orc_fit p = {m, peptides, K, stride, 0, 0.0, 0.0, 0.0, 0, 0, 1, &om};
            status = max_lik_core(&p, om.rho, 0.001, 10.0, &gen, &lpr[h], &anc, NULL, NULL);
            evals += p.evals;
            if (status) break;
            orc_fit pk = {m, peptides, K, stride, 0, 0.0, 0.0, 0.0, 0, 0, 2, &om};
            status = max_lik_core(&pk, om.qs[0], 1.0, 10.0, &gen, &lpr[h], &anc, NULL, NULL);
            evals += pk.evals;
        }
    }
    *score = 10.0 * (lpr[1] - lpr[0]) / log(10.0);
    if (info) { info[0] = om.rho; info[1] = om.qs[0]; info[2] = lpr[0]; info[3] = lpr[1]; info[4] = (double)evals; }
    free(m->P);
    free(m);
    return status;
}

/* ------------------------------------------------------------------------------------------------
 * additional_scores.hpp:5-41 newick_sum_branch_lengths on the flattened tree.
 * present[leaf] != 0 marks the subset.  bl64 are the pointer tree's double branch lengths indexed by
 * flattened node id.  Summation order is the reference's: own, then += left, then += right. */
static int overlap_size(const int16_t *c1, const int16_t *c2, int node, const uint8_t *present) {
    if (c1[node] < 0) return present[node] ? 1 : 0;
    return overlap_size(c1, c2, c1[node], present) + overlap_size(c1, c2, c2[node], present);
}

static double sum_bl(const int16_t *c1, const int16_t *c2, const double *bl64, int node,
                     const uint8_t *present, int arrived, int overlap_parent) {
    if (c1[node] < 0) return bl64[node];
    if (overlap_parent == -1) overlap_parent = overlap_size(c1, c2, node, present);
    const int ol = overlap_size(c1, c2, c1[node], present);
    const int orr = overlap_parent - ol;
    double bl = 0.0;
    if (arrived) bl = bl64[node];
    if (ol > 0 && orr > 0) arrived = 1;
    if (ol > 0) bl += sum_bl(c1, c2, bl64, c1[node], present, arrived, ol);
    if (orr > 0) bl += sum_bl(c1, c2, bl64, c2[node], present, arrived, orr);
    return bl;
}

/* additional_scores.hpp:43-84 compute_bls_score.  seqs: nl rows x L ASCII, row stride `stride`.
 * per_base (may be NULL) gets L entries.  Returns bl_total / (bl(all) * L).  *bad_char is set to 1 if a
 * character outside "ACGTacgt.-Nn" is met (the reference exit(37)s). */
double orc_bls(int nl, const int16_t *c1, const int16_t *c2, const double *bl64, const char *seqs,
               int64_t L, int64_t stride, double *per_base, int *bad_char) {
    const int root = 2 * nl - 2;
    uint8_t *present = (uint8_t *)malloc((size_t)nl);
    memset(present, 1, (size_t)nl);
    const double all = sum_bl(c1, c2, bl64, root, present, 0, -1);
    double total = 0.0;
    for (int64_t i = 0; i < L; ++i) {
        int cnt = 0;
        for (int s = 0; s < nl; ++s) {
            uint8_t id = orc_dna_id(seqs[(int64_t)s * stride + i]);
            if (id == 99 && bad_char) *bad_char = 1;
            present[s] = id <= 3;
            cnt += present[s];
        }
        if (cnt >= 2) {
            double bl = sum_bl(c1, c2, bl64, root, present, 0, -1);
            total += bl;
            if (per_base) per_base[i] = bl / all;
        } else if (per_base) {
            per_base[i] = 0.0;
        }
    }
    free(present);
    return total / (all * (double)L);
}

/* ------------------------------------------------------------------------------------------------
 * Window formulation used by the CUDA path: for every forward offset
 * o in [0, L-3], the '+' codon n[o..o+2] and the '-' codon comp(n[o+2]),comp(n[o+1]),comp(n[o]).
 * plus/minus: nl rows x (L-2) codon ids, row stride L-2.  This is a convenience for tests; its
 * equivalence to update_seqs + reverse-complement is itself checked in tests/test_oracle_frames.py. */
void orc_window_codons(const char *seqs, int nl, int64_t L, int64_t stride, uint8_t *plus, uint8_t *minus) {
    const int64_t W = L - 2;
    for (int s = 0; s < nl; ++s) {
        const char *r = seqs + (int64_t)s * stride;
        for (int64_t o = 0; o < W; ++o) {
            plus[(int64_t)s * W + o] = orc_codon_id(r[o], r[o + 1], r[o + 2]);
            minus[(int64_t)s * W + o] = orc_codon_id(orc_complement(r[o + 2]), orc_complement(r[o + 1]),
                                                      orc_complement(r[o]));
        }
    }
}

/* run.hpp:35-55 run_tracks for codon ids already in hand: per-codon decibans
 * 10 (lprC - lprNC) / ln 10.  Both models must be at rho = 1. */
void orc_run_tracks(const orc_model *mc, const orc_model *mnc, const uint8_t *peptides, int64_t K,
                    int64_t stride, double *decibans) {
    double lc, lnc, ac, anc;
    double *nc = (double *)malloc(sizeof(double) * (size_t)(K > 0 ? K : 1));
    orc_lpr_leaves(mc, peptides, K, stride, 0, &lc, &ac, decibans, NULL);
    orc_lpr_leaves(mnc, peptides, K, stride, 0, &lnc, &anc, nc, NULL);
    for (int64_t k = 0; k < K; ++k) decibans[k] = 10.0 * (decibans[k] - nc[k]) / log(10.0);
    free(nc);
}
