/*
 * gsl_shim.h — a from-scratch stand-in for the handful of GNU Scientific Library entry points that
 * PhyloCSF++'s likelihood core calls (SURVEY.md §8c lists the call sites).
 *
 * TEST INFRASTRUCTURE ONLY.  GSL is a system dependency of the reference (CMakeLists.txt:43, unpinned) and is
 * absent from this image.  With this header set on the include path the reference's UNMODIFIED sources
 * (run.hpp, instance.hpp, fixed_lik.hpp, omega.hpp, additional_scores.hpp, phylocsf++build_tracks.hpp,
 * phylocsf++score_msa.hpp, ...) compile where they lie under /root/reference/src into oracle/_ref/ — the CPU
 * baseline (`cpu_baseline.kind == "reference"`) and the generator of golden files for inputs the reference's own
 * tests do not cover.  Nothing under phylocsfpp_b200/ includes or links this.
 *
 * Written from the published GSL API / algorithms, not from GSL sources:
 *   - containers: row-major blocks with {size1,size2,tda,data} / {size,stride,data}; views alias the parent;
 *   - gsl_blas_ddot / dgemm / zgemm: the reference-BLAS loop orders (sequential dot; i-k-j gemm);
 *   - gsl_eigen_nonsymmv: Householder reduction to Hessenberg form + Francis double-shift QR with accumulated
 *     transformations + back-substitution for the eigenvectors (the EISPACK orthes/hqr2 pair, as GSL documents),
 *     eigenvectors normalised to unit 2-norm;
 *   - gsl_linalg_complex_LU_decomp/invert: partial-pivoting LU;
 *   - gsl_min_fminimizer_brent: the golden-section / parabolic step of GSL 2.x min/brent.c with min/fsolver.c's
 *     set/iterate protocol (reproduces the reference's MLE goldens to 6 decimals);
 *   - gsl_ran_gamma_pdf, gsl_sf_exp, gsl_complex_{abs,rect,exp,mul}.
 * P(t) = S exp(Lt) S^-1 is independent of the eigenvector scaling and ordering, so shim-vs-GSL differences stay
 * ~1e-13 — far below the printed precision of every golden file (checked: tests/test_reference_build.py).
 */
#ifndef PCSF_GSL_SHIM_H
#define PCSF_GSL_SHIM_H

#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

/* ------------------------------------------------------------------------------------------------ containers */
typedef struct { size_t size; size_t stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; void *block; int owner; } gsl_matrix;
typedef struct { double dat[2]; } gsl_complex;
typedef struct { size_t size; size_t stride; double *data; void *block; int owner; } gsl_vector_complex;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; void *block; int owner; } gsl_matrix_complex;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct { size_t size; size_t *data; } gsl_permutation;

#define GSL_REAL(z) ((z).dat[0])
#define GSL_IMAG(z) ((z).dat[1])
#define GSL_SUCCESS 0
#define GSL_CONTINUE (-2)
#define GSL_EINVAL 4
#define GSL_FAILURE (-1)

static const gsl_complex GSL_COMPLEX_ONE = {{1.0, 0.0}};
static const gsl_complex GSL_COMPLEX_ZERO = {{0.0, 0.0}};

static inline void gsl_shim_die(const char *what) {
    std::fprintf(stderr, "gsl shim: %s\n", what);
    std::abort(); /* GSL's default error handler aborts as well */
}

static inline gsl_vector *gsl_vector_alloc(size_t n) {
    gsl_vector *v = (gsl_vector *)std::malloc(sizeof(gsl_vector));
    v->size = n; v->stride = 1; v->data = (double *)std::malloc(sizeof(double) * (n ? n : 1)); v->block = v->data; v->owner = 1;
    return v;
}
static inline void gsl_vector_free(gsl_vector *v) { if (!v) return; if (v->owner) std::free(v->data); std::free(v); }
static inline double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i * v->stride] = x; }
static inline void gsl_vector_set_all(gsl_vector *v, double x) { for (size_t i = 0; i < v->size; ++i) v->data[i * v->stride] = x; }
static inline void gsl_vector_set_zero(gsl_vector *v) { gsl_vector_set_all(v, 0.0); }
static inline int gsl_vector_memcpy(gsl_vector *d, const gsl_vector *s) {
    if (d->size != s->size) gsl_shim_die("gsl_vector_memcpy: sizes differ");
    for (size_t i = 0; i < s->size; ++i) d->data[i * d->stride] = s->data[i * s->stride];
    return GSL_SUCCESS;
}

static inline gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) {
    gsl_matrix *m = (gsl_matrix *)std::malloc(sizeof(gsl_matrix));
    m->size1 = n1; m->size2 = n2; m->tda = n2;
    m->data = (double *)std::malloc(sizeof(double) * (n1 * n2 ? n1 * n2 : 1)); m->block = m->data; m->owner = 1;
    return m;
}
static inline void gsl_matrix_free(gsl_matrix *m) { if (!m) return; if (m->owner) std::free(m->data); std::free(m); }
static inline double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
static inline void gsl_matrix_set_zero(gsl_matrix *m) {
    for (size_t i = 0; i < m->size1; ++i) for (size_t j = 0; j < m->size2; ++j) m->data[i * m->tda + j] = 0.0;
}
static inline int gsl_matrix_memcpy(gsl_matrix *d, const gsl_matrix *s) {
    if (d->size1 != s->size1 || d->size2 != s->size2) gsl_shim_die("gsl_matrix_memcpy: sizes differ");
    for (size_t i = 0; i < s->size1; ++i) for (size_t j = 0; j < s->size2; ++j) d->data[i * d->tda + j] = s->data[i * s->tda + j];
    return GSL_SUCCESS;
}
static inline int gsl_matrix_scale(gsl_matrix *m, double x) {
    for (size_t i = 0; i < m->size1; ++i) for (size_t j = 0; j < m->size2; ++j) m->data[i * m->tda + j] *= x;
    return GSL_SUCCESS;
}
static inline gsl_vector_view gsl_matrix_row(gsl_matrix *m, size_t i) {
    gsl_vector_view v; v.vector.size = m->size2; v.vector.stride = 1; v.vector.data = m->data + i * m->tda;
    v.vector.block = m->block; v.vector.owner = 0; return v;
}
static inline gsl_matrix_view gsl_matrix_submatrix(gsl_matrix *m, size_t k1, size_t k2, size_t n1, size_t n2) {
    gsl_matrix_view v; v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = m->tda;
    v.matrix.data = m->data + k1 * m->tda + k2; v.matrix.block = m->block; v.matrix.owner = 0; return v;
}

static inline gsl_vector_complex *gsl_vector_complex_alloc(size_t n) {
    gsl_vector_complex *v = (gsl_vector_complex *)std::malloc(sizeof(gsl_vector_complex));
    v->size = n; v->stride = 1; v->data = (double *)std::malloc(sizeof(double) * 2 * (n ? n : 1)); v->block = v->data; v->owner = 1;
    return v;
}
static inline void gsl_vector_complex_free(gsl_vector_complex *v) { if (!v) return; if (v->owner) std::free(v->data); std::free(v); }
static inline gsl_complex gsl_vector_complex_get(const gsl_vector_complex *v, size_t i) {
    gsl_complex z; z.dat[0] = v->data[2 * i * v->stride]; z.dat[1] = v->data[2 * i * v->stride + 1]; return z;
}
static inline void gsl_vector_complex_set(gsl_vector_complex *v, size_t i, gsl_complex z) {
    v->data[2 * i * v->stride] = z.dat[0]; v->data[2 * i * v->stride + 1] = z.dat[1];
}
static inline int gsl_vector_complex_memcpy(gsl_vector_complex *d, const gsl_vector_complex *s) {
    if (d->size != s->size) gsl_shim_die("gsl_vector_complex_memcpy: sizes differ");
    for (size_t i = 0; i < s->size; ++i) gsl_vector_complex_set(d, i, gsl_vector_complex_get(s, i));
    return GSL_SUCCESS;
}
static inline gsl_matrix_complex *gsl_matrix_complex_alloc(size_t n1, size_t n2) {
    gsl_matrix_complex *m = (gsl_matrix_complex *)std::malloc(sizeof(gsl_matrix_complex));
    m->size1 = n1; m->size2 = n2; m->tda = n2;
    m->data = (double *)std::malloc(sizeof(double) * 2 * (n1 * n2 ? n1 * n2 : 1)); m->block = m->data; m->owner = 1;
    return m;
}
static inline void gsl_matrix_complex_free(gsl_matrix_complex *m) { if (!m) return; if (m->owner) std::free(m->data); std::free(m); }
static inline gsl_complex gsl_matrix_complex_get(const gsl_matrix_complex *m, size_t i, size_t j) {
    gsl_complex z; z.dat[0] = m->data[2 * (i * m->tda + j)]; z.dat[1] = m->data[2 * (i * m->tda + j) + 1]; return z;
}
static inline void gsl_matrix_complex_set(gsl_matrix_complex *m, size_t i, size_t j, gsl_complex z) {
    m->data[2 * (i * m->tda + j)] = z.dat[0]; m->data[2 * (i * m->tda + j) + 1] = z.dat[1];
}
static inline int gsl_matrix_complex_memcpy(gsl_matrix_complex *d, const gsl_matrix_complex *s) {
    if (d->size1 != s->size1 || d->size2 != s->size2) gsl_shim_die("gsl_matrix_complex_memcpy: sizes differ");
    for (size_t i = 0; i < s->size1; ++i) for (size_t j = 0; j < s->size2; ++j) gsl_matrix_complex_set(d, i, j, gsl_matrix_complex_get(s, i, j));
    return GSL_SUCCESS;
}

static inline gsl_permutation *gsl_permutation_alloc(size_t n) {
    gsl_permutation *p = (gsl_permutation *)std::malloc(sizeof(gsl_permutation));
    p->size = n; p->data = (size_t *)std::malloc(sizeof(size_t) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) p->data[i] = i;
    return p;
}
static inline void gsl_permutation_free(gsl_permutation *p) { if (!p) return; std::free(p->data); std::free(p); }

/* ------------------------------------------------------------------------------------------------ complex math */
static inline gsl_complex gsl_complex_rect(double x, double y) { gsl_complex z; z.dat[0] = x; z.dat[1] = y; return z; }
static inline double gsl_complex_abs(gsl_complex z) { return std::hypot(z.dat[0], z.dat[1]); }
static inline gsl_complex gsl_complex_mul(gsl_complex a, gsl_complex b) {
    return gsl_complex_rect(a.dat[0] * b.dat[0] - a.dat[1] * b.dat[1], a.dat[0] * b.dat[1] + a.dat[1] * b.dat[0]);
}
static inline gsl_complex gsl_complex_exp(gsl_complex a) {
    const double rho = std::exp(a.dat[0]);
    return gsl_complex_rect(rho * std::cos(a.dat[1]), rho * std::sin(a.dat[1]));
}
static inline double gsl_sf_exp(double x) { return std::exp(x); }

/* ------------------------------------------------------------------------------------------------ BLAS subset */
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
typedef enum CBLAS_TRANSPOSE CBLAS_TRANSPOSE_t;

static inline int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result) {
    if (x->size != y->size) gsl_shim_die("gsl_blas_ddot: sizes differ");
    double r = 0.0;
    for (size_t i = 0; i < x->size; ++i) r += x->data[i * x->stride] * y->data[i * y->stride];
    *result = r;
    return GSL_SUCCESS;
}
static inline int gsl_blas_dgemm(CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, double alpha, const gsl_matrix *A, const gsl_matrix *B,
                                 double beta, gsl_matrix *C) {
    if (ta != CblasNoTrans || tb != CblasNoTrans) gsl_shim_die("gsl_blas_dgemm: only NoTrans/NoTrans is implemented");
    if (A->size2 != B->size1 || C->size1 != A->size1 || C->size2 != B->size2) gsl_shim_die("gsl_blas_dgemm: shapes");
    for (size_t i = 0; i < C->size1; ++i)
        for (size_t j = 0; j < C->size2; ++j) C->data[i * C->tda + j] = (beta == 0.0) ? 0.0 : beta * C->data[i * C->tda + j];
    for (size_t i = 0; i < A->size1; ++i)
        for (size_t k = 0; k < A->size2; ++k) {
            const double t = alpha * A->data[i * A->tda + k];
            if (t != 0.0)
                for (size_t j = 0; j < B->size2; ++j) C->data[i * C->tda + j] += t * B->data[k * B->tda + j];
        }
    return GSL_SUCCESS;
}
static inline int gsl_blas_zgemm(CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, const gsl_complex alpha, const gsl_matrix_complex *A,
                                 const gsl_matrix_complex *B, const gsl_complex beta, gsl_matrix_complex *C) {
    if (ta != CblasNoTrans || tb != CblasNoTrans) gsl_shim_die("gsl_blas_zgemm: only NoTrans/NoTrans is implemented");
    if (A->size2 != B->size1 || C->size1 != A->size1 || C->size2 != B->size2) gsl_shim_die("gsl_blas_zgemm: shapes");
    for (size_t i = 0; i < C->size1; ++i)
        for (size_t j = 0; j < C->size2; ++j) {
            gsl_complex acc = (beta.dat[0] == 0.0 && beta.dat[1] == 0.0) ? GSL_COMPLEX_ZERO : gsl_complex_mul(beta, gsl_matrix_complex_get(C, i, j));
            for (size_t k = 0; k < A->size2; ++k) {
                const gsl_complex t = gsl_complex_mul(alpha, gsl_complex_mul(gsl_matrix_complex_get(A, i, k), gsl_matrix_complex_get(B, k, j)));
                acc.dat[0] += t.dat[0]; acc.dat[1] += t.dat[1];
            }
            gsl_matrix_complex_set(C, i, j, acc);
        }
    return GSL_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------ complex LU */
static inline int gsl_linalg_complex_LU_decomp(gsl_matrix_complex *A, gsl_permutation *p, int *signum) {
    typedef std::complex<double> cd;
    const size_t n = A->size1;
    if (A->size2 != n || p->size != n) gsl_shim_die("gsl_linalg_complex_LU_decomp: shapes");
    cd *a = reinterpret_cast<cd *>(A->data);
    const size_t ld = A->tda;
    *signum = 1;
    for (size_t i = 0; i < n; ++i) p->data[i] = i;
    for (size_t j = 0; j + 1 < n; ++j) {
        size_t piv = j; double best = std::abs(a[j * ld + j]);
        for (size_t i = j + 1; i < n; ++i) { const double v = std::abs(a[i * ld + j]); if (v > best) { best = v; piv = i; } }
        if (piv != j) {
            for (size_t k = 0; k < n; ++k) std::swap(a[j * ld + k], a[piv * ld + k]);
            std::swap(p->data[j], p->data[piv]);
            *signum = -*signum;
        }
        const cd ajj = a[j * ld + j];
        if (ajj != cd(0.0, 0.0))
            for (size_t i = j + 1; i < n; ++i) {
                const cd f = a[i * ld + j] / ajj;
                a[i * ld + j] = f;
                for (size_t k = j + 1; k < n; ++k) a[i * ld + k] -= f * a[j * ld + k];
            }
    }
    return GSL_SUCCESS;
}
static inline int gsl_linalg_complex_LU_invert(const gsl_matrix_complex *LU, const gsl_permutation *p, gsl_matrix_complex *inv) {
    typedef std::complex<double> cd;
    const size_t n = LU->size1;
    const cd *a = reinterpret_cast<const cd *>(LU->data);
    cd *x = reinterpret_cast<cd *>(inv->data);
    const size_t ld = LU->tda, lx = inv->tda;
    std::vector<cd> col(n);
    for (size_t c = 0; c < n; ++c) {
        /* solve L U x = P e_c */
        for (size_t i = 0; i < n; ++i) col[i] = (p->data[i] == c) ? cd(1.0, 0.0) : cd(0.0, 0.0);
        for (size_t i = 0; i < n; ++i) { cd s = col[i]; for (size_t k = 0; k < i; ++k) s -= a[i * ld + k] * col[k]; col[i] = s; }
        for (size_t ii = n; ii-- > 0;) { cd s = col[ii]; for (size_t k = ii + 1; k < n; ++k) s -= a[ii * ld + k] * col[k]; col[ii] = s / a[ii * ld + ii]; }
        for (size_t i = 0; i < n; ++i) x[i * lx + c] = col[i];
    }
    return GSL_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------ nonsymmetric eigensystem */
typedef struct { size_t size; } gsl_eigen_nonsymmv_workspace;
typedef enum { GSL_EIGEN_SORT_VAL_ASC, GSL_EIGEN_SORT_VAL_DESC, GSL_EIGEN_SORT_ABS_ASC, GSL_EIGEN_SORT_ABS_DESC } gsl_eigen_sort_t;
static inline gsl_eigen_nonsymmv_workspace *gsl_eigen_nonsymmv_alloc(size_t n) {
    gsl_eigen_nonsymmv_workspace *w = (gsl_eigen_nonsymmv_workspace *)std::malloc(sizeof(gsl_eigen_nonsymmv_workspace));
    w->size = n; return w;
}
static inline void gsl_eigen_nonsymmv_free(gsl_eigen_nonsymmv_workspace *w) { std::free(w); }

namespace gsl_shim_detail {

static inline void cdiv(double xr, double xi, double yr, double yi, double &cr, double &ci) {
    double r, d;
    if (std::fabs(yr) > std::fabs(yi)) { r = yi / yr; d = yr + r * yi; cr = (xr + r * xi) / d; ci = (xi - r * xr) / d; }
    else { r = yr / yi; d = yi + r * yr; cr = (r * xr + xi) / d; ci = (r * xi - xr) / d; }
}

/* Householder reduction of H (n x n, row-major in H[i][j]) to upper Hessenberg form; V accumulates the transformations. */
static inline void hessenberg(int n, std::vector<std::vector<double> > &H, std::vector<std::vector<double> > &V) {
    std::vector<double> ort(n, 0.0);
    const int low = 0, high = n - 1;
    for (int m = low + 1; m <= high - 1; ++m) {
        double scale = 0.0;
        for (int i = m; i <= high; ++i) scale += std::fabs(H[i][m - 1]);
        if (scale != 0.0) {
            double h = 0.0;
            for (int i = high; i >= m; --i) { ort[i] = H[i][m - 1] / scale; h += ort[i] * ort[i]; }
            double g = std::sqrt(h);
            if (ort[m] > 0) g = -g;
            h -= ort[m] * g;
            ort[m] -= g;
            for (int j = m; j < n; ++j) {
                double f = 0.0;
                for (int i = high; i >= m; --i) f += ort[i] * H[i][j];
                f /= h;
                for (int i = m; i <= high; ++i) H[i][j] -= f * ort[i];
            }
            for (int i = 0; i <= high; ++i) {
                double f = 0.0;
                for (int j = high; j >= m; --j) f += ort[j] * H[i][j];
                f /= h;
                for (int j = m; j <= high; ++j) H[i][j] -= f * ort[j];
            }
            ort[m] = scale * ort[m];
            H[m][m - 1] = scale * g;
        }
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int m = high - 1; m >= low + 1; --m) {
        if (H[m][m - 1] != 0.0) {
            for (int i = m + 1; i <= high; ++i) ort[i] = H[i][m - 1];
            for (int j = m; j <= high; ++j) {
                double g = 0.0;
                for (int i = m; i <= high; ++i) g += ort[i] * V[i][j];
                g = (g / ort[m]) / H[m][m - 1];
                for (int i = m; i <= high; ++i) V[i][j] += g * ort[i];
            }
        }
    }
}

/* Francis double-shift QR on the Hessenberg matrix H with accumulation into V, followed by back-substitution:
 * on return d/e hold the eigenvalues and the columns of V the (real-packed) eigenvectors. */
static inline void schur_vectors(int nn, std::vector<std::vector<double> > &H, std::vector<std::vector<double> > &V,
                                 std::vector<double> &d, std::vector<double> &e) {
    int n = nn - 1;
    const int low = 0, high = nn - 1;
    const double eps = std::pow(2.0, -52.0);
    double exshift = 0.0, p = 0, q = 0, r = 0, s = 0, z = 0, t, w, x, y;
    double norm = 0.0;
    for (int i = 0; i < nn; ++i) for (int j = (i - 1 > 0 ? i - 1 : 0); j < nn; ++j) norm += std::fabs(H[i][j]);
    int iter = 0;
    while (n >= low) {
        int l = n;
        while (l > low) {
            s = std::fabs(H[l - 1][l - 1]) + std::fabs(H[l][l]);
            if (s == 0.0) s = norm;
            if (std::fabs(H[l][l - 1]) < eps * s) break;
            --l;
        }
        if (l == n) { /* one root */
            H[n][n] += exshift; d[n] = H[n][n]; e[n] = 0.0; --n; iter = 0;
        } else if (l == n - 1) { /* two roots */
            w = H[n][n - 1] * H[n - 1][n];
            p = (H[n - 1][n - 1] - H[n][n]) / 2.0;
            q = p * p + w;
            /* A conjugate pair whose imaginary part is below n ulps of its real part (the backward error of the QR sweep) is rounding noise around a real
             * double eigenvalue (reversible rate matrices with symmetric structure have many): keep it real, as GSL and
             * LAPACK end up doing through their deflation tests.  The reference drops imaginary parts that small anyway
             * (check_real, src/instance.hpp:21-27) but would then use only the real half of the eigenvector pair. */
            if (q < 0 && std::sqrt(-q) <= nn * eps * (std::fabs(H[n - 1][n - 1] + exshift) + std::fabs(H[n][n] + exshift))) q = 0.0;
            z = std::sqrt(std::fabs(q));
            H[n][n] += exshift; H[n - 1][n - 1] += exshift;
            x = H[n][n];
            if (q >= 0) { /* real pair */
                z = (p >= 0) ? p + z : p - z;
                d[n - 1] = x + z; d[n] = d[n - 1];
                if (z != 0.0) d[n] = x - w / z;
                e[n - 1] = 0.0; e[n] = 0.0;
                x = H[n][n - 1];
                s = std::fabs(x) + std::fabs(z);
                if (s == 0.0) { x = 0.0; z = 1.0; s = 1.0; } /* already diagonal */
                p = x / s; q = z / s;
                r = std::sqrt(p * p + q * q);
                p /= r; q /= r;
                for (int j = n - 1; j < nn; ++j) { z = H[n - 1][j]; H[n - 1][j] = q * z + p * H[n][j]; H[n][j] = q * H[n][j] - p * z; }
                for (int i = 0; i <= n; ++i) { z = H[i][n - 1]; H[i][n - 1] = q * z + p * H[i][n]; H[i][n] = q * H[i][n] - p * z; }
                for (int i = low; i <= high; ++i) { z = V[i][n - 1]; V[i][n - 1] = q * z + p * V[i][n]; V[i][n] = q * V[i][n] - p * z; }
            } else { /* complex pair */
                d[n - 1] = x + p; d[n] = x + p; e[n - 1] = z; e[n] = -z;
            }
            n -= 2; iter = 0;
        } else {
            x = H[n][n]; y = 0.0; w = 0.0;
            if (l < n) { y = H[n - 1][n - 1]; w = H[n][n - 1] * H[n - 1][n]; }
            if (iter == 10) { /* exceptional shift */
                exshift += x;
                for (int i = low; i <= n; ++i) H[i][i] -= x;
                s = std::fabs(H[n][n - 1]) + std::fabs(H[n - 1][n - 2]);
                x = y = 0.75 * s; w = -0.4375 * s * s;
            }
            if (iter == 30) {
                s = (y - x) / 2.0; s = s * s + w;
                if (s > 0) {
                    s = std::sqrt(s);
                    if (y < x) s = -s;
                    s = x - w / ((y - x) / 2.0 + s);
                    for (int i = low; i <= n; ++i) H[i][i] -= s;
                    exshift += s;
                    x = y = w = 0.964;
                }
            }
            ++iter;
            if (iter > 10000) gsl_shim_die("gsl_eigen_nonsymmv: QR iteration did not converge");
            int m = n - 2;
            while (m >= l) {
                z = H[m][m];
                r = x - z; s = y - z;
                p = (r * s - w) / H[m + 1][m] + H[m][m + 1];
                q = H[m + 1][m + 1] - z - r - s;
                r = H[m + 2][m + 1];
                s = std::fabs(p) + std::fabs(q) + std::fabs(r);
                p /= s; q /= s; r /= s;
                if (m == l) break;
                if (std::fabs(H[m][m - 1]) * (std::fabs(q) + std::fabs(r)) <
                    eps * (std::fabs(p) * (std::fabs(H[m - 1][m - 1]) + std::fabs(z) + std::fabs(H[m + 1][m + 1])))) break;
                --m;
            }
            for (int i = m + 2; i <= n; ++i) { H[i][i - 2] = 0.0; if (i > m + 2) H[i][i - 3] = 0.0; }
            for (int k = m; k <= n - 1; ++k) {
                const bool notlast = (k != n - 1);
                if (k != m) {
                    p = H[k][k - 1]; q = H[k + 1][k - 1]; r = notlast ? H[k + 2][k - 1] : 0.0;
                    x = std::fabs(p) + std::fabs(q) + std::fabs(r);
                    if (x == 0.0) continue;
                    p /= x; q /= x; r /= x;
                }
                s = std::sqrt(p * p + q * q + r * r);
                if (p < 0) s = -s;
                if (s != 0) {
                    if (k != m) H[k][k - 1] = -s * x;
                    else if (l != m) H[k][k - 1] = -H[k][k - 1];
                    p += s; x = p / s; y = q / s; z = r / s; q /= p; r /= p;
                    for (int j = k; j < nn; ++j) {
                        p = H[k][j] + q * H[k + 1][j];
                        if (notlast) { p += r * H[k + 2][j]; H[k + 2][j] -= p * z; }
                        H[k][j] -= p * x; H[k + 1][j] -= p * y;
                    }
                    const int lim = (n < k + 3) ? n : k + 3;
                    for (int i = 0; i <= lim; ++i) {
                        p = x * H[i][k] + y * H[i][k + 1];
                        if (notlast) { p += z * H[i][k + 2]; H[i][k + 2] -= p * r; }
                        H[i][k] -= p; H[i][k + 1] -= p * q;
                    }
                    for (int i = low; i <= high; ++i) {
                        p = x * V[i][k] + y * V[i][k + 1];
                        if (notlast) { p += z * V[i][k + 2]; V[i][k + 2] -= p * r; }
                        V[i][k] -= p; V[i][k + 1] -= p * q;
                    }
                }
            }
        }
    }
    if (norm == 0.0) return;
    /* back-substitution: eigenvectors of the quasi-triangular form */
    for (n = nn - 1; n >= 0; --n) {
        p = d[n]; q = e[n];
        if (q == 0) {
            int l = n;
            H[n][n] = 1.0;
            for (int i = n - 1; i >= 0; --i) {
                w = H[i][i] - p;
                r = 0.0;
                for (int j = l; j <= n; ++j) r += H[i][j] * H[j][n];
                if (e[i] < 0.0) { z = w; s = r; }
                else {
                    l = i;
                    if (e[i] == 0.0) {
                        if (w != 0.0) H[i][n] = -r / w; else H[i][n] = -r / (eps * norm);
                    } else {
                        x = H[i][i + 1]; y = H[i + 1][i];
                        q = (d[i] - p) * (d[i] - p) + e[i] * e[i];
                        t = (x * s - z * r) / q;
                        H[i][n] = t;
                        if (std::fabs(x) > std::fabs(z)) H[i + 1][n] = (-r - w * t) / x; else H[i + 1][n] = (-s - y * t) / z;
                    }
                    t = std::fabs(H[i][n]);
                    if ((eps * t) * t > 1) for (int j = i; j <= n; ++j) H[j][n] /= t;
                }
            }
        } else if (q < 0) {
            int l = n - 1;
            if (std::fabs(H[n][n - 1]) > std::fabs(H[n - 1][n])) {
                H[n - 1][n - 1] = q / H[n][n - 1];
                H[n - 1][n] = -(H[n][n] - p) / H[n][n - 1];
            } else {
                double cr, ci;
                cdiv(0.0, -H[n - 1][n], H[n - 1][n - 1] - p, q, cr, ci);
                H[n - 1][n - 1] = cr; H[n - 1][n] = ci;
            }
            H[n][n - 1] = 0.0; H[n][n] = 1.0;
            for (int i = n - 2; i >= 0; --i) {
                double ra = 0.0, sa = 0.0, vr, vi, cr, ci;
                for (int j = l; j <= n; ++j) { ra += H[i][j] * H[j][n - 1]; sa += H[i][j] * H[j][n]; }
                w = H[i][i] - p;
                if (e[i] < 0.0) { z = w; r = ra; s = sa; }
                else {
                    l = i;
                    if (e[i] == 0) {
                        cdiv(-ra, -sa, w, q, cr, ci);
                        H[i][n - 1] = cr; H[i][n] = ci;
                    } else {
                        x = H[i][i + 1]; y = H[i + 1][i];
                        vr = (d[i] - p) * (d[i] - p) + e[i] * e[i] - q * q;
                        vi = (d[i] - p) * 2.0 * q;
                        if (vr == 0.0 && vi == 0.0) vr = eps * norm * (std::fabs(w) + std::fabs(q) + std::fabs(x) + std::fabs(y) + std::fabs(z));
                        cdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, cr, ci);
                        H[i][n - 1] = cr; H[i][n] = ci;
                        if (std::fabs(x) > (std::fabs(z) + std::fabs(q))) {
                            H[i + 1][n - 1] = (-ra - w * H[i][n - 1] + q * H[i][n]) / x;
                            H[i + 1][n] = (-sa - w * H[i][n] - q * H[i][n - 1]) / x;
                        } else {
                            cdiv(-r - y * H[i][n - 1], -s - y * H[i][n], z, q, cr, ci);
                            H[i + 1][n - 1] = cr; H[i + 1][n] = ci;
                        }
                    }
                    t = std::fabs(H[i][n - 1]) > std::fabs(H[i][n]) ? std::fabs(H[i][n - 1]) : std::fabs(H[i][n]);
                    if ((eps * t) * t > 1) for (int j = i; j <= n; ++j) { H[j][n - 1] /= t; H[j][n] /= t; }
                }
            }
        }
    }
    /* back-transformation to the eigenvectors of the original matrix */
    for (int j = nn - 1; j >= low; --j)
        for (int i = low; i <= high; ++i) {
            z = 0.0;
            const int lim = (j < high) ? j : high;
            for (int k = low; k <= lim; ++k) z += V[i][k] * H[k][j];
            V[i][j] = z;
        }
}

} // namespace gsl_shim_detail

/* gsl_eigen_nonsymmv: eigenvalues into eval, unit-norm eigenvectors into the columns of evec; A is destroyed
 * (GSL leaves the Schur form T in it; the reference only reads it back in OMEGA's `cpy == false` mode, where the
 * value is immediately overwritten by the next Q build). */
static inline int gsl_eigen_nonsymmv(gsl_matrix *A, gsl_vector_complex *eval, gsl_matrix_complex *evec, gsl_eigen_nonsymmv_workspace *w) {
    const int n = (int)A->size1;
    if ((size_t)n != A->size2 || eval->size != (size_t)n || evec->size1 != (size_t)n || w->size != (size_t)n) gsl_shim_die("gsl_eigen_nonsymmv: shapes");
    std::vector<std::vector<double> > H(n, std::vector<double>(n)), V(n, std::vector<double>(n));
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) H[i][j] = gsl_matrix_get(A, i, j);
    std::vector<double> d(n, 0.0), e(n, 0.0);
    gsl_shim_detail::hessenberg(n, H, V);
    gsl_shim_detail::schur_vectors(n, H, V, d, e);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) gsl_matrix_set(A, i, j, (j >= i - 1) ? H[i][j] : 0.0);
    for (int j = 0; j < n; ++j) {
        gsl_vector_complex_set(eval, j, gsl_complex_rect(d[j], e[j]));
        if (e[j] == 0.0) {
            double nrm = 0.0;
            for (int i = 0; i < n; ++i) nrm += V[i][j] * V[i][j];
            nrm = std::sqrt(nrm);
            for (int i = 0; i < n; ++i) gsl_matrix_complex_set(evec, i, j, gsl_complex_rect(V[i][j] / nrm, 0.0));
        } else if (e[j] > 0.0) { /* columns j, j+1 hold Re and Im of the pair's first vector */
            double nrm = 0.0;
            for (int i = 0; i < n; ++i) nrm += V[i][j] * V[i][j] + V[i][j + 1] * V[i][j + 1];
            nrm = std::sqrt(nrm);
            for (int i = 0; i < n; ++i) {
                gsl_matrix_complex_set(evec, i, j, gsl_complex_rect(V[i][j] / nrm, V[i][j + 1] / nrm));
                gsl_matrix_complex_set(evec, i, j + 1, gsl_complex_rect(V[i][j] / nrm, -V[i][j + 1] / nrm));
            }
        }
    }
    return GSL_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------ 1-D minimisation (Brent) */
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)

typedef struct { const char *name; } gsl_min_fminimizer_type;
static const gsl_min_fminimizer_type gsl_shim_brent_type = {"brent"};
static const gsl_min_fminimizer_type *const gsl_min_fminimizer_brent = &gsl_shim_brent_type;

typedef struct {
    const gsl_min_fminimizer_type *type;
    gsl_function *function;
    double x_minimum, x_lower, x_upper;
    double f_minimum, f_lower, f_upper;
    /* brent state */
    double d, e, v, w, f_v, f_w;
} gsl_min_fminimizer;

static inline gsl_min_fminimizer *gsl_min_fminimizer_alloc(const gsl_min_fminimizer_type *T) {
    gsl_min_fminimizer *s = (gsl_min_fminimizer *)std::calloc(1, sizeof(gsl_min_fminimizer));
    s->type = T; return s;
}
static inline void gsl_min_fminimizer_free(gsl_min_fminimizer *s) { std::free(s); }
static inline double gsl_min_fminimizer_x_minimum(const gsl_min_fminimizer *s) { return s->x_minimum; }
static inline double gsl_min_fminimizer_x_lower(const gsl_min_fminimizer *s) { return s->x_lower; }
static inline double gsl_min_fminimizer_x_upper(const gsl_min_fminimizer *s) { return s->x_upper; }

static inline int gsl_min_fminimizer_set(gsl_min_fminimizer *s, gsl_function *f, double x_minimum, double x_lower, double x_upper) {
    /* fsolver.c: evaluates f at the guess, the lower and the upper bound, in this order */
    const double f_minimum = GSL_FN_EVAL(f, x_minimum);
    const double f_lower = GSL_FN_EVAL(f, x_lower);
    const double f_upper = GSL_FN_EVAL(f, x_upper);
    s->function = f;
    s->x_minimum = x_minimum; s->x_lower = x_lower; s->x_upper = x_upper;
    if (x_lower > x_upper) gsl_shim_die("gsl_min_fminimizer_set: invalid interval (lower > upper)");
    if (x_minimum >= x_upper || x_minimum <= x_lower) gsl_shim_die("gsl_min_fminimizer_set: x_minimum must lie inside interval");
    s->f_lower = f_lower; s->f_upper = f_upper; s->f_minimum = f_minimum;
    if (f_minimum >= f_lower || f_minimum >= f_upper) gsl_shim_die("gsl_min_fminimizer_set: endpoints do not enclose a minimum");
    /* brent_init */
    const double golden = 0.3819660;
    s->v = x_lower + golden * (x_upper - x_lower);
    s->w = s->v;
    s->d = 0; s->e = 0;
    const double f_vw = GSL_FN_EVAL(f, s->v);
    s->f_v = f_vw; s->f_w = f_vw;
    return GSL_SUCCESS;
}

static inline int gsl_min_fminimizer_iterate(gsl_min_fminimizer *s) {
    gsl_function *f = s->function;
    const double x_left = s->x_lower, x_right = s->x_upper;
    const double z = s->x_minimum;
    double d = s->e, e = s->d; /* sic: loaded swapped */
    double u, f_u;
    const double v = s->v, w = s->w;
    const double f_v = s->f_v, f_w = s->f_w, f_z = s->f_minimum;
    const double golden = 0.3819660;
    const double w_lower = (z - x_left), w_upper = (x_right - z);
    const double tolerance = 1.4901161193847656e-08 * std::fabs(z);
    double p = 0, q = 0, r = 0;
    const double midpoint = 0.5 * (x_left + x_right);

    if (std::fabs(e) > tolerance) {
        r = (z - w) * (f_z - f_v);
        q = (z - v) * (f_z - f_w);
        p = (z - v) * q - (z - w) * r;
        q = 2 * (q - r);
        if (q > 0) p = -p; else q = -q;
        r = e;
        e = d;
    }
    if (std::fabs(p) < std::fabs(0.5 * q * r) && p < q * w_lower && p < q * w_upper) {
        const double t2 = 2 * tolerance;
        d = p / q;
        u = z + d;
        if ((u - x_left) < t2 || (x_right - u) < t2) d = (z < midpoint) ? tolerance : -tolerance;
    } else {
        e = (z < midpoint) ? x_right - z : -(z - x_left);
        d = golden * e;
    }
    if (std::fabs(d) >= tolerance) u = z + d;
    else u = z + ((d > 0) ? tolerance : -tolerance);

    s->e = e; s->d = d;
    f_u = GSL_FN_EVAL(f, u);

    if (f_u <= f_z) {
        if (u < z) { s->x_upper = z; s->f_upper = f_z; }
        else { s->x_lower = z; s->f_lower = f_z; }
        s->v = w; s->f_v = f_w;
        s->w = z; s->f_w = f_z;
        s->x_minimum = u; s->f_minimum = f_u;
        return GSL_SUCCESS;
    } else {
        if (u < z) { s->x_lower = u; s->f_lower = f_u; }
        else { s->x_upper = u; s->f_upper = f_u; }
        if (f_u <= f_w || w == z) {
            s->v = w; s->f_v = f_w;
            s->w = u; s->f_w = f_u;
            return GSL_SUCCESS;
        } else if (f_u <= f_v || v == z || v == w) {
            s->v = u; s->f_v = f_u;
            return GSL_SUCCESS;
        }
    }
    return GSL_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------ gamma density */
static inline double gsl_ran_gamma_pdf(const double x, const double a, const double b) {
    if (x < 0) return 0;
    if (x == 0) return (a == 1) ? 1 / b : 0;
    if (a == 1) return std::exp(-x / b) / b;
    return std::exp((a - 1) * std::log(x / b) - x / b - std::lgamma(a)) / b;
}

#endif /* PCSF_GSL_SHIM_H */
