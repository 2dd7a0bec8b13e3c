/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
 *
 * Plain-C restatement of the four f2py Fortran modules on pypde's hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product path (pypde_b200) never does.
 *
 * Arrays are addressed with explicit element strides (s0 = stride of axis 0,
 * s1 = stride of axis 1) so that both C-ordered NumPy arrays and the
 * F-ordered views the reference produces with swapaxes are handled without
 * copies.  All loops keep the floating-point operation order of the Fortran
 * sources so results are bit-identical to a gfortran build without
 * -ffast-math (verified against the shipped .so symbols where callable, see
 * tests/golden/make_golden.py).
 *
 * Follows:
 *   pypde/bases/fortran/differentiate_cheby.f90:1-53      (diff_1d / diff_2d)
 *   pypde/bases/linalg/fortran/tdma.f90:1-106             (solve_tdma_1d / _2d)
 *   pypde/solver/linalg/fortran/fdma.f90:1-195            (solve_fdma_1d / _2d, init_fdma, solve_fdma_type2)
 *   pypde/solver/linalg/fortran/twodma.f90:1-61           (solve_twodma_1d / _2d)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AT(p, i, j) (p)[(long)(i) * s0 + (long)(j) * s1]

/* ---- differentiate_cheby.f90:28-53 (diff_2d; diff_1d is m == 1) ----------
 * dc is intent(out): f2py hands the routine a zero-filled array, and row n
 * (1-based) is never written, so dc(n,:) == 0 is part of the result.
 * c, dc: (n, m), derivative along axis 0. Input strides (cs0, cs1), output C-ordered. */
void orc_diff_2d(const double *c, long cs0, long cs1, double *dc, int n, int m)
{
    for (long t = 0; t < (long)n * m; ++t) dc[t] = 0.0;
    if (n < 2) return;
    /* dc(n-1,:) = 2*(n-1)*c(n,:) */
    for (int j = 0; j < m; ++j)
        dc[(long)(n - 2) * m + j] = (double)(2 * (n - 1)) * c[(long)(n - 1) * cs0 + (long)j * cs1];
    /* do i=n-2,2,-1: dc(i,:) = dc(i+2,:) + 2*i*c(i+1,:)  (1-based i) */
    for (int i = n - 2; i >= 2; --i) {
        const double f = (double)(2 * i);
        for (int j = 0; j < m; ++j)
            dc[(long)(i - 1) * m + j] = dc[(long)(i + 1) * m + j] + f * c[(long)i * cs0 + (long)j * cs1];
    }
    /* dc(1,:) = dc(3,:)/2. + c(2,:) */
    if (n >= 3) {
        for (int j = 0; j < m; ++j)
            dc[j] = dc[(long)2 * m + j] / 2.0 + c[cs0 + (long)j * cs1];
    }
}

/* ---- tdma.f90:55-106 (solve_tdma_2d; 1d is m == 1) ------------------------
 * Solve A x = d, A banded at offsets -k, 0, +k; a: n-k, b: n, c: n-k.
 * d: (n, m) with strides (ds0, ds1); x: (n, m) C-ordered output.
 * w and the denominators are recomputed on every call, as in the Fortran. */
void orc_solve_tdma_2d(const double *a, const double *b, const double *c,
                       const double *d, long ds0, long ds1, int k, double *x, int n, int m)
{
    double *w = (double *)malloc(sizeof(double) * (size_t)(n > k ? n - k : 1));
    double *g = (double *)malloc(sizeof(double) * (size_t)n * (size_t)m);
    for (int i = 1; i <= n - k; ++i) {
        if (i < k + 1) w[i - 1] = c[i - 1] / b[i - 1];
        else w[i - 1] = c[i - 1] / (b[i - 1] - a[i - k - 1] * w[i - k - 1]);
    }
    for (int i = 1; i <= n; ++i) {
        if (i < k + 1) {
            for (int j = 0; j < m; ++j)
                g[(long)(i - 1) * m + j] = d[(long)(i - 1) * ds0 + (long)j * ds1] / b[i - 1];
        } else {
            const double den = b[i - 1] - a[i - k - 1] * w[i - k - 1];
            const double al = a[i - k - 1];
            for (int j = 0; j < m; ++j)
                g[(long)(i - 1) * m + j] =
                    (d[(long)(i - 1) * ds0 + (long)j * ds1] - al * g[(long)(i - k - 1) * m + j]) / den;
        }
    }
    /* x(n-k:n,:) = g(n-k:n,:) */
    for (int i = n - k; i <= n; ++i)
        if (i >= 1)
            for (int j = 0; j < m; ++j) x[(long)(i - 1) * m + j] = g[(long)(i - 1) * m + j];
    /* do i=n-k+1,2,-1: x(i-1,:) = g(i-1,:) - w(i-1)*x(i+k-1,:) */
    for (int i = n - k + 1; i >= 2; --i) {
        const double wi = w[i - 2];
        for (int j = 0; j < m; ++j)
            x[(long)(i - 2) * m + j] = g[(long)(i - 2) * m + j] - wi * x[(long)(i + k - 2) * m + j];
    }
    free(w);
    free(g);
}

/* ---- fdma.f90:1-38 (solve_fdma_1d) on a strided vector -------------------- */
static void fdma_1d(const double *l, const double *d, const double *u1, const double *u2,
                    double *x, long s, int n)
{
#define X(i) x[(long)((i) - 1) * s]
    for (int i = 3; i <= n; ++i) X(i) = X(i) - l[i - 3] * X(i - 2);
    X(n) = X(n) / d[n - 1];
    X(n - 1) = X(n - 1) / d[n - 2];
    X(n - 2) = (X(n - 2) - u1[n - 3] * X(n)) / d[n - 3];
    X(n - 3) = (X(n - 3) - u1[n - 4] * X(n - 1)) / d[n - 4];
    for (int i = n - 4; i >= 1; --i)
        X(i) = (X(i) - u1[i - 1] * X(i + 2) - u2[i - 1] * X(i + 4)) / d[i - 1];
#undef X
}

void orc_solve_fdma_1d(const double *l, const double *d, const double *u1, const double *u2,
                       double *x, long s, int n)
{
    fdma_1d(l, d, u1, u2, x, s, n);
}

/* ---- fdma.f90:40-98 (solve_fdma_2d), in place on x(n,m), strides (s0,s1) --
 * The Fortran vectorises over the non-solve axis; per element the operation
 * order equals the 1-D sweep, so looping vector-by-vector is bit-identical. */
void orc_solve_fdma_2d(const double *l, const double *d, const double *u1, const double *u2,
                       double *x, long s0, long s1, int axis, int n, int m)
{
    if (axis == 0) {
        for (int j = 0; j < m; ++j) fdma_1d(l, d, u1, u2, x + (long)j * s1, s0, n);
    } else {
        for (int i = 0; i < n; ++i) fdma_1d(l, d, u1, u2, x + (long)i * s0, s1, m);
    }
}

/* ---- fdma.f90:102-143 (init_fdma) from the four diagonals of M = A + lam*C --
 * The Fortran forms the dense temporary A + C*lam(i) and init_fdma reads only
 * its diagonals at offsets -2, 0, 2, 4 (fdma.f90:128-133); evaluating
 * A(i,j) + C(i,j)*lam on those entries only is the same floating-point
 * expression.  A, C are dense row-major (n0 x n0); `off` drops the leading
 * rows/cols (the singular branch passes A(2:,2:), fdma.f90:177). */
static void init_fdma_sum(const double *A, const double *C, int n0, double lam, int off, int n,
                          double *d, double *u1, double *u2, double *l)
{
#define MIJ(i, j) (A[(long)((i) - 1 + off) * n0 + ((j) - 1 + off)] + C[(long)((i) - 1 + off) * n0 + ((j) - 1 + off)] * lam)
    for (int i = 0; i < n; ++i) d[i] = 0.0;
    for (int i = 0; i < n - 2; ++i) u1[i] = 0.0, l[i] = 0.0;
    for (int i = 0; i < n - 4; ++i) u2[i] = 0.0;
    for (int i = 1; i <= n; ++i) {
        d[i - 1] = MIJ(i, i);
        if (i > 2) l[i - 3] = MIJ(i, i - 2);
        if (i < n - 1) u1[i - 1] = MIJ(i, i + 2);
        if (i < n - 3) u2[i - 1] = MIJ(i, i + 4);
    }
    for (int i = 3; i <= n; ++i) {
        l[i - 3] = l[i - 3] / d[i - 3];
        d[i - 1] = d[i - 1] - l[i - 3] * u1[i - 3];
        if (i < n - 1) u1[i - 1] = u1[i - 1] - l[i - 3] * u2[i - 3];
    }
#undef MIJ
}

/* ---- fdma.f90:146-195 (solve_fdma_type2) ------------------------------------
 * (A + lam_i C) x_i = b_i, in place on x (n, m) with strides (s0, s1).
 * axis 0: one system per column i (A, C are n x n, lam has m entries).
 * axis 1: one system per row i    (A, C are m x m, lam has n entries); the
 *         Fortran has no singular branch there (fdma.f90:187-193). */
void orc_solve_fdma_type2(const double *A, const double *C, const double *lam,
                          double *x, long s0, long s1, int axis, int singular, int n, int m)
{
    const int len = axis == 0 ? n : m;
    double *d = (double *)malloc(sizeof(double) * (size_t)(4 * len + 8));
    double *u1 = d + len + 2, *u2 = u1 + len + 2, *l = u2 + len + 2;
    if (axis == 0) {
        for (int i = 0; i < m; ++i) {
            /* 1e-10 is a default-real literal in the Fortran (fdma.f90:176) */
            if (singular && fabs(lam[i]) < (double)1e-10f) {
                init_fdma_sum(A, C, n, lam[i], 1, n - 1, d, u1, u2, l);
                fdma_1d(l, d, u1, u2, x + s0 + (long)i * s1, s0, n - 1);
                x[(long)i * s1] = 0.0;
            } else {
                init_fdma_sum(A, C, n, lam[i], 0, n, d, u1, u2, l);
                fdma_1d(l, d, u1, u2, x + (long)i * s1, s0, n);
            }
        }
    } else {
        for (int i = 0; i < n; ++i) {
            init_fdma_sum(A, C, m, lam[i], 0, m, d, u1, u2, l);
            fdma_1d(l, d, u1, u2, x + (long)i * s0, s1, m);
        }
    }
    free(d);
}

/* ---- twodma.f90:1-25 / 27-61 ------------------------------------------------ */
static void twodma_1d(const double *d, const double *u, double *x, long s, int n)
{
#define X(i) x[(long)((i) - 1) * s]
    X(n) = X(n) / d[n - 1];
    X(n - 1) = X(n - 1) / d[n - 2];
    for (int i = n - 2; i >= 1; --i) X(i) = (X(i) - u[i - 1] * X(i + 2)) / d[i - 1];
#undef X
}

void orc_solve_twodma_1d(const double *d, const double *u, double *x, long s, int n)
{
    twodma_1d(d, u, x, s, n);
}

void orc_solve_twodma_2d(const double *d, const double *u, double *x, long s0, long s1,
                         int axis, int n, int m)
{
    if (axis == 0) {
        for (int j = 0; j < m; ++j) twodma_1d(d, u, x + (long)j * s1, s0, n);
    } else {
        for (int i = 0; i < n; ++i) twodma_1d(d, u, x + (long)i * s0, s1, m);
    }
}
