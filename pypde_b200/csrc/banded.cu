// Banded / recurrence kernels of the Chebyshev-Galerkin hot path (sm_100a).
//
// This translation unit is compiled with --fmad=false: every recurrence keeps
// the operation order of the reference's Fortran (no FMA contraction there
// either), so these kernels are BIT-IDENTICAL to the CPU oracle.  They are
// latency / HBM bound, so the missing FMAs cost nothing.
//
// All sequential operators decouple into an even and an odd index chain
// (offset-2 couplings only), so the unit of work is one (problem, parity) chain:
//   axis 0: one thread per (column, parity); lanes run over consecutive columns,
//           so every global access is a coalesced row segment;
//   axis 1: a CTA stages R full rows in shared memory with coalesced loads
//           (row pitch odd -> conflict-free), 2R threads sweep, coalesced store.
#include "common.cuh"

namespace pde {

struct Acc {            // strided view of one problem
    double *p;
    long s;
    __device__ __forceinline__ double ld(int i) const { return p[(long)i * s]; }
    __device__ __forceinline__ void st(int i, double v) const { p[(long)i * s] = v; }
};

// ---------------------------------------------------------------------------
// chain operators
// ---------------------------------------------------------------------------

// differentiate_cheby.f90:28-53.  dc[n-1] = 0 (f2py zero fill), dc[n-2] = 2(n-1)c[n-1],
// dc[k] = dc[k+2] + 2(k+1)c[k+1] (k = n-3..1), dc[0] = dc[2]/2 + c[1].
// The stored value is dc/div (grad()'s `dvhat /= scale**deriv`), the recurrence
// runs on the undivided value.
struct DiffOp {
    int n;
    double div;
    int use_div;
    static constexpr bool in_place = false;
    __device__ __forceinline__ int n_in() const { return n; }
    __device__ __forceinline__ int n_out() const { return n; }
    __device__ void chain(Acc c, Acc dc, int p) const
    {
        int k = ((n - 1 - p) & 1) ? n - 2 : n - 1;   // largest index of parity p
        if (k < 0) return;
        double cur;
        if (k == n - 1) cur = 0.0;
        else cur = (double)(2 * (n - 1)) * c.ld(n - 1);
        dc.st(k, use_div ? cur / div : cur);
        for (k -= 2; k >= 1; k -= 2) {
            cur = cur + (double)(2 * (k + 1)) * c.ld(k + 1);
            dc.st(k, use_div ? cur / div : cur);
        }
        if (p == 0 && n >= 3) {
            cur = cur / 2.0 + c.ld(1);
            dc.st(0, use_div ? cur / div : cur);
        }
    }
};

// tdma.f90:55-106 with k = 2 and host-precomputed den / w (tdma.f90:82-89).
// Optional fused S^T product in front (chebyshev.py:327): d_k = u_k + s_k u_{k+2}.
struct TdmaOp {
    int n;                 // number of unknowns (M)
    const double *s;       // stencil sub-diagonal (nullptr: plain tdma, input has n entries)
    const double *a, *den, *w;
    static constexpr bool in_place = false;
    __device__ __forceinline__ int n_in() const { return s ? n + 2 : n; }
    __device__ __forceinline__ int n_out() const { return n; }
    __device__ __forceinline__ double rhs(const Acc &u, int i) const
    {
        if (!s) return u.ld(i);
        return u.ld(i) + __ldg(s + i) * u.ld(i + 2);
    }
    __device__ void chain(Acc u, Acc x, int p) const
    {
        if (p >= n) return;
        int i = p;
        double g = rhs(u, i) / __ldg(den + i);
        x.st(i, g);
        for (i += 2; i < n; i += 2) {
            g = (rhs(u, i) - __ldg(a + i - 2) * g) / __ldg(den + i);
            x.st(i, g);
        }
        i -= 2;                     // top of the chain: x = g
        double xv = g;
        for (i -= 2; i >= 0; i -= 2) {
            xv = x.ld(i) - __ldg(w + i) * xv;
            x.st(i, xv);
        }
    }
};

// fdma.f90:26-36 / :68-80
struct FdmaOp {
    int n;
    const double *l, *d, *u1, *u2;
    static constexpr bool in_place = true;
    __device__ __forceinline__ int n_in() const { return n; }
    __device__ __forceinline__ int n_out() const { return n; }
    __device__ void chain(Acc x, Acc, int p) const
    {
        if (p >= n) return;
        int i = p;
        double prev = x.ld(i);
        for (i += 2; i < n; i += 2) {
            prev = x.ld(i) - __ldg(l + i - 2) * prev;
            x.st(i, prev);
        }
        i -= 2;                      // top index of this parity (n-1 or n-2)
        double x2 = prev / __ldg(d + i);
        x.st(i, x2);
        i -= 2;
        if (i < 0) return;
        double x4 = x2;
        x2 = (x.ld(i) - __ldg(u1 + i) * x4) / __ldg(d + i);
        x.st(i, x2);
        for (i -= 2; i >= 0; i -= 2) {
            double v = (x.ld(i) - __ldg(u1 + i) * x2 - __ldg(u2 + i) * x4) / __ldg(d + i);
            x.st(i, v);
            x4 = x2;
            x2 = v;
        }
    }
};

// twodma.f90:17-22 / :45-58
struct TwodmaOp {
    int n;
    const double *d, *u;
    static constexpr bool in_place = true;
    __device__ __forceinline__ int n_in() const { return n; }
    __device__ __forceinline__ int n_out() const { return n; }
    __device__ void chain(Acc x, Acc, int p) const
    {
        int i = ((n - 1 - p) & 1) ? n - 2 : n - 1;
        if (i < 0) return;
        double x2 = x.ld(i) / __ldg(d + i);
        x.st(i, x2);
        for (i -= 2; i >= 0; i -= 2) {
            x2 = (x.ld(i) - __ldg(u + i) * x2) / __ldg(d + i);
            x.st(i, x2);
        }
    }
};

// ---------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------
template <class Op>
__global__ void k_chain_cols(Op op, const double *in, long ldin, double *out, long ldout, int batch)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= batch) return;
    op.chain(Acc{const_cast<double *>(in) + j, ldin}, Acc{out + j, ldout}, (int)threadIdx.y);
}

template <class Op>
__global__ void k_chain_rows(Op op, const double *in, long ldin, double *out, long ldout,
                             int nrows, int R, int W)
{
    extern __shared__ double sm[];
    double *tin = sm;
    double *tout = Op::in_place ? sm : sm + (long)R * W;
    const int r0 = blockIdx.x * R;
    const int rows = min(R, nrows - r0);
    const int nin = op.n_in(), nout = op.n_out();
    for (int r = 0; r < rows; ++r) {
        const double *src = in + (long)(r0 + r) * ldin;
        for (int i = threadIdx.x; i < nin; i += blockDim.x) tin[(long)r * W + i] = src[i];
    }
    __syncthreads();
    if ((int)threadIdx.x < 2 * rows) {
        const int r = threadIdx.x >> 1, p = threadIdx.x & 1;
        op.chain(Acc{tin + (long)r * W, 1}, Acc{tout + (long)r * W, 1}, p);
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
        double *dst = out + (long)(r0 + r) * ldout;
        for (int i = threadIdx.x; i < nout; i += blockDim.x) dst[i] = tout[(long)r * W + i];
    }
}

template <class Op>
static int launch_chain(const Op &op, const double *in, long ldin, double *out, long ldout,
                        int batch, int axis, int nmax, cudaStream_t st, const char *what)
{
    if (batch <= 0 || nmax <= 0) return PDE_OK;
    if (axis == 0) {
        const int tx = batch >= 148 * 128 ? 128 : (batch >= 148 * 64 ? 64 : 32);
        dim3 block(tx, 2);
        k_chain_cols<Op><<<ceil_div(batch, tx), block, 0, st>>>(op, in, ldin, out, ldout, batch);
        return after_launch(what);
    }
    const int W = nmax | 1;
    const int tiles = Op::in_place ? 1 : 2;
    const long budget = 200 * 1024;
    int R = (int)(budget / ((long)tiles * W * 8));
    if (R < 1) {
        set_error("%s: axis-1 problem of length %d does not fit the shared-memory row tile", what, nmax);
        return PDE_ERR_UNSUPPORTED;
    }
    // keep >= ~2 CTAs per SM when there are enough rows
    const int want = ceil_div(batch, 2 * sm_count());
    if (R > 32) R = 32;
    if (R > want) R = want < 1 ? 1 : want;
    const size_t smem = (size_t)tiles * R * W * 8;
    static bool attr_done = false;   // one flag per Op instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_rows<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)budget + 8 * 1024);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return PDE_ERR_CUDA;
        }
        attr_done = true;
    }
    k_chain_rows<Op><<<ceil_div(batch, R), 128, smem, st>>>(op, in, ldin, out, ldout, batch, R, W);
    return after_launch(what);
}

// ---------------------------------------------------------------------------
// pointwise stencils: to_cheb and the banded product (no sequential coupling,
// so both axes use the same coalesced 2-D elementwise kernel)
// ---------------------------------------------------------------------------
__global__ void k_to_cheb(const double *__restrict__ s, const double *__restrict__ v, long ldv, int M,
                          double *__restrict__ u, long ldu, int n0, int n1, int axis)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n0 || j >= n1) return;
    const int k = axis == 0 ? i : j;
    const long step = axis == 0 ? ldv : 1;
    const double *vp = v + (long)i * ldv + j;
    double acc = 0.0;
    if (k >= 2 && k - 2 < M) {
        const double sk = __ldg(s + k - 2);
        if (sk != 0.0) acc = sk * vp[-2 * step];     // CSC product skips stored zeros (tosparse)
    }
    if (k < M) acc = acc + vp[0];
    u[(long)i * ldu + j] = acc;
}

#define MAX_DIAG 8
struct DiagSpec {
    int ndiag;
    int off[MAX_DIAG];
};

__global__ void k_banded_mul(DiagSpec spec, const double *__restrict__ diags, const double *__restrict__ x,
                             long ldx, int n_in, double *__restrict__ y, long ldy, int n_out,
                             int n0, int n1, int axis, int accumulate)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n0 || j >= n1) return;
    const int r = axis == 0 ? i : j;
    double acc = 0.0;
    for (int d = 0; d < spec.ndiag; ++d) {
        const int c = r + spec.off[d];
        if (c < 0 || c >= n_in) continue;
        const double a = __ldg(diags + (long)d * n_out + r);
        if (a == 0.0) continue;                     // CSR stores no explicit zeros
        const double xv = axis == 0 ? x[(long)c * ldx + j] : x[(long)i * ldx + c];
        acc = acc + a * xv;
    }
    double *yp = y + (long)i * ldy + j;
    *yp = accumulate ? *yp + acc : acc;
}

// ---------------------------------------------------------------------------
// Poisson (A + lam_i C) plan: per-column LU (init_fdma, fdma.f90:102-143) kept on
// the device, solve = solve_fdma_1d per column (fdma.f90:26-36, :173-185).
// ---------------------------------------------------------------------------
struct PoissonTables {
    double *l, *d, *u1, *u2;   // (n x m) row-major each
    int *off;                  // 1 where the singular branch drops row/col 0
};

__global__ void k_poisson_factor(const double *__restrict__ Ad, const double *__restrict__ Cd,
                                 const double *__restrict__ lam, PoissonTables t, int n, int m, int singular)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double lm = lam[j];
    // `abs(lam(i)) < 1e-10` with a default-real literal (fdma.f90:176)
    const int off = (singular && fabs(lm) < (double)1e-10f) ? 1 : 0;
    t.off[j] = off;
    const int ne = n - off;
    auto M = [&](int dg, int q) {    // sub-matrix entry (q, q + offset[dg]), offsets -2,0,2,4
        const int r = q + off;
        return Ad[(long)dg * n + r] + Cd[(long)dg * n + r] * lm;
    };
#define T(arr, q) arr[(long)((q) + off) * m + j]
    if (off) { t.l[j] = 0; t.d[j] = 1; t.u1[j] = 0; t.u2[j] = 0; }
    for (int q = 0; q < ne; ++q) {
        T(t.d, q) = M(1, q);
        T(t.l, q) = (q + 2 < ne) ? M(0, q + 2) : 0.0;     // l(i-2) = A(i,i-2)
        T(t.u1, q) = (q + 2 < ne) ? M(2, q) : 0.0;
        T(t.u2, q) = (q + 4 < ne) ? M(3, q) : 0.0;
    }
    for (int q = 2; q < ne; ++q) {
        const double lf = T(t.l, q - 2) / T(t.d, q - 2);
        T(t.l, q - 2) = lf;
        T(t.d, q) = T(t.d, q) - lf * T(t.u1, q - 2);
        if (q < ne - 2) T(t.u1, q) = T(t.u1, q) - lf * T(t.u2, q - 2);
    }
#undef T
}

__global__ void k_poisson_solve(PoissonTables t, double *x, long ldx, int n, int m)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int p = threadIdx.y;
    const int off = t.off[j];
    const int ne = n - off;
    if (off && p == 0) x[j] = 0.0;                 // x(1,i) = 0.0
    if (p >= ne) return;
    double *xp = x + (long)off * ldx + j;
    const long tb = (long)off * m + j;
#define X(q) xp[(long)(q) * ldx]
#define T(arr, q) arr[tb + (long)(q) * m]
    int i = p;
    double prev = X(i);
    for (i += 2; i < ne; i += 2) {
        prev = X(i) - T(t.l, i - 2) * prev;
        X(i) = prev;
    }
    i -= 2;
    double x2 = prev / T(t.d, i);
    X(i) = x2;
    i -= 2;
    if (i < 0) return;
    double x4 = x2;
    x2 = (X(i) - T(t.u1, i) * x4) / T(t.d, i);
    X(i) = x2;
    for (i -= 2; i >= 0; i -= 2) {
        const double v = (X(i) - T(t.u1, i) * x2 - T(t.u2, i) * x4) / T(t.d, i);
        X(i) = v;
        x4 = x2;
        x2 = v;
    }
#undef X
#undef T
}

// tiled transpose, 32x32 tiles, conflict-free
__global__ void k_transpose(const double *__restrict__ in, long ldin, double *__restrict__ out, long ldout,
                            int n0, int n1)
{
    __shared__ double tile[32][33];
    int j = blockIdx.x * 32 + threadIdx.x;
    int i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        if (i0 + r < n0 && j < n1) tile[r][threadIdx.x] = in[(long)(i0 + r) * ldin + j];
    __syncthreads();
    int i = i0 + threadIdx.x;
    int j0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        if (j0 + r < n1 && i < n0) out[(long)(j0 + r) * ldout + i] = tile[threadIdx.x][r];
}

}  // namespace pde

using namespace pde;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int pde_cheb_diff(const double *c, long ldc, double *dc, long lddc, int n, int batch, int axis,
                  int order, double div, void *stream)
{
    PDE_REQUIRE(c && dc && c != dc, "c, dc must be distinct device pointers");
    PDE_REQUIRE(n >= 3 && batch >= 0, "n >= 3");
    PDE_REQUIRE(order >= 1 && order <= 2, "order must be 1 or 2");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    cudaStream_t st = as_stream(stream);
    if (order == 1) {
        DiffOp op{n, div, div != 1.0};
        return launch_chain(op, c, ldc, dc, lddc, batch, axis, n, st, "pde_cheb_diff");
    }
    // order 2: c -> tmp -> dc
    double *tmp = nullptr;
    const long n0 = axis == 0 ? n : batch, n1 = axis == 0 ? batch : n;
    PDE_CUDA(cudaMallocAsync(&tmp, sizeof(double) * n0 * n1, st));
    DiffOp op1{n, 1.0, 0};
    int rc = launch_chain(op1, c, ldc, tmp, n1, batch, axis, n, st, "pde_cheb_diff(1/2)");
    if (rc == PDE_OK) {
        DiffOp op2{n, div, div != 1.0};
        rc = launch_chain(op2, tmp, n1, dc, lddc, batch, axis, n, st, "pde_cheb_diff(2/2)");
    }
    cudaFreeAsync(tmp, st);
    return rc;
}

int pde_tdma2_solve(const double *a, const double *den, const double *w, const double *d, long ldd,
                    int n, double *x, long ldx, int batch, int axis, void *stream)
{
    PDE_REQUIRE(a && den && w && d && x, "null pointer");
    PDE_REQUIRE(n >= 3, "n >= 3");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    TdmaOp op{n, nullptr, a, den, w};
    return launch_chain(op, d, ldd, x, ldx, batch, axis, n, as_stream(stream), "pde_tdma2_solve");
}

int pde_from_cheb(const double *s, const double *a, const double *den, const double *w,
                  const double *u, long ldu, int M, double *v, long ldv, int batch, int axis, void *stream)
{
    PDE_REQUIRE(s && a && den && w && u && v, "null pointer");
    PDE_REQUIRE(M >= 3, "M >= 3");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    TdmaOp op{M, s, a, den, w};
    return launch_chain(op, u, ldu, v, ldv, batch, axis, M + 2, as_stream(stream), "pde_from_cheb");
}

int pde_fdma_solve(const double *l, const double *d, const double *u1, const double *u2, double *x,
                   long ldx, int n, int batch, int axis, void *stream)
{
    PDE_REQUIRE(l && d && u1 && u2 && x, "null pointer");
    PDE_REQUIRE(n >= 5, "n >= 5");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    FdmaOp op{n, l, d, u1, u2};
    return launch_chain(op, x, ldx, x, ldx, batch, axis, n, as_stream(stream), "pde_fdma_solve");
}

int pde_twodma_solve(const double *d, const double *u, double *x, long ldx, int n, int batch, int axis,
                     void *stream)
{
    PDE_REQUIRE(d && u && x, "null pointer");
    PDE_REQUIRE(n >= 3, "n >= 3");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    TwodmaOp op{n, d, u};
    return launch_chain(op, x, ldx, x, ldx, batch, axis, n, as_stream(stream), "pde_twodma_solve");
}

int pde_to_cheb(const double *s, const double *v, long ldv, int M, double *u, long ldu, int n_out,
                int batch, int axis, void *stream)
{
    PDE_REQUIRE(s && v && u && v != u, "null/aliased pointer");
    PDE_REQUIRE(M >= 1 && n_out >= 1, "sizes");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    if (batch <= 0) return PDE_OK;
    const int n0 = axis == 0 ? n_out : batch, n1 = axis == 0 ? batch : n_out;
    dim3 block(64, 4), grid(ceil_div(n1, 64), ceil_div(n0, 4));
    k_to_cheb<<<grid, block, 0, as_stream(stream)>>>(s, v, ldv, M, u, ldu, n0, n1, axis);
    return after_launch("pde_to_cheb");
}

int pde_banded_mul(const double *diags, const int *offsets, int ndiag, const double *x, long ldx, int n_in,
                   double *y, long ldy, int n_out, int batch, int axis, int accumulate, void *stream)
{
    PDE_REQUIRE(diags && offsets && x && y && x != y, "null/aliased pointer");
    PDE_REQUIRE(ndiag >= 1 && ndiag <= MAX_DIAG, "1..8 diagonals");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    if (batch <= 0) return PDE_OK;
    DiagSpec spec;
    spec.ndiag = ndiag;
    for (int d = 0; d < ndiag; ++d) {
        spec.off[d] = offsets[d];
        if (d) PDE_REQUIRE(offsets[d] > offsets[d - 1], "offsets must ascend");
    }
    const int n0 = axis == 0 ? n_out : batch, n1 = axis == 0 ? batch : n_out;
    dim3 block(64, 4), grid(ceil_div(n1, 64), ceil_div(n0, 4));
    k_banded_mul<<<grid, block, 0, as_stream(stream)>>>(spec, diags, x, ldx, n_in, y, ldy, n_out, n0, n1,
                                                         axis, accumulate);
    return after_launch("pde_banded_mul");
}

struct pde_poisson_plan_s {
    PoissonTables t;
    int n, m;
};

int pde_poisson_plan_create(pde_poisson_plan_t *plan, const double *Adiag, const double *Cdiag,
                            const double *lam, int n, int m, int singular)
{
    PDE_REQUIRE(plan && Adiag && Cdiag && lam, "null pointer");
    PDE_REQUIRE(n >= 6 && m >= 1, "n >= 6");
    pde_poisson_plan_s *p = new pde_poisson_plan_s();
    p->n = n;
    p->m = m;
    const size_t tb = sizeof(double) * (size_t)n * m;
    double *dA = nullptr, *dC = nullptr, *dl = nullptr;
    PDE_CUDA(cudaMalloc(&p->t.l, tb));
    PDE_CUDA(cudaMalloc(&p->t.d, tb));
    PDE_CUDA(cudaMalloc(&p->t.u1, tb));
    PDE_CUDA(cudaMalloc(&p->t.u2, tb));
    PDE_CUDA(cudaMalloc(&p->t.off, sizeof(int) * m));
    PDE_CUDA(cudaMalloc(&dA, sizeof(double) * 4 * n));
    PDE_CUDA(cudaMalloc(&dC, sizeof(double) * 4 * n));
    PDE_CUDA(cudaMalloc(&dl, sizeof(double) * m));
    PDE_CUDA(cudaMemcpy(dA, Adiag, sizeof(double) * 4 * n, cudaMemcpyHostToDevice));
    PDE_CUDA(cudaMemcpy(dC, Cdiag, sizeof(double) * 4 * n, cudaMemcpyHostToDevice));
    PDE_CUDA(cudaMemcpy(dl, lam, sizeof(double) * m, cudaMemcpyHostToDevice));
    k_poisson_factor<<<ceil_div(m, 64), 64>>>(dA, dC, dl, p->t, n, m, singular);
    int rc = after_launch("pde_poisson_plan_create");
    PDE_CUDA(cudaDeviceSynchronize());
    cudaFree(dA);
    cudaFree(dC);
    cudaFree(dl);
    if (rc != PDE_OK) return rc;
    *plan = p;
    return PDE_OK;
}

int pde_poisson_plan_destroy(pde_poisson_plan_t p)
{
    if (!p) return PDE_OK;
    cudaFree(p->t.l);
    cudaFree(p->t.d);
    cudaFree(p->t.u1);
    cudaFree(p->t.u2);
    cudaFree(p->t.off);
    delete p;
    return PDE_OK;
}

int pde_poisson_solve(pde_poisson_plan_t p, double *x, long ldx, void *stream)
{
    PDE_REQUIRE(p && x, "null pointer");
    const int tx = 32;
    dim3 block(tx, 2);
    k_poisson_solve<<<ceil_div(p->m, tx), block, 0, as_stream(stream)>>>(p->t, x, ldx, p->n, p->m);
    return after_launch("pde_poisson_solve");
}

int pde_transpose(const double *in, long ldin, double *out, long ldout, int n0, int n1, void *stream)
{
    PDE_REQUIRE(in && out && in != out, "null/aliased pointer");
    if (n0 <= 0 || n1 <= 0) return PDE_OK;
    dim3 block(32, 8), grid(ceil_div(n1, 32), ceil_div(n0, 32));
    k_transpose<<<grid, block, 0, as_stream(stream)>>>(in, ldin, out, ldout, n0, n1);
    return after_launch("pde_transpose");
}

}  // extern "C"
