// Banded / recurrence kernels of the Chebyshev-Galerkin hot path (sm_100a).
//
// This translation unit is compiled with --fmad=false: every recurrence keeps
// the operation order of the reference's Fortran (no FMA contraction there
// either), so these kernels are BIT-IDENTICAL to the CPU oracle.  They are
// latency / HBM bound, so the missing FMAs cost nothing.
//
// The sequential operators run as sequence-per-thread sweeps with a cp.async prefetch
// ring (sweeps.cuh); this file holds the pointwise stencils, the Poisson LU set-up,
// the transpose and the C-ABI entry points.
#include "common.cuh"
#include "sweeps.cuh"

namespace pde {

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
    double *rd;                // RN(1/d), for the correctly rounded fast division
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
    for (int q = 0; q < ne; ++q) T(t.rd, q) = 1.0 / T(t.d, q);
    if (off) t.rd[j] = 1.0;
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

static SweepJob make_job(const double *in, long ldin, double *out, long ldout, int nseq)
{
    SweepJob j{};
    j.in[0] = in;
    j.ldin[0] = ldin;
    j.out = out;
    j.ldout = ldout;
    j.nseq = nseq;
    return j;
}

int pde_cheb_diff(const double *c, long ldc, double *dc, long lddc, int n, int batch, int axis,
                  int order, double div, void *stream)
{
    PDE_REQUIRE(c && dc && c != dc, "c, dc must be distinct device pointers");
    PDE_REQUIRE(n >= 3 && batch >= 0, "n >= 3");
    PDE_REQUIRE(order >= 1 && order <= 2, "order must be 1 or 2");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    cudaStream_t st = as_stream(stream);
    SweepJobs jobs{};
    jobs.njobs = 1;
    jobs.n = n;
    if (order == 1) {
        jobs.j[0] = make_job(c, ldc, dc, lddc, batch);
        jobs.j[0].flag = div != 1.0;
        jobs.j[0].sc = div;
        return launch_sweep<DiffDesc<false>>(jobs, axis, st, "pde_cheb_diff");
    }
    double *tmp = nullptr;
    const long n0 = axis == 0 ? n : batch, n1 = axis == 0 ? batch : n;
    PDE_CUDA(cudaMallocAsync(&tmp, sizeof(double) * n0 * n1, st));
    jobs.j[0] = make_job(c, ldc, tmp, n1, batch);
    int rc = launch_sweep<DiffDesc<false>>(jobs, axis, st, "pde_cheb_diff(1/2)");
    if (rc == PDE_OK) {
        jobs.j[0] = make_job(tmp, n1, dc, lddc, batch);
        jobs.j[0].flag = div != 1.0;
        jobs.j[0].sc = div;
        rc = launch_sweep<DiffDesc<false>>(jobs, axis, st, "pde_cheb_diff(2/2)");
    }
    cudaFreeAsync(tmp, st);
    return rc;
}

static int tdma_run(const double *s, const double *a, const double *den, const double *w, const double *u,
                    long ldu, int n, double *x, long ldx, int batch, int axis, cudaStream_t st, const char *what)
{
    SweepJobs jobs{};
    jobs.njobs = 1;
    jobs.n = n;
    jobs.j[0] = make_job(u, ldu, x, ldx, batch);
    jobs.j[0].in[1] = s ? u : nullptr;
    jobs.j[0].ldin[1] = ldu;
    jobs.j[0].tab[0] = s;
    jobs.j[0].tab[1] = a;
    jobs.j[0].tab[2] = den;
    jobs.j[0].tab[3] = w;
    int rc = launch_sweep<TdmaFwd<false>>(jobs, axis, st, what);
    if (rc != PDE_OK) return rc;
    jobs.j[0].in[0] = x;
    jobs.j[0].ldin[0] = ldx;
    jobs.j[0].in[1] = nullptr;
    return launch_sweep<TdmaBwd<false>>(jobs, axis, st, what);
}

int pde_tdma2_solve(const double *a, const double *den, const double *w, const double *d, long ldd,
                    int n, double *x, long ldx, int batch, int axis, void *stream)
{
    PDE_REQUIRE(a && den && w && d && x && d != x, "null/aliased pointer");
    PDE_REQUIRE(n >= 3, "n >= 3");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    return tdma_run(nullptr, a, den, w, d, ldd, n, x, ldx, batch, axis, as_stream(stream), "pde_tdma2_solve");
}

int pde_from_cheb(const double *s, const double *a, const double *den, const double *w,
                  const double *u, long ldu, int M, double *v, long ldv, int batch, int axis, void *stream)
{
    PDE_REQUIRE(s && a && den && w && u && v && u != v, "null/aliased pointer");
    PDE_REQUIRE(M >= 3, "M >= 3");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    return tdma_run(s, a, den, w, u, ldu, M, v, ldv, batch, axis, as_stream(stream), "pde_from_cheb");
}

int pde_fdma_solve(const double *l, const double *d, const double *u1, const double *u2, double *x,
                   long ldx, int n, int batch, int axis, void *stream)
{
    PDE_REQUIRE(l && d && u1 && u2 && x, "null pointer");
    PDE_REQUIRE(n >= 5, "n >= 5");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    SweepJobs jobs{};
    jobs.njobs = 1;
    jobs.n = n;
    jobs.j[0] = make_job(x, ldx, x, ldx, batch);
    jobs.j[0].tab[0] = l;
    jobs.j[0].tab[1] = d;
    jobs.j[0].tab[2] = u1;
    jobs.j[0].tab[3] = u2;
    int rc = launch_sweep<FdmaFwd<false>>(jobs, axis, as_stream(stream), "pde_fdma_solve(fwd)");
    if (rc != PDE_OK) return rc;
    return launch_sweep<FdmaBwd<false>>(jobs, axis, as_stream(stream), "pde_fdma_solve(bwd)");
}

int pde_twodma_solve(const double *d, const double *u, double *x, long ldx, int n, int batch, int axis,
                     void *stream)
{
    PDE_REQUIRE(d && u && x, "null pointer");
    PDE_REQUIRE(n >= 3, "n >= 3");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    SweepJobs jobs{};
    jobs.njobs = 1;
    jobs.n = n;
    jobs.j[0] = make_job(x, ldx, x, ldx, batch);
    jobs.j[0].tab[0] = d;
    jobs.j[0].tab[1] = u;
    return launch_sweep<TwodmaBwd<false>>(jobs, axis, as_stream(stream), "pde_twodma_solve");
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
    PDE_CUDA(cudaMalloc(&p->t.rd, tb));
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
    cudaFree(p->t.rd);
    cudaFree(p->t.off);
    delete p;
    return PDE_OK;
}

int pde_poisson_plan_export(pde_poisson_plan_t p, int which, void *dst)
{
    PDE_REQUIRE(p && dst && which >= 0 && which <= 5, "arguments");
    const double *src[5] = {p->t.l, p->t.d, p->t.u1, p->t.u2, p->t.rd};
    if (which < 5) PDE_CUDA(cudaMemcpy(dst, src[which], sizeof(double) * (size_t)p->n * p->m, cudaMemcpyDeviceToDevice));
    else PDE_CUDA(cudaMemcpy(dst, p->t.off, sizeof(int) * (size_t)p->m, cudaMemcpyDeviceToDevice));
    return PDE_OK;
}

int pde_poisson_solve(pde_poisson_plan_t p, double *x, long ldx, void *stream)
{
    PDE_REQUIRE(p && x, "null pointer");
    SweepJobs jobs{};
    jobs.njobs = 1;
    jobs.n = p->n;
    jobs.j[0] = make_job(x, ldx, x, ldx, p->m);
    jobs.j[0].itab = p->t.off;
    jobs.j[0].in[1] = p->t.l;
    jobs.j[0].ldin[1] = p->m;
    int rc = launch_sweep<PoissonFwd<true>>(jobs, 0, as_stream(stream), "pde_poisson_solve(fwd)");
    if (rc != PDE_OK) return rc;
    jobs.j[0].in[1] = p->t.d;
    jobs.j[0].in[2] = p->t.u1;
    jobs.j[0].in[3] = p->t.u2;
    jobs.j[0].in[4] = p->t.rd;
    jobs.j[0].ldin[1] = jobs.j[0].ldin[2] = jobs.j[0].ldin[3] = jobs.j[0].ldin[4] = p->m;
    return launch_sweep<PoissonBwd<true>>(jobs, 0, as_stream(stream), "pde_poisson_solve(bwd)");
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
