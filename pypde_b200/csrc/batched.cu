// Batched (multi-array) entry points of the C ABI: pointwise stencil / banded / linear-
// combination kernels, the pseudo-spectral products, and the generic sweep dispatcher.
// Compiled with --fmad=false: every product and sum is rounded separately, like the
// NumPy / SciPy expressions of the reference.
#include "common.cuh"
#include <cmath>
#include "sweeps.cuh"

namespace pde {

struct StencilJobs {
    int njobs;
    pde_stencil_job j[PDE_MAX_JOBS];
};

// The three pointwise kernels below compute RPT rows per thread and issue EVERY load of those rows before
// the first use.  The first versions tested the loaded coefficient (skip zero diagonals entries) before
// loading the operand, one tap after the other: four serialised memory round trips per output, 1.8 TB/s.
// The selects keep the arithmetic of the reference expression exactly (same products, same order of sums,
// zero coefficients skipped).
constexpr int RPT = 4;

__global__ void k_to_cheb_multi(StencilJobs jobs, int axis)
{
    const pde_stencil_job &jb = jobs.j[blockIdx.z];
    const int n0 = axis == 0 ? jb.n_out : jb.batch, n1 = axis == 0 ? jb.batch : jb.n_out;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * (blockDim.y * RPT) + threadIdx.y;
    if (j >= n1) return;
    const long step = axis == 0 ? jb.ldv : 1;
    double sk[RPT], vm[RPT], v0[RPT];
    bool lo[RPT], hi[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = i0 + r * blockDim.y;
        const int k = axis == 0 ? i : j;
        const double *vp = jb.v + (long)i * jb.ldv + j;
        lo[r] = i < n0 && k >= 2 && k - 2 < jb.M;
        hi[r] = i < n0 && k < jb.M;
        sk[r] = lo[r] ? __ldg(jb.s + k - 2) : 0.0;
        vm[r] = lo[r] ? vp[-2 * step] : 0.0;
        v0[r] = hi[r] ? vp[0] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = i0 + r * blockDim.y;
        if (i >= n0) continue;
        double acc = (lo[r] && sk[r] != 0.0) ? sk[r] * vm[r] : 0.0;
        if (hi[r]) acc = acc + v0[r];
        jb.u[(long)i * jb.ldu + j] = acc;
    }
}

struct BandJobs {
    int njobs;
    pde_band_job j[PDE_MAX_JOBS];
};

constexpr int RPB = 2;       // rows per thread of the banded product (4 needs 93 registers and was slower)

template <int MAXD>
__global__ void __launch_bounds__(256, 4) k_banded_multi(BandJobs jobs, int axis)
{
    const pde_band_job &jb = jobs.j[blockIdx.z];
    const int n0 = axis == 0 ? jb.n_out : jb.batch, n1 = axis == 0 ? jb.batch : jb.n_out;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * (blockDim.y * RPB) + threadIdx.y;
    if (j >= n1) return;
    double a[RPB][MAXD], xv[RPB][MAXD], yo[RPB];
    bool ok[RPB][MAXD];
#pragma unroll
    for (int r = 0; r < RPB; ++r) {
        const int i = i0 + r * blockDim.y;
        const int row = axis == 0 ? i : j;
#pragma unroll
        for (int d = 0; d < MAXD; ++d) {
            const int c = row + (d < jb.ndiag ? jb.off[d] : 0);
            ok[r][d] = i < n0 && d < jb.ndiag && c >= 0 && c < jb.n_in;
            a[r][d] = ok[r][d] ? __ldg(jb.diags + (long)d * jb.n_out + row) : 0.0;
            xv[r][d] = ok[r][d] ? (axis == 0 ? jb.x[(long)c * jb.ldx + j] : jb.x[(long)i * jb.ldx + c]) : 0.0;
        }
        yo[r] = (jb.accumulate && i < n0) ? jb.y[(long)i * jb.ldy + j] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < RPB; ++r) {
        const int i = i0 + r * blockDim.y;
        if (i >= n0) continue;
        double acc = 0.0;
#pragma unroll
        for (int d = 0; d < MAXD; ++d)
            if (ok[r][d] && a[r][d] != 0.0) acc = acc + a[r][d] * xv[r][d];
        jb.y[(long)i * jb.ldy + j] = jb.accumulate ? yo[r] + acc : acc;
    }
}

struct LincombJobs {
    int njobs;
    pde_lincomb_job j[PDE_MAX_JOBS];
};

__global__ void k_lincomb_multi(LincombJobs jobs)
{
    const pde_lincomb_job &jb = jobs.j[blockIdx.z];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * (blockDim.y * RPT) + threadIdx.y;
    if (j >= jb.n1) return;
    constexpr int MAXT = 4;
    double xv[RPT][MAXT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = i0 + r * blockDim.y;
#pragma unroll
        for (int t = 0; t < MAXT; ++t)
            xv[r][t] = (i < jb.n0 && t < jb.nterm) ? jb.x[t][(long)i * jb.ldx[t] + j] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = i0 + r * blockDim.y;
        if (i >= jb.n0) continue;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            if (t >= jb.nterm) break;
            const double term = jb.coef[t] == 1.0 ? xv[r][t] : jb.coef[t] * xv[r][t];
            acc = t == 0 ? term : acc + term;
        }
        jb.y[(long)i * jb.ldy + j] = acc;
    }
}

__global__ void k_conv_products(long n, double b, double c, const double *__restrict__ u,
                                const double *__restrict__ w, const double *__restrict__ uo,
                                const double *__restrict__ wo, double *dxU, const double *__restrict__ dzU,
                                double *dxV, const double *__restrict__ dzV, double *dxT,
                                const double *__restrict__ dzT, const double *__restrict__ dTbc)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double ub = u[i], wb = w[i];
        if (uo != nullptr) {
            ub = b * ub + c * uo[i];
            wb = b * wb + c * wo[i];
        } else if (b != 1.0) {
            ub = b * ub;
            wb = b * wb;
        }
        dxU[i] = dxU[i] * ub + dzU[i] * wb;
        dxV[i] = dxV[i] * ub + dzV[i] * wb;
        double t = dxT[i] * ub + dzT[i] * wb;
        if (dTbc != nullptr) t = t + wb * dTbc[i];
        dxT[i] = t;
    }
}

struct SlabOff {
    int n;
    int off[17];
};

__global__ void k_slab_repack(int dir, double *__restrict__ bundle, double *__restrict__ blocked, int rows, int K,
                              int cols, SlabOff so)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.z;
    if (j >= cols) return;
    int s = 0;
    while (s + 1 < so.n && j >= so.off[s + 1]) ++s;
    const int w = so.off[s + 1] - so.off[s];
    const long bbase = (long)rows * K * so.off[s] + (long)k * w + (j - so.off[s]);
    const long gbase = (long)k * cols + j;
    const long bstep = (long)K * w, gstep = (long)K * cols;
    // 8 rows per thread: independent 8-byte transfers in flight, index math amortised
    const int i0 = blockIdx.y * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int i = i0 + e;
        if (i < rows) {
            if (dir) blocked[bbase + i * bstep] = bundle[gbase + i * gstep];
            else bundle[gbase + i * gstep] = blocked[bbase + i * bstep];
        }
    }
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_sweep(int op, int axis, int n, int njobs, const pde_sweep_job *jobs, void *stream)
{
    PDE_REQUIRE(jobs && njobs >= 1 && njobs <= PDE_MAX_JOBS, "1..8 jobs");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    PDE_REQUIRE(n >= 5, "n >= 5");
    SweepJobs sj{};
    sj.njobs = njobs;
    sj.n = n;
    for (int j = 0; j < njobs; ++j) sj.j[j] = jobs[j];
    cudaStream_t st = as_stream(stream);
    // FULL instantiations (no run-time feature tests) when every job supplies all optional tables
    bool full = true, pow2 = true;
    for (int j = 0; j < njobs; ++j) {
        const pde_sweep_job &jb = jobs[j];
        switch (op) {
        case PDE_SWEEP_DIFF: {
            full = full && jb.flag != 0;
            // x / 2^k == x * 2^-k bit for bit: the division sequence becomes one multiplication
            int ex = 0;
            const double sc = jb.flag != 0 ? jb.sc : 1.0;
            pow2 = pow2 && sc > 0.0 && std::isfinite(sc) && std::frexp(sc, &ex) == 0.5 && ex > -1000 && ex < 1000;
            break;
        }
        case PDE_SWEEP_TDMA_FWD: full = full && jb.tab[0] && jb.tab[4] && jb.in[1]; break;
        case PDE_SWEEP_FDMA_BWD: full = full && jb.tab[4]; break;
        default: break;
        }
    }
#define PDE_SW(OP, what) (full ? launch_sweep<OP<true>>(sj, axis, st, what) : launch_sweep<OP<false>>(sj, axis, st, what))
    switch (op) {
    case PDE_SWEEP_DIFF:
        if (pow2) {
            for (int j = 0; j < njobs; ++j)
                if (sj.j[j].flag == 0) sj.j[j].sc = 1.0;
            return launch_sweep<DiffDesc<true, true>>(sj, axis, st, "pde_sweep(diff, pow2 scale)");
        }
        return PDE_SW(DiffDesc, "pde_sweep(diff)");
    case PDE_SWEEP_TDMA_FWD: return PDE_SW(TdmaFwd, "pde_sweep(tdma fwd)");
    case PDE_SWEEP_TDMA_BWD: return launch_sweep<TdmaBwd<true>>(sj, axis, st, "pde_sweep(tdma bwd)");
    case PDE_SWEEP_FDMA_FWD: return launch_sweep<FdmaFwd<true>>(sj, axis, st, "pde_sweep(fdma fwd)");
    case PDE_SWEEP_FDMA_BWD: return PDE_SW(FdmaBwd, "pde_sweep(fdma bwd)");
    case PDE_SWEEP_TWODMA_BWD: return launch_sweep<TwodmaBwd<false>>(sj, axis, st, "pde_sweep(twodma)");
    default: set_error("pde_sweep: unknown op %d", op); return PDE_ERR_ARG;
    }
#undef PDE_SW
}

int pde_to_cheb_multi(int axis, int njobs, const pde_stencil_job *jobs, void *stream)
{
    PDE_REQUIRE(jobs && njobs >= 1 && njobs <= PDE_MAX_JOBS, "1..8 jobs");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    StencilJobs sj{};
    sj.njobs = njobs;
    int m0 = 0, m1 = 0;
    for (int j = 0; j < njobs; ++j) {
        sj.j[j] = jobs[j];
        const int n0 = axis == 0 ? jobs[j].n_out : jobs[j].batch, n1 = axis == 0 ? jobs[j].batch : jobs[j].n_out;
        m0 = n0 > m0 ? n0 : m0;
        m1 = n1 > m1 ? n1 : m1;
    }
    if (m0 <= 0 || m1 <= 0) return PDE_OK;
    dim3 block(64, 4), grid(ceil_div(m1, 64), ceil_div(m0, 4 * RPT), njobs);
    k_to_cheb_multi<<<grid, block, 0, as_stream(stream)>>>(sj, axis);
    return after_launch("pde_to_cheb_multi");
}

int pde_banded_multi(int axis, int njobs, const pde_band_job *jobs, void *stream)
{
    PDE_REQUIRE(jobs && njobs >= 1 && njobs <= PDE_MAX_JOBS, "1..8 jobs");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    BandJobs bj{};
    bj.njobs = njobs;
    int m0 = 0, m1 = 0, maxd = 0;
    for (int j = 0; j < njobs; ++j) {
        PDE_REQUIRE(jobs[j].ndiag >= 1 && jobs[j].ndiag <= 8, "1..8 diagonals");
        maxd = jobs[j].ndiag > maxd ? jobs[j].ndiag : maxd;
        bj.j[j] = jobs[j];
        const int n0 = axis == 0 ? jobs[j].n_out : jobs[j].batch, n1 = axis == 0 ? jobs[j].batch : jobs[j].n_out;
        m0 = n0 > m0 ? n0 : m0;
        m1 = n1 > m1 ? n1 : m1;
    }
    if (m0 <= 0 || m1 <= 0) return PDE_OK;
    dim3 block(64, 4), grid(ceil_div(m1, 64), ceil_div(m0, 4 * RPB), njobs);
    if (maxd <= 4) k_banded_multi<4><<<grid, block, 0, as_stream(stream)>>>(bj, axis);
    else k_banded_multi<8><<<grid, block, 0, as_stream(stream)>>>(bj, axis);
    return after_launch("pde_banded_multi");
}

int pde_lincomb_multi(int njobs, const pde_lincomb_job *jobs, void *stream)
{
    PDE_REQUIRE(jobs && njobs >= 1 && njobs <= PDE_MAX_JOBS, "1..8 jobs");
    LincombJobs lj{};
    lj.njobs = njobs;
    int m0 = 0, m1 = 0;
    for (int j = 0; j < njobs; ++j) {
        PDE_REQUIRE(jobs[j].nterm >= 1 && jobs[j].nterm <= 4, "1..4 terms");
        lj.j[j] = jobs[j];
        m0 = jobs[j].n0 > m0 ? jobs[j].n0 : m0;
        m1 = jobs[j].n1 > m1 ? jobs[j].n1 : m1;
    }
    if (m0 <= 0 || m1 <= 0) return PDE_OK;
    dim3 block(64, 4), grid(ceil_div(m1, 64), ceil_div(m0, 4 * RPT), njobs);
    k_lincomb_multi<<<grid, block, 0, as_stream(stream)>>>(lj);
    return after_launch("pde_lincomb_multi");
}

int pde_slab_repack(int dir, double *bundle, double *blocked, int rows, int K, int cols, int nranks,
                    const int *col_off, void *stream)
{
    PDE_REQUIRE(bundle && blocked && col_off, "null pointer");
    PDE_REQUIRE(nranks >= 1 && nranks <= 16, "1..16 ranks");
    if (rows <= 0 || K <= 0 || cols <= 0) return PDE_OK;
    SlabOff so{};
    so.n = nranks;
    for (int s = 0; s <= nranks; ++s) so.off[s] = col_off[s];
    dim3 block(256), grid(ceil_div(cols, 256), ceil_div(rows, 8), K);
    k_slab_repack<<<grid, block, 0, as_stream(stream)>>>(dir, bundle, blocked, rows, K, cols, so);
    return after_launch("pde_slab_repack");
}

int pde_conv_products(long n, double b, double c, const double *u, const double *w, const double *u_old,
                      const double *w_old, double *dxU, const double *dzU, double *dxV, const double *dzV,
                      double *dxT, const double *dzT, const double *dTbc, void *stream)
{
    PDE_REQUIRE(u && w && dxU && dzU && dxV && dzV && dxT && dzT, "null pointer");
    PDE_REQUIRE((u_old == nullptr) == (w_old == nullptr), "u_old and w_old go together");
    if (n <= 0) return PDE_OK;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_conv_products<<<grid, 256, 0, as_stream(stream)>>>(n, b, c, u, w, u_old, w_old, dxU, dzU, dxV, dzV, dxT,
                                                         dzT, dTbc);
    return after_launch("pde_conv_products");
}

}  // extern "C"
