// Batched (multi-array) entry points of the C ABI: pointwise stencil / banded / linear-
// combination kernels, the pseudo-spectral products, and the generic sweep dispatcher.
// Compiled with --fmad=false: every product and sum is rounded separately, like the
// NumPy / SciPy expressions of the reference.
#include "common.cuh"
#include <cmath>
#include <cstdint>
#include <cstdlib>
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

// Strip form of the banded product for the stepper's operators (diagonals at off0, off0 + 2, ..., at most 4,
// even sizes and pitches, 16-byte aligned rows): a thread owns TWO adjacent columns (16-byte accesses) and a
// strip of rows, so every operand element is loaded once per thread instead of once per tap.
//   axis 0 (taps along the rows): a register window of RS0 + 6 input rows feeds RS0 output rows
//           (1.75 loads per 2 outputs instead of 8); the coefficients of the strip are warp-uniform 16-byte loads;
//   axis 1 (taps along the contiguous axis): the coefficients of the two columns are loaded once per thread and
//           reused over RS1 rows; the taps of a row are nd aligned 16-byte loads.
// Same products, same order of sums, zero coefficients skipped: bit-identical to k_banded_multi.
#ifndef PDE_BAND_RS0
#define PDE_BAND_RS0 4
#endif
#ifndef PDE_BAND_RS1
#define PDE_BAND_RS1 4
#endif
#ifndef PDE_BAND_MINB
#define PDE_BAND_MINB 2
#endif
constexpr int RS0 = PDE_BAND_RS0, RS1 = PDE_BAND_RS1;

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

template <int AXIS>
__global__ void __launch_bounds__(256, PDE_BAND_MINB) k_banded_strip(BandJobs jobs)
{
    const pde_band_job &jb = jobs.j[blockIdx.z];
    const int n0 = AXIS == 0 ? jb.n_out : jb.batch, n1 = AXIS == 0 ? jb.batch : jb.n_out;
    const int jp = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (jp >= n1) return;
    const int nd = jb.ndiag, o0 = jb.off[0];
    if (AXIS == 0) {
        const int i0 = (blockIdx.y * blockDim.y + threadIdx.y) * RS0;
        if (i0 >= n0) return;
        constexpr int W = RS0 + 6;
        double2 xw[W], yo[RS0];
        double a[4][RS0];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int c = i0 + o0 + w;
            xw[w] = (w < RS0 + 2 * (nd - 1) && c >= 0 && c < jb.n_in) ? ld2(jb.x + (long)c * jb.ldx + jp)
                                                                      : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int d = 0; d < 4; ++d)
#pragma unroll
            for (int r = 0; r < RS0; r += 2) {
                // rows i0 + r, i0 + r + 1 of diagonal d (n_out even, i0 a multiple of RS0: aligned, in range together)
                const double2 t = (d < nd && i0 + r < n0) ? ldg2(jb.diags + (long)d * jb.n_out + i0 + r)
                                                          : make_double2(0.0, 0.0);
                a[d][r] = t.x;
                a[d][r + 1] = t.y;
            }
        if (jb.accumulate) {
#pragma unroll
            for (int r = 0; r < RS0; ++r)
                yo[r] = i0 + r < n0 ? ld2(jb.y + (long)(i0 + r) * jb.ldy + jp) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int r = 0; r < RS0; ++r) {
            const int i = i0 + r;
            if (i >= n0) break;
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int c = i + o0 + 2 * d;
                if (d < nd && c >= 0 && c < jb.n_in && a[d][r] != 0.0) {
                    ax = ax + a[d][r] * xw[r + 2 * d].x;
                    ay = ay + a[d][r] * xw[r + 2 * d].y;
                }
            }
            double2 *yp = reinterpret_cast<double2 *>(jb.y + (long)i * jb.ldy + jp);
            *yp = jb.accumulate ? make_double2(yo[r].x + ax, yo[r].y + ay) : make_double2(ax, ay);
        }
    } else {
        const int i0 = (blockIdx.y * blockDim.y + threadIdx.y) * RS1;
        if (i0 >= n0) return;
        double2 a[4], xd[RS1][4], yo[RS1];
        bool ok[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int c = jp + o0 + 2 * d;
            ok[d] = d < nd && c >= 0 && c < jb.n_in;
            a[d] = ok[d] ? ldg2(jb.diags + (long)d * jb.n_out + jp) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int r = 0; r < RS1; ++r) {
            const int i = i0 + r;
#pragma unroll
            for (int d = 0; d < 4; ++d)
                xd[r][d] = (ok[d] && i < n0) ? ld2(jb.x + (long)i * jb.ldx + jp + o0 + 2 * d) : make_double2(0.0, 0.0);
            if (jb.accumulate) yo[r] = i < n0 ? ld2(jb.y + (long)i * jb.ldy + jp) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int r = 0; r < RS1; ++r) {
            const int i = i0 + r;
            if (i >= n0) break;
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (ok[d] && a[d].x != 0.0) ax = ax + a[d].x * xd[r][d].x;
                if (ok[d] && a[d].y != 0.0) ay = ay + a[d].y * xd[r][d].y;
            }
            double2 *yp = reinterpret_cast<double2 *>(jb.y + (long)i * jb.ldy + jp);
            *yp = jb.accumulate ? make_double2(yo[r].x + ax, yo[r].y + ay) : make_double2(ax, ay);
        }
    }
}

static bool band_strip_ok(const pde_band_job &jb, int axis)
{
    if (jb.ndiag < 1 || jb.ndiag > 4) return false;
    for (int d = 1; d < jb.ndiag; ++d)
        if (jb.off[d] != jb.off[0] + 2 * d) return false;
    if (jb.off[0] % 2 != 0 || jb.off[0] < -2 || jb.off[0] > 0) return false;
    auto al = [](const void *q) { return ((uintptr_t)q % 16) == 0; };
    if (!al(jb.x) || !al(jb.y) || !al(jb.diags) || jb.ldx % 2 || jb.ldy % 2) return false;
    if (jb.n_out % 2 || jb.n_in % 2 || jb.batch % 2) return false;
    (void)axis;
    return true;
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

// the same products for `gridDim.y` members whose arrays lie `stride` elements apart (dTbc is shared)
__global__ void k_conv_products_members(long n, long stride, double b, double c, const double *__restrict__ u,
                                        const double *__restrict__ w, const double *__restrict__ uo,
                                        const double *__restrict__ wo, double *dxU, const double *__restrict__ dzU,
                                        double *dxV, const double *__restrict__ dzV, double *dxT,
                                        const double *__restrict__ dzT, const double *__restrict__ dTbc)
{
    const long o = (long)blockIdx.y * stride;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double ub = u[o + i], wb = w[o + i];
        if (uo != nullptr) {
            ub = b * ub + c * uo[o + i];
            wb = b * wb + c * wo[o + i];
        } else if (b != 1.0) {
            ub = b * ub;
            wb = b * wb;
        }
        dxU[o + i] = dxU[o + i] * ub + dzU[o + i] * wb;
        dxV[o + i] = dxV[o + i] * ub + dzV[o + i] * wb;
        double t = dxT[o + i] * ub + dzT[o + i] * wb;
        if (dTbc != nullptr) t = t + wb * dTbc[i];
        dxT[o + i] = t;
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

namespace pde {
// DiffDesc with the power-of-two scale, in the <FULL, TAB> shape the tiled launcher expects
template <bool FULL, class TAB>
using DiffPow2 = DiffDesc<FULL, true, TAB>;
}  // namespace pde

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
    // axis 1: tiled kernel (k_sweep_tile) when sizes / alignment allow; PDE_SWEEP_TILE=0 keeps k_sweep
    static const bool tile = !(getenv("PDE_SWEEP_TILE") && atoi(getenv("PDE_SWEEP_TILE")) == 0);
#define PDE_TILE(OPT, FULLV, what)                                                           \
    if (tile && axis == 1 && sweep_tile_ok<OPT<FULLV, TabTile>>(sj)) return launch_sweep_tile<OPT, FULLV>(sj, st, what)
#define PDE_SW(OP, what) (full ? launch_sweep<OP<true>>(sj, axis, st, what) : launch_sweep<OP<false>>(sj, axis, st, what))
    switch (op) {
    case PDE_SWEEP_DIFF:
        if (pow2) {
            for (int j = 0; j < njobs; ++j)
                if (sj.j[j].flag == 0) sj.j[j].sc = 1.0;
            PDE_TILE(DiffPow2, true, "pde_sweep(diff, pow2 scale, tiled)");
            return launch_sweep<DiffDesc<true, true>>(sj, axis, st, "pde_sweep(diff, pow2 scale)");
        }
        return PDE_SW(DiffDesc, "pde_sweep(diff)");
    case PDE_SWEEP_TDMA_FWD:
        if (full) PDE_TILE(TdmaFwd, true, "pde_sweep(tdma fwd, tiled)");
        return PDE_SW(TdmaFwd, "pde_sweep(tdma fwd)");
    case PDE_SWEEP_TDMA_BWD:
        PDE_TILE(TdmaBwd, true, "pde_sweep(tdma bwd, tiled)");
        return launch_sweep<TdmaBwd<true>>(sj, axis, st, "pde_sweep(tdma bwd)");
    case PDE_SWEEP_FDMA_FWD:
        PDE_TILE(FdmaFwd, true, "pde_sweep(fdma fwd, tiled)");
        return launch_sweep<FdmaFwd<true>>(sj, axis, st, "pde_sweep(fdma fwd)");
    case PDE_SWEEP_FDMA_BWD:
        if (full) PDE_TILE(FdmaBwd, true, "pde_sweep(fdma bwd, tiled)");
        return PDE_SW(FdmaBwd, "pde_sweep(fdma bwd)");
    case PDE_SWEEP_TWODMA_BWD: return launch_sweep<TwodmaBwd<false>>(sj, axis, st, "pde_sweep(twodma)");
    default: set_error("pde_sweep: unknown op %d", op); return PDE_ERR_ARG;
    }
#undef PDE_TILE
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
    static const bool strip = !(getenv("PDE_BANDED_STRIP") && atoi(getenv("PDE_BANDED_STRIP")) == 0);
    bool strip_ok = strip;
    for (int j = 0; j < njobs && strip_ok; ++j) strip_ok = band_strip_ok(jobs[j], axis);
    if (strip_ok) {
        dim3 block(64, 4), grid(ceil_div(m1, 128), ceil_div(m0, 4 * (axis == 0 ? RS0 : RS1)), njobs);
        if (axis == 0) k_banded_strip<0><<<grid, block, 0, as_stream(stream)>>>(bj);
        else k_banded_strip<1><<<grid, block, 0, as_stream(stream)>>>(bj);
        return after_launch("pde_banded_multi(strip)");
    }
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

int pde_conv_products_members(long n, int nmembers, long stride, double b, double c, const double *u, const double *w,
                              const double *u_old, const double *w_old, double *dxU, const double *dzU, double *dxV,
                              const double *dzV, double *dxT, const double *dzT, const double *dTbc, void *stream)
{
    PDE_REQUIRE(u && w && dxU && dzU && dxV && dzV && dxT && dzT, "null pointer");
    PDE_REQUIRE((u_old == nullptr) == (w_old == nullptr), "u_old and w_old go together");
    PDE_REQUIRE(nmembers >= 1 && nmembers <= 65535 && stride >= n, "members / stride");
    if (n <= 0) return PDE_OK;
    const int gx = (int)((n + 255) / 256 < 148 * 4 ? (n + 255) / 256 : 148 * 4);
    k_conv_products_members<<<dim3(gx, nmembers), 256, 0, as_stream(stream)>>>(n, stride, b, c, u, w, u_old, w_old, dxU,
                                                                              dzU, dxV, dzV, dxT, dzT, dTbc);
    return after_launch("pde_conv_products_members");
}

}  // extern "C"
