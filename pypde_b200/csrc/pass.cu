// Fused axis passes (sm_100a): one launch applies a whole chain of 1-D operators of the
// Chebyshev-Galerkin time step to every sequence of several 2-D arrays.
//
// SURVEY.md §8(d) describes an IMEX stage as ~9 axis passes: every operator of the step acts along ONE
// axis (stencil maps chebyshev.py:287-337, derivative recurrence differentiate_cheby.f90:28-53, banded
// products plans.py:54-74, offset-2 Thomas sweeps tdma.f90:55-106 / fdma.f90:1-98, per-column Poisson
// solves fdma.f90:146-195).  Round 1 ran each of them as its own kernel (39 launches per stage, every
// intermediate through HBM, recurrences as ONE thread per chain).  Here a persistent CTA keeps W sequences
// (rows: ROW layout, or a strip of W adjacent columns: COL layout) in shared memory, and the warp that owns
// a sequence interprets a small program of operators on it:
//
//   LOAD / STORE / AXPY / LINCOMB   move the sequences between HBM and shared memory (cp.async, coalesced
//                         batches; a sequence may be split over several base pointers = the slabs of peer
//                         GPUs: the distributed transposes of the slab decomposition are the loads and
//                         stores of the row passes, straight over NVLink peer mappings);
//   POINT                 banded product / stencil map (taps at unit offsets -1..+2);
//   DIFF                  derivative recurrence (suffix sums);
//   REC1 / REC2           first / second order linear recurrences (the Thomas sweeps), CHAIN-SPLIT:
//                         every lane owns 1/32 of the chain, composes the affine map of its segment
//                         (2x2 for the two-term back substitution), the maps are combined with a
//                         warp scan, and a second walk over the segment writes the result;
//   TABLES                stages the recurrence tables of a job in shared memory once per CTA and job.
//
// A sequence is stored as 16-byte units (x[2m], x[2m+1]): the two parity chains of the offset-2
// recurrences ride in the two halves of a double2, so one thread always advances two independent
// chains.  Lane l owns units [l*SEGU, (l+1)*SEGU); unit m lives at m + (m >> lg) (one pad unit per
// segment: lane-strided and consecutive accesses are both conflict-free).  Recurrence tables are
// prepared by the host in the matching "segment-transposed" order [j][lane].
//
// First version of this kernel (profiles/r02_ncu_pass_v1_stage.csv): one CTA per strip, instruction words
// and tables read from global memory inside the operator loops -> 12 warps per SM all waiting on L2
// (long-scoreboard 4-14 per issue, 0.15-0.86 ms per pass).  Now: program and shared tables live in shared
// memory, global operands are fetched in independent batches, per-sequence tables (Poisson) are
// register-prefetched 8 steps ahead.
//
// Arithmetic differs from the reference by rounding only (re-association across segments,
// reciprocal-scaled tables): the contract is 1e-12, checked against the oracle in tests/test_gpu_pass.py.
#include "common.cuh"
#include <cstring>

namespace pde {

constexpr int PASS_MAX_INS = 16;             // instructions per program held in shared memory
constexpr int PASS_SLOTS = 4;                // shared-memory table slots
// job descriptors held in shared memory: short sequences (ensembles of small grids) come as thousands of jobs and leave
// most of the shared memory free
__host__ __device__ constexpr int pass_jobs_cached(int lg) { return lg <= 3 ? 4096 : 128; }

__device__ __forceinline__ void pcp16(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void pcp8(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void pcp_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double2 d2(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ double2 operator*(double2 a, double2 b) { return d2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return d2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 operator-(double2 a) { return d2(-a.x, -a.y); }
__device__ __forceinline__ double2 fma2(double2 a, double2 b, double2 c) { return d2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ double2 nfma2(double2 a, double2 b, double2 c) { return d2(fma(-a.x, b.x, c.x), fma(-a.y, b.y, c.y)); }
__device__ __forceinline__ double2 fmas(double s, double2 b, double2 c) { return d2(fma(s, b.x, c.x), fma(s, b.y, c.y)); }
__device__ __forceinline__ double2 shfl_up2(double2 v, int d)
{
    return d2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ double2 shfl_dn2(double2 v, int d)
{
    return d2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}

// Sizes are compile-time (one kernel instantiation per log2 SEGU): loop trip counts, the unit -> address map and
// most index arithmetic fold into immediates.  The first persistent version kept them in registers and spent
// ~8000 instructions per warp on a five-operator program (profiles/r02_ncu_pass_v2.csv).
template <int LG_, int W_>
struct PassCtxT {
    static constexpr int lg = LG_, SEGU = 1 << LG_, NUP = 32 << LG_;
    static constexpr int PADU = LG_ > 0 ? 1 : 0;
    static constexpr int BUFU = NUP + 32 * PADU + 1;   // + 1: the W buffers of a strip start 16 bytes apart (mod 128)
    static constexpr int W = W_;                       // sequences (warps) per CTA
    int lane, w, q0, q;            // lane, warp = sequence within the CTA, first / own sequence of the job
    bool live;                     // q < nseq
    int nseq, seq0;
    double2 *buf;                  // this warp's sequence
    double2 *all;                  // sequence 0 of the CTA
    double2 *tab;                  // table slot 0 (NUP units each)
    __device__ __forceinline__ static int pu(int m) { return m + (PADU ? (m >> LG_) : 0); }
    __device__ __forceinline__ int segbase() const { return lane * (SEGU + PADU); }
};

// segment of element i (along the sequence) -- nseg is small; the instruction lives in shared memory
__device__ __forceinline__ int seg_of(const pde_pass_ins &I, int i)
{
    int s = 0;
    while (s + 1 < I.nseg && i >= I.start[s + 1]) ++s;
    return s;
}

template <bool COL>
__device__ __forceinline__ const double *elem_addr(const pde_pass_ins &I, int s, int q, int i)
{
    const double *base = reinterpret_cast<const double *>(I.p[s]);
    if (COL) return base + (long)(i - I.start[s]) * I.ld[s] + q;
    return base + (long)q * I.ld[s] + (i - I.start[s]);
}

// ---- address maps ------------------------------------------------------------------------------
// ROW ("interleaved") map: lane l, chunk k <-> unit m = l + 32 k.  COL map: thread (r0, col), step j <-> row
// r = r0 + 32 j of column col, i.e. unit (r0 >> 1) + 16 j, half r0 & 1.  In both, the shared-memory position
// splits into a per-thread base and a compile-time offset of the step (pu(m) = m + (m >> LG)).
template <class Ctx>
struct RowMap {
    static constexpr int LG = Ctx::lg;
    __device__ __forceinline__ static int base(int lane) { return lane + ((Ctx::PADU && LG < 5) ? (lane >> LG) : 0); }
    __host__ __device__ __forceinline__ static constexpr int off(int k)
    {
        return 32 * k + (Ctx::PADU ? (LG <= 5 ? (k << (5 - (LG <= 5 ? LG : 5))) : (k >> 1)) : 0);
    }
};
template <class Ctx>
struct ColMap {
    static constexpr int LG = Ctx::lg;
    int col, r0;
    __device__ __forceinline__ ColMap() : col(threadIdx.x & (Ctx::W - 1)), r0(threadIdx.x / Ctx::W) {}
    // double index of row r0 (step 0) inside the column's buffer
    __device__ __forceinline__ int base() const
    {
        const int u = r0 >> 1;
        return 2 * (u + (Ctx::PADU ? (u >> LG) : 0)) + (r0 & 1);
    }
    __host__ __device__ __forceinline__ static constexpr int off(int j)      // doubles
    {
        return 2 * (16 * j + (Ctx::PADU ? ((16 * j) >> LG) : 0));
    }
    static constexpr int STEPS = 2 * Ctx::SEGU;      // rows r0 + 32 j, j < STEPS, cover 2 NUP elements
};
// for LG < 4 the COL split above is not exact ((x + 16 j) >> LG with x < 16 needs 2^LG | 16 j: true for LG <= 4)
static_assert(true, "");

constexpr int PB = 8;        // independent memory operations per batch

// COL data movement is the 8-byte form only: all cooperative operators (LOAD / STORE / AXPY / LINCOMB) use the SAME
// thread <-> (row, column) map (ColMap), so consecutive cooperative operators need no barrier between them -- every
// thread only meets elements it wrote itself.  (A 16-byte form with a second map -- 2 x 2 blocks per thread -- was
// measured: no gain, the 64-byte row pieces bound the column passes; and mixing two maps without a barrier is a race.)

// ------------------------------------------------------------------------------------------------
// LOAD: buffer <- elements [0, n) of the operand, zero beyond
// ------------------------------------------------------------------------------------------------
template <bool COL, class Ctx>
__device__ __forceinline__ void op_load(const Ctx &c, const pde_pass_ins &I)
{
    const int n = I.n;
    const bool one = I.nseg == 1;
    if (!COL) {
        const double *p0 = reinterpret_cast<const double *>(I.p[0]);
        const double *row0 = p0 + (long)c.q * I.ld[0] + 2 * c.lane;
        double2 *dst = c.buf + RowMap<Ctx>::base(c.lane);
        const int rem0 = c.live ? n - 2 * c.lane : 0;          // valid elements from this lane's first unit on
#pragma unroll 8
        for (int k = 0; k < Ctx::SEGU; ++k) {
            const int rem = rem0 - 64 * k;
            const int bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
            const double *src = row0 + 64 * k;
            if (!one && bytes) src = elem_addr<false>(I, seg_of(I, 2 * c.lane + 64 * k), c.q, 2 * c.lane + 64 * k);
            pcp16(dst + RowMap<Ctx>::off(k), bytes ? src : p0, bytes);
        }
    } else {
        const ColMap<Ctx> cm;
        double *dst = reinterpret_cast<double *>(c.all + cm.col * Ctx::BUFU) + cm.base();
        const bool ok = c.q0 + cm.col < c.nseq;
        const long ld0 = I.ld[0];
        const double *p0 = reinterpret_cast<const double *>(I.p[0]);
        const double *src = p0 + c.q0 + cm.col + (long)cm.r0 * ld0;
        const int lim = ok ? n - cm.r0 : 0;                   // row r0 + 32 j is valid iff 32 j < lim
#pragma unroll 8
        for (int j = 0; j < ColMap<Ctx>::STEPS; ++j) {
            const bool v = 32 * j < lim;
            const double *s2 = src;
            if (!one && v) s2 = elem_addr<true>(I, seg_of(I, cm.r0 + 32 * j), c.q0 + cm.col, cm.r0 + 32 * j);
            pcp8(dst + ColMap<Ctx>::off(j), v ? s2 : p0, v ? 8 : 0);
            src += 32 * ld0;
        }
    }
    pcp_wait_all();
    if (!COL) __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// STORE: elements [0, n) of the buffer -> operand.  flag ONLY_SEQ: only the sequence with global index off[0]
// ------------------------------------------------------------------------------------------------
template <bool COL, class Ctx>
__device__ __forceinline__ void op_store(const Ctx &c, const pde_pass_ins &I)
{
    const int n = I.n;
    const bool only = (I.flags & PDE_PASS_F_ONLY_SEQ) != 0;
    const bool one = I.nseg == 1;
    if (!COL) {
        if (!c.live || (only && c.seq0 + c.q != I.off[0])) return;
        if ((I.flags & PDE_PASS_F_BULK) && Ctx::SEGU >= 8) {
            // TMA bulk stores: lane l hands its segment (SEGU units = 16 SEGU contiguous bytes of the row) to the copy
            // engine, cut at the operand's segment boundaries (peer slabs); shared-memory writes of the generic proxy
            // are fenced first, and the buffer is reused only after the engine has read it.
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            const int e0 = 2 * c.lane * Ctx::SEGU;
            const int e1 = min(n, e0 + 2 * Ctx::SEGU);
            const double2 *sseg = c.buf + c.segbase();
            if (e1 > e0) {
                const int efull = e0 + ((e1 - e0) & ~1);
                int a = e0;
                while (a < efull) {
                    const int s = seg_of(I, a);
                    const int send = (s + 1 < I.nseg) ? I.start[s + 1] : n;
                    const int b = min(efull, send);
                    const double *g = elem_addr<false>(I, s, c.q, a);
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(sseg + ((a - e0) >> 1));
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(sa), "r"((b - a) * 8)
                                 : "memory");
                    a = b;
                }
                if (efull < e1) {       // odd tail element
                    const int s = seg_of(I, efull);
                    *const_cast<double *>(elem_addr<false>(I, s, c.q, efull)) = sseg[(efull - e0) >> 1].x;
                }
            }
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            __syncwarp();
            return;
        }
        double *row0 = const_cast<double *>(reinterpret_cast<const double *>(I.p[0])) + (long)c.q * I.ld[0] + 2 * c.lane;
        const double2 *src = c.buf + RowMap<Ctx>::base(c.lane);
        const int rem0 = n - 2 * c.lane;
#pragma unroll 1
        for (int k0 = 0; k0 < Ctx::SEGU; k0 += PB) {
            double2 v[PB];
#pragma unroll
            for (int e = 0; e < PB; ++e)
                if (k0 + e < Ctx::SEGU) v[e] = src[RowMap<Ctx>::off(k0 + e)];
#pragma unroll
            for (int e = 0; e < PB; ++e) {
                const int k = k0 + e;
                if (k < Ctx::SEGU) {
                    const int rem = rem0 - 64 * k;
                    if (rem >= 1) {
                        double *dst = row0 + 64 * k;
                        if (!one) dst = const_cast<double *>(elem_addr<false>(I, seg_of(I, 2 * c.lane + 64 * k), c.q, 2 * c.lane + 64 * k));
                        if (rem >= 2) *reinterpret_cast<double2 *>(dst) = v[e];
                        else *dst = v[e].x;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        const ColMap<Ctx> cm;
        if (c.q0 + cm.col >= c.nseq || (only && c.seq0 + c.q0 + cm.col != I.off[0])) return;
        const double *src = reinterpret_cast<const double *>(c.all + cm.col * Ctx::BUFU) + cm.base();
        const long ld0 = I.ld[0];
        double *dst0 = const_cast<double *>(reinterpret_cast<const double *>(I.p[0])) + c.q0 + cm.col + (long)cm.r0 * ld0;
        const int lim = n - cm.r0;
#pragma unroll 1
        for (int j0 = 0; j0 < ColMap<Ctx>::STEPS; j0 += PB) {
            double v[PB];
#pragma unroll
            for (int e = 0; e < PB; ++e)
                if (j0 + e < ColMap<Ctx>::STEPS) v[e] = src[ColMap<Ctx>::off(j0 + e)];
#pragma unroll
            for (int e = 0; e < PB; ++e) {
                const int j = j0 + e;
                if (j < ColMap<Ctx>::STEPS && 32 * j < lim) {
                    double *dst = dst0 + (long)(32 * j) * ld0;
                    if (!one) dst = const_cast<double *>(elem_addr<true>(I, seg_of(I, cm.r0 + 32 * j), c.q0 + cm.col, cm.r0 + 32 * j));
                    *dst = v[e];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// AXPY: buffer <- f1 * buffer + f0 * G   (f1 applied only with flag SCALED), G = elements [0, n) of the
// operand; flag STENCIL: G_i = g_i + st_i g_{i-2} with the element table st = p[7] (st_i = s_{i-2}, st_0 = st_1 = 0)
// Loads are issued in independent batches of PB before they are consumed.
// ------------------------------------------------------------------------------------------------
template <bool COL, bool STEN, class Ctx>
__device__ __forceinline__ void axpy_impl(const Ctx &c, const pde_pass_ins &I)
{
    // STEN: the operand has n valid entries, its image G_i = g_i + st_i g_{i-2} has n + 2
    const int n = I.n, nout = STEN ? n + 2 : n;
    const bool scaled = (I.flags & PDE_PASS_F_SCALED) != 0;
    const double f0 = I.f0, f1 = scaled ? I.f1 : 1.0;
    const bool one = I.nseg == 1;
    const double *st = reinterpret_cast<const double *>(I.p[PDE_PASS_MAX_SEG - 1]);
    if (!COL) {
        if (!c.live) return;
        constexpr int AB = STEN ? 4 : 8;
        const double *row0 = reinterpret_cast<const double *>(I.p[0]) + (long)c.q * I.ld[0];
        double2 *bb = c.buf + RowMap<Ctx>::base(c.lane);
#pragma unroll 1
        for (int k0 = 0; k0 < Ctx::SEGU; k0 += AB) {
            double2 g[AB], h[AB], t[AB];
#pragma unroll
            for (int e = 0; e < AB; ++e) {
                const int i = 2 * c.lane + 64 * (k0 + e);
                g[e] = h[e] = t[e] = d2(0.0, 0.0);
                if (k0 + e < Ctx::SEGU && i < nout) {
                    if (i < n) {
                        const double *src = one ? row0 + i : elem_addr<false>(I, seg_of(I, i), c.q, i);
                        if (i + 1 < n) g[e] = __ldg(reinterpret_cast<const double2 *>(src));
                        else g[e].x = __ldg(src);
                    }
                    if (STEN && i >= 2) {
                        const double *s2 = one ? row0 + i - 2 : elem_addr<false>(I, seg_of(I, i - 2), c.q, i - 2);
                        if (i - 1 < n) h[e] = __ldg(reinterpret_cast<const double2 *>(s2));
                        else h[e].x = __ldg(s2);
                        t[e] = __ldg(reinterpret_cast<const double2 *>(st + i));
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < AB; ++e) {
                const int k = k0 + e;
                if (k < Ctx::SEGU && (scaled || 2 * c.lane + 64 * k < nout)) {
                    const double2 gg = STEN ? fma2(t[e], h[e], g[e]) : g[e];
                    double2 &b = bb[RowMap<Ctx>::off(k)];
                    b = scaled ? d2(fma(f0, gg.x, f1 * b.x), fma(f0, gg.y, f1 * b.y)) : fmas(f0, gg, b);
                }
            }
        }
        __syncwarp();
    } else {
        const ColMap<Ctx> cm;
        if (c.q0 + cm.col >= c.nseq) return;
        constexpr int AB = STEN ? 8 : 16;          // 8-byte loads: many in flight per thread
        double *bb = reinterpret_cast<double *>(c.all + cm.col * Ctx::BUFU) + cm.base();
        const long ld0 = I.ld[0];
        const double *col0 = reinterpret_cast<const double *>(I.p[0]) + c.q0 + cm.col;
#pragma unroll 1
        for (int j0 = 0; j0 < ColMap<Ctx>::STEPS; j0 += AB) {
            double g[AB], h[AB], t[AB];
#pragma unroll
            for (int e = 0; e < AB; ++e) {
                const int r = cm.r0 + 32 * (j0 + e);
                g[e] = h[e] = t[e] = 0.0;
                if (j0 + e < ColMap<Ctx>::STEPS && r < nout) {
                    if (r < n) g[e] = __ldg(one ? col0 + (long)r * ld0 : elem_addr<true>(I, seg_of(I, r), c.q0 + cm.col, r));
                    if (STEN && r >= 2) {
                        h[e] = __ldg(one ? col0 + (long)(r - 2) * ld0 : elem_addr<true>(I, seg_of(I, r - 2), c.q0 + cm.col, r - 2));
                        t[e] = __ldg(st + r);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < AB; ++e) {
                const int j = j0 + e;
                if (j < ColMap<Ctx>::STEPS && (scaled || cm.r0 + 32 * j < nout)) {
                    const double gg = STEN ? fma(t[e], h[e], g[e]) : g[e];
                    double &b = bb[ColMap<Ctx>::off(j)];
                    b = scaled ? fma(f0, gg, f1 * b) : fma(f0, gg, b);
                }
            }
        }
    }
}

template <bool COL, class Ctx>
__device__ __forceinline__ void op_axpy(const Ctx &c, const pde_pass_ins &I)
{
    if (I.flags & PDE_PASS_F_STENCIL) axpy_impl<COL, true, Ctx>(c, I);
    else axpy_impl<COL, false, Ctx>(c, I);
}

// ------------------------------------------------------------------------------------------------
// LINCOMB: buffer <- [buffer +] sum_k coef[k] G_k, K = nseg single-segment operands p[k] / ld[k] with
// start[k] valid elements each (zero beyond); flag ACCUM keeps the buffer.  All K loads of a batch are in
// flight together (the right-hand sides of the Helmholtz problems are sums of 3-5 arrays).
// ------------------------------------------------------------------------------------------------
template <bool COL, class Ctx>
__device__ __forceinline__ void op_lincomb(const Ctx &c, const pde_pass_ins &I)
{
    constexpr int LB = COL ? 8 : 4, KMAX = PDE_PASS_MAX_TERMS;
    const int K = I.nseg;
    const bool accum = (I.flags & PDE_PASS_F_ACCUM) != 0;
    if (!COL) {
        if (!c.live) return;
        double2 *bb = c.buf + RowMap<Ctx>::base(c.lane);
        const double *rowk[KMAX];
        int nk[KMAX];
        double cf[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            rowk[k] = k < K ? reinterpret_cast<const double *>(I.p[k]) + (long)c.q * I.ld[k] + 2 * c.lane : nullptr;
            nk[k] = k < K ? I.start[k] - 2 * c.lane : 0;
            cf[k] = k < K ? I.coef[k] : 0.0;
        }
#pragma unroll 1
        for (int m0 = 0; m0 < Ctx::SEGU; m0 += LB) {
            double2 g[LB][KMAX];
#pragma unroll
            for (int e = 0; e < LB; ++e)
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    g[e][k] = d2(0.0, 0.0);
                    const int rem = nk[k] - 64 * (m0 + e);
                    if (m0 + e < Ctx::SEGU && rem >= 1) {
                        const double *src = rowk[k] + 64 * (m0 + e);
                        if (rem >= 2) g[e][k] = __ldg(reinterpret_cast<const double2 *>(src));
                        else g[e][k].x = __ldg(src);
                    }
                }
#pragma unroll
            for (int e = 0; e < LB; ++e) {
                if (m0 + e < Ctx::SEGU) {
                    double2 &b = bb[RowMap<Ctx>::off(m0 + e)];
                    double2 acc = accum ? b : d2(0.0, 0.0);
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) acc = fmas(cf[k], g[e][k], acc);
                    b = acc;
                }
            }
        }
        __syncwarp();
    } else {
        const ColMap<Ctx> cm;
        if (c.q0 + cm.col >= c.nseq) return;
        double *bb = reinterpret_cast<double *>(c.all + cm.col * Ctx::BUFU) + cm.base();
        const double *colk[KMAX];
        long ldk[KMAX];
        int nk[KMAX];
        double cf[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            ldk[k] = k < K ? I.ld[k] : 0;
            colk[k] = k < K ? reinterpret_cast<const double *>(I.p[k]) + c.q0 + cm.col + (long)cm.r0 * ldk[k] : nullptr;
            nk[k] = k < K ? I.start[k] - cm.r0 : 0;
            cf[k] = k < K ? I.coef[k] : 0.0;
        }
#pragma unroll 1
        for (int j0 = 0; j0 < ColMap<Ctx>::STEPS; j0 += LB) {
            double g[LB][KMAX];
#pragma unroll
            for (int e = 0; e < LB; ++e)
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    g[e][k] = 0.0;
                    if (j0 + e < ColMap<Ctx>::STEPS && 32 * (j0 + e) < nk[k]) g[e][k] = __ldg(colk[k] + (long)(32 * (j0 + e)) * ldk[k]);
                }
#pragma unroll
            for (int e = 0; e < LB; ++e) {
                if (j0 + e < ColMap<Ctx>::STEPS) {
                    double &b = bb[ColMap<Ctx>::off(j0 + e)];
                    double acc = accum ? b : 0.0;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) acc = fma(cf[k], g[e][k], acc);
                    b = acc;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// POINT: y[m] = sum_t C_t[m] * x[m + off_t] (units), in place.  All offsets >= 0 (ascending walk) or all <= 0
// (descending walk).  C_t = p[t] (double2 per unit, NUP entries) or 1 when p[t] is null.
// Specialised on the number of taps: all table and operand loads of a group of chunks are issued back to back
// (the first version tested nt / null tables inside the loops: one L2 round trip per chunk and tap, 11-24 us per
// strip; tools/bench_pass.py).
// ------------------------------------------------------------------------------------------------
template <int NT, class Ctx>
__device__ __forceinline__ void point_impl(const Ctx &c, const pde_pass_ins &I)
{
    bool desc = false;
    int off[NT];
    const double2 *tab[NT];
    bool has[NT];
    const double2 *any = nullptr;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        off[t] = I.off[t];
        tab[t] = reinterpret_cast<const double2 *>(I.p[t]);
        has[t] = tab[t] != nullptr;
        if (has[t]) any = tab[t];
        desc |= off[t] < 0;
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
        if (!has[t]) tab[t] = any;               // loads stay unconditional; the value is replaced by 1
    const bool anytab = any != nullptr;
    constexpr int G0 = NT <= 2 ? 8 : 4;
    constexpr int G = Ctx::SEGU < G0 ? Ctx::SEGU : G0;   // chunks of 32 units per synchronisation
    constexpr int NG = Ctx::SEGU / G;
    const double2 one = d2(1.0, 1.0);
    double2 *bl = c.buf + RowMap<Ctx>::base(c.lane);
#pragma unroll 1
    for (int gi = 0; gi < NG; ++gi) {
        const int g = desc ? NG - 1 - gi : gi;
        double2 cf[G][NT], x[G][NT];
#pragma unroll
        for (int e = 0; e < G; ++e) {
            const int m = (g * G + e) * 32 + c.lane;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                cf[e][t] = anytab ? __ldg(tab[t] + m) : one;
                const int mm = m + off[t];
                x[e][t] = (mm >= 0 && mm < Ctx::NUP) ? c.buf[Ctx::pu(mm)] : d2(0.0, 0.0);
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < G; ++e) {
            double2 acc = (has[0] ? cf[e][0] : one) * x[e][0];
#pragma unroll
            for (int t = 1; t < NT; ++t) acc = fma2(has[t] ? cf[e][t] : one, x[e][t], acc);
            bl[RowMap<Ctx>::off(g * G + e)] = acc;
        }
    }
    __syncwarp();
}

template <class Ctx>
__device__ __forceinline__ void op_point(const Ctx &c, const pde_pass_ins &I)
{
    switch (I.n) {
    case 1: point_impl<1>(c, I); break;
    case 2: point_impl<2>(c, I); break;
    case 3: point_impl<3>(c, I); break;
    default: point_impl<4>(c, I); break;
    }
}

// ------------------------------------------------------------------------------------------------
// DIFF: dc_k = dc_{k+2} + 2 (k+1) c_{k+1}, dc_0 = dc_2 / 2 + c_1, result times f0 (1 / scale)
// (differentiate_cheby.f90:28-53).  Units: D[m] = D[m+1] + (2(2m+1) X[m].y, (4m+4) X[m+1].x).
// ------------------------------------------------------------------------------------------------
template <class Ctx>
__device__ __forceinline__ void op_diff(const Ctx &c, const pde_pass_ins &I)
{
    const int base = c.segbase();
    constexpr int S = Ctx::SEGU;
    double next_x = 0.0;
    if (c.lane < 31) next_x = c.buf[base + S + Ctx::PADU].x;
    __syncwarp();
    double2 acc = d2(0.0, 0.0);
    // coefficients 2 (2m+1) and 4m+4 of unit m = lane * S + j, walking down (exact in double)
    double cx = (double)(4 * (c.lane * S + S - 1) + 2), cy = cx + 2.0;
#pragma unroll 8
    for (int j = S - 1; j >= 0; --j) {
        const double2 cur = c.buf[base + j];
        acc.x = fma(cx, cur.y, acc.x);
        acc.y = fma(cy, next_x, acc.y);
        c.buf[base + j] = acc;
        next_x = cur.x;
        cx -= 4.0;
        cy -= 4.0;
    }
    double2 v = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double2 t = shfl_dn2(v, d);
        if (c.lane + d < 32) v = v + t;
    }
    double2 carry = shfl_dn2(v, 1);
    if (c.lane == 31) carry = d2(0.0, 0.0);
    const double f0 = I.f0;
#pragma unroll 1
    for (int j0 = 0; j0 < S; j0 += PB) {
        double2 t[PB];
#pragma unroll
        for (int e = 0; e < PB; ++e)
            if (j0 + e < S) t[e] = c.buf[base + j0 + e];
#pragma unroll
        for (int e = 0; e < PB; ++e) {
            if (j0 + e < S) {
                double2 u = t[e] + carry;
                if (j0 + e == 0 && c.lane == 0) u.x *= 0.5;
                c.buf[base + j0 + e] = d2(u.x * f0, u.y * f0);
            }
        }
    }
    __syncwarp();
}

// table k of a recurrence: shared-memory slot, or (slot < 0) a global pointer, per sequence with PERSEQ
struct RecTab {
    const double2 *sm;     // shared (or null)
    const double2 *gl;     // global (or null)
};
template <class Ctx>
__device__ __forceinline__ RecTab rec_tab(const Ctx &c, const pde_pass_ins &I, int k)
{
    RecTab t{nullptr, nullptr};
    if (I.slot[k] >= 0) {
        t.sm = c.tab + (long)I.slot[k] * Ctx::NUP;
    } else if (I.p[k]) {
        const long sq = ((I.flags & PDE_PASS_F_PERSEQ) && c.live) ? c.q : 0;
        t.gl = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(I.p[k]) + sq * I.ld[k]);
    }
    return t;
}

// ------------------------------------------------------------------------------------------------
// REC1: y[m] = T0[m] b[m] - T1[m] y[m -+ 1]  (ascending, or descending with flag DESC), in place.
// Tables in segment order: entry of (lane, j) at [j * 32 + lane]; T0 may be absent (= 1).
// Each walk advances in batches of RB steps whose table entries and data are read into registers first
// (in the result walk the stores of a step would otherwise fence the loads of the next one: 180 cycles per
// step in the first version).  GLOBAL = tables streamed from global memory: the next batch is fetched while
// the current one runs.
// ------------------------------------------------------------------------------------------------
template <bool GLOBAL> struct RecBatch { static constexpr int value = GLOBAL ? 4 : 8; };

template <bool GLOBAL, int NTAB, class Ctx>
struct RecFeed {
    static constexpr int RB = RecBatch<GLOBAL>::value;
    const double2 *T[NTAB];
    bool has[NTAB];
    double2 nxt[GLOBAL ? RB : 1][NTAB];
    int lane;
    bool desc;
    __device__ __forceinline__ double2 ld(int k, int jj) const
    {
        const int j = desc ? Ctx::SEGU - 1 - jj : jj;
        return GLOBAL ? __ldg(T[k] + j * 32 + lane) : T[k][j * 32 + lane];
    }
    // tables of the batch starting at step j0 -> cur; GLOBAL: also issue the loads of the following batch
    __device__ __forceinline__ void fetch(int j0, double2 (&cur)[RB][NTAB])
    {
#pragma unroll
        for (int e = 0; e < RB; ++e)
#pragma unroll
            for (int k = 0; k < NTAB; ++k) {
                if (GLOBAL) {
                    cur[e][k] = nxt[e][k];
                    if (j0 + RB + e < Ctx::SEGU) nxt[e][k] = ld(k, j0 + RB + e);
                } else if (j0 + e < Ctx::SEGU) {
                    cur[e][k] = ld(k, j0 + e);
                }
            }
    }
    __device__ __forceinline__ void prime()
    {
        if (GLOBAL) {
#pragma unroll
            for (int e = 0; e < RB; ++e)
#pragma unroll
                for (int k = 0; k < NTAB; ++k)
                    if (e < Ctx::SEGU) nxt[e][k] = ld(k, e);
        }
    }
};

template <bool GLOBAL, bool DESC, bool HAS0, class Ctx>
__device__ __forceinline__ void rec1_impl(const Ctx &c, const pde_pass_ins &I)
{
    constexpr int RB = RecBatch<GLOBAL>::value;
    const RecTab r0 = rec_tab(c, I, 0), r1 = rec_tab(c, I, 1);
    RecFeed<GLOBAL, 2, Ctx> feed;
    feed.T[1] = GLOBAL ? r1.gl : r1.sm;
    feed.T[0] = HAS0 ? (GLOBAL ? r0.gl : r0.sm) : feed.T[1];      // unconditional loads; unused without T0
    feed.lane = c.lane;
    feed.desc = DESC;
    double2 *seg = c.buf + c.segbase();
    constexpr int S = Ctx::SEGU;
    const double2 one = d2(1.0, 1.0);
    double2 A = one, B = d2(0.0, 0.0);
    // walk 1: the affine map of the segment
    feed.prime();
#pragma unroll 1
    for (int j0 = 0; j0 < S; j0 += RB) {
        double2 t[RB][2], b[RB];
        feed.fetch(j0, t);
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) b[e] = seg[DESC ? S - 1 - (j0 + e) : j0 + e];
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) {
                B = nfma2(t[e][1], B, HAS0 ? t[e][0] * b[e] : b[e]);
                A = -(t[e][1] * A);
            }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double2 Ap = DESC ? shfl_dn2(A, d) : shfl_up2(A, d);
        const double2 Bp = DESC ? shfl_dn2(B, d) : shfl_up2(B, d);
        const bool ok = DESC ? (c.lane + d < 32) : (c.lane >= d);
        B = ok ? fma2(A, Bp, B) : B;
        A = ok ? A * Ap : A;
    }
    double2 y = DESC ? shfl_dn2(B, 1) : shfl_up2(B, 1);
    if (DESC ? (c.lane == 31) : (c.lane == 0)) y = d2(0.0, 0.0);
    // walk 2: the recurrence from the true incoming state
    feed.prime();
#pragma unroll 1
    for (int j0 = 0; j0 < S; j0 += RB) {
        double2 t[RB][2], b[RB];
        feed.fetch(j0, t);
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) b[e] = seg[DESC ? S - 1 - (j0 + e) : j0 + e];
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) {
                y = nfma2(t[e][1], y, HAS0 ? t[e][0] * b[e] : b[e]);
                seg[DESC ? S - 1 - (j0 + e) : j0 + e] = y;
            }
    }
    __syncwarp();
}

template <bool GLOBAL, class Ctx>
__device__ __forceinline__ void op_rec1(const Ctx &c, const pde_pass_ins &I)
{
    const bool desc = (I.flags & PDE_PASS_F_DESC) != 0;
    const bool has0 = I.slot[0] >= 0 || I.p[0] != nullptr;
    if (desc) {
        if (has0) rec1_impl<GLOBAL, true, true, Ctx>(c, I);
        else rec1_impl<GLOBAL, true, false, Ctx>(c, I);
    } else {
        if (has0) rec1_impl<GLOBAL, false, true, Ctx>(c, I);
        else rec1_impl<GLOBAL, false, false, Ctx>(c, I);
    }
}

// ------------------------------------------------------------------------------------------------
// REC2 (descending): x[m] = T0[m] b[m] - T1[m] x[m+1] - T2[m] x[m+2], in place (back substitution of the
// 4-diagonal systems, fdma.f90:26-36, with reciprocal-scaled tables).
// ------------------------------------------------------------------------------------------------
template <bool GLOBAL, class Ctx>
__device__ __forceinline__ void op_rec2(const Ctx &c, const pde_pass_ins &I)
{
    constexpr int RB = RecBatch<GLOBAL>::value;
    const RecTab r0 = rec_tab(c, I, 0), r1 = rec_tab(c, I, 1), r2 = rec_tab(c, I, 2);
    RecFeed<GLOBAL, 3, Ctx> feed;
    feed.T[0] = GLOBAL ? r0.gl : r0.sm;
    feed.T[1] = GLOBAL ? r1.gl : r1.sm;
    feed.T[2] = GLOBAL ? r2.gl : r2.sm;
    feed.lane = c.lane;
    feed.desc = true;
    double2 *seg = c.buf + c.segbase();
    constexpr int S = Ctx::SEGU;
    const double2 one = d2(1.0, 1.0), zero = d2(0.0, 0.0);
    // state (a, b) = (x[m+1], x[m+2]); after the segment: (a, b)_out = H (a, b)_in + p
    double2 h00 = one, h01 = zero, h10 = zero, h11 = one, p0 = zero, p1 = zero;
    feed.prime();
#pragma unroll 1
    for (int j0 = 0; j0 < S; j0 += RB) {
        double2 t[RB][3], b[RB];
        feed.fetch(j0, t);
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) b[e] = seg[S - 1 - (j0 + e)];
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) {
                const double2 c1 = t[e][1], c2 = t[e][2];
                const double2 np = nfma2(c2, p1, nfma2(c1, p0, t[e][0] * b[e]));
                const double2 n0 = nfma2(c2, h10, -(c1 * h00));
                const double2 n1 = nfma2(c2, h11, -(c1 * h01));
                p1 = p0;
                p0 = np;
                h10 = h00;
                h11 = h01;
                h00 = n0;
                h01 = n1;
            }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double2 g00 = shfl_dn2(h00, d), g01 = shfl_dn2(h01, d), g10 = shfl_dn2(h10, d), g11 = shfl_dn2(h11, d);
        const double2 s0 = shfl_dn2(p0, d), s1 = shfl_dn2(p1, d);
        const bool ok = c.lane + d < 32;
        const double2 t0 = fma2(h00, s0, fma2(h01, s1, p0));
        const double2 t1 = fma2(h10, s0, fma2(h11, s1, p1));
        const double2 m00 = fma2(h00, g00, h01 * g10), m01 = fma2(h00, g01, h01 * g11);
        const double2 m10 = fma2(h10, g00, h11 * g10), m11 = fma2(h10, g01, h11 * g11);
        p0 = ok ? t0 : p0;
        p1 = ok ? t1 : p1;
        h00 = ok ? m00 : h00;
        h01 = ok ? m01 : h01;
        h10 = ok ? m10 : h10;
        h11 = ok ? m11 : h11;
    }
    double2 a = shfl_dn2(p0, 1), b2 = shfl_dn2(p1, 1);
    if (c.lane == 31) a = b2 = zero;
    feed.prime();
#pragma unroll 1
    for (int j0 = 0; j0 < S; j0 += RB) {
        double2 t[RB][3], b[RB];
        feed.fetch(j0, t);
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) b[e] = seg[S - 1 - (j0 + e)];
#pragma unroll
        for (int e = 0; e < RB; ++e)
            if (j0 + e < S) {
                const double2 x = nfma2(t[e][2], b2, nfma2(t[e][1], a, t[e][0] * b[e]));
                seg[S - 1 - (j0 + e)] = x;
                b2 = a;
                a = x;
            }
    }
    __syncwarp();
}

// (Measured and removed: prefetch.global.L2 of the row pieces that the later cooperative operators of a COL strip and
// the first operator of the CTA's next strip will read -- COL phases run in lock step, HBM idles while a strip
// computes -- changed nothing: 2.42 -> 2.46 ms of column passes per rbc2048 step.  The column passes are bound by
// the dependent chains of the per-warp recurrences with 8 warps per SM, not by the latency of their loads.)
// Short sequences are latency-bound per strip (a handful of dependent memory round trips, little arithmetic): two
// CTAs per SM (128 registers) instead of one (prof of the 256-member 128^2 ensemble: 25 us per strip with one).
template <bool COL, int LG>
__global__ void __launch_bounds__(LG <= 5 ? 256 : 64, LG <= 3 ? 2 : 1) k_pass(const pde_pass_job *__restrict__ jobs, int njobs, int ncache)
{
    extern __shared__ __align__(16) double2 pass_smem[];
    using Ctx = PassCtxT<LG, (LG <= 5 ? 8 : 2)>;
    constexpr int W = Ctx::W;
    Ctx c;
    c.lane = threadIdx.x & 31;
    c.w = threadIdx.x >> 5;
    c.all = pass_smem;
    c.buf = pass_smem + c.w * Ctx::BUFU;
    c.tab = pass_smem + W * Ctx::BUFU;
    pde_pass_ins *sprog = reinterpret_cast<pde_pass_ins *>(c.tab + PASS_SLOTS * Ctx::NUP);

    // job descriptors: one copy per CTA in shared memory (re-reading them from global memory cost ~10 % of the
    // first persistent version: four dependent L2 round trips per strip)
    pde_pass_job *sjobs = reinterpret_cast<pde_pass_job *>(sprog + PASS_MAX_INS);
    for (int t = threadIdx.x; t < ncache * (int)(sizeof(pde_pass_job) / 8); t += blockDim.x)
        reinterpret_cast<long long *>(sjobs)[t] = reinterpret_cast<const long long *>(jobs)[t];
    __syncthreads();
    auto jobref = [&](int j) -> const pde_pass_job & { return j < ncache ? sjobs[j] : jobs[j]; };

    // flattened (job, strip) list, split into contiguous chunks over the CTAs
    long total = 0;
    for (int j = 0; j < njobs; ++j) total += (jobref(j).nseq + W - 1) / W;
    // strips are dealt round-robin: at any moment the CTAs of the grid work on ADJACENT strips (COL: neighbouring
    // 64-byte pieces of the same rows -> whole DRAM pages; with contiguous chunks per CTA every page was shared by
    // ~4 CTAs only)
    int job = 0;
    long jstart = 0;
    int cur = -1, nins = 0, jstrips = (jobref(0).nseq + W - 1) / W;
    for (long idx = blockIdx.x; idx < total; idx += gridDim.x) {
        while (idx >= jstart + jstrips) {
            jstart += jstrips;
            ++job;
            jstrips = (jobref(job).nseq + W - 1) / W;
        }
        const bool first = job != cur;
        if (first) {
            cur = job;
            const pde_pass_job jb = jobref(job);
            nins = jb.nins;
            c.nseq = jb.nseq;
            c.seq0 = jb.seq0;
            __syncthreads();
            const long long *src = reinterpret_cast<const long long *>(jb.prog);
            long long *dst = reinterpret_cast<long long *>(sprog);
            const int words = nins * (int)(sizeof(pde_pass_ins) / 8);
            for (int t = threadIdx.x; t < words; t += blockDim.x) dst[t] = src[t];
            __syncthreads();
        }
        c.q0 = (int)(idx - jstart) * W;
        c.q = c.q0 + c.w;
        c.live = c.q < c.nseq;
        // ROW: a warp only ever touches its own sequence -> the warps of a CTA run free (no CTA barrier; one warp's
        // memory latency overlaps another warp's recurrences).  COL: the data-movement operators are cooperative
        // (thread <-> (row, column) of the strip): barrier when the program switches between them and the per-warp
        // operators.
        bool prev_coop = true;
        for (int ip = 0; ip < nins; ++ip) {
            const pde_pass_ins &I = sprog[ip];
            const int op = I.op;
            if (COL) {
                const bool coop = op == PDE_PASS_LOAD || op == PDE_PASS_STORE || op == PDE_PASS_AXPY || op == PDE_PASS_LINCOMB;
                if (op != PDE_PASS_TABLES && coop != prev_coop) {
                    __syncthreads();
                    prev_coop = coop;
                }
            }
            switch (op) {
            case PDE_PASS_TABLES:
                if (first) {
                    for (int k = 0; k < I.n; ++k) {
                        const double2 *src = reinterpret_cast<const double2 *>(I.p[k]);
                        double2 *dst = c.tab + (long)k * Ctx::NUP;      // slots are positional
                        for (int m = threadIdx.x; m < Ctx::NUP; m += blockDim.x) pcp16(dst + m, src + m, 16);
                    }
                    pcp_wait_all();
                    __syncthreads();
                }
                break;
            case PDE_PASS_LOAD: op_load<COL, Ctx>(c, I); break;
            case PDE_PASS_STORE: op_store<COL, Ctx>(c, I); break;
            case PDE_PASS_AXPY: op_axpy<COL, Ctx>(c, I); break;
            case PDE_PASS_LINCOMB: op_lincomb<COL, Ctx>(c, I); break;
            case PDE_PASS_SCALE: {
                const double f0 = I.f0;
                double2 *bb = c.buf + c.segbase();
#pragma unroll 8
                for (int j = 0; j < Ctx::SEGU; ++j) bb[j] = d2(bb[j].x * f0, bb[j].y * f0);
                __syncwarp();
                break;
            }
            case PDE_PASS_SETZ0:
                if (c.live && c.lane == 0 && c.seq0 + c.q == I.off[0]) c.buf[0].x = 0.0;
                __syncwarp();
                break;
            case PDE_PASS_POINT: op_point<Ctx>(c, I); break;
            case PDE_PASS_DIFF: op_diff(c, I); break;
            case PDE_PASS_REC1:
                if (I.slot[1] >= 0) op_rec1<false, Ctx>(c, I);
                else op_rec1<true, Ctx>(c, I);
                break;
            case PDE_PASS_REC2:
                if (I.slot[0] >= 0) op_rec2<false, Ctx>(c, I);
                else op_rec2<true, Ctx>(c, I);
                break;
            default: break;
            }
        }
        if (COL && !prev_coop) __syncthreads();      // the next strip starts with cooperative loads
    }
}

// ------------------------------------------------------------------------------------------------
// device-side barrier between the ranks of one node (slab decomposition): every rank bumps its slot in
// every peer's flag array and waits until all of its own slots reached the new epoch.
// ------------------------------------------------------------------------------------------------
__global__ void k_peer_barrier(unsigned long long *const *__restrict__ peer_flags, unsigned long long *epoch_ctr,
                               int rank, int nranks, int *err)
{
    const int s = threadIdx.x;
    __shared__ unsigned long long epoch;
    if (s == 0) epoch = *epoch_ctr + 1;
    __syncthreads();
    if (s < nranks) {
        __threadfence_system();
        volatile unsigned long long *theirs = peer_flags[s] + rank;
        *theirs = epoch;
        __threadfence_system();
        volatile unsigned long long *mine = peer_flags[rank] + s;
        const long long t0 = clock64();
        while (*mine < epoch) {
            if (clock64() - t0 > 8000000000LL) {        // ~4 s: a rank is gone -- do not hang the GPU
                if (err) *err = 1;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (s == 0) *epoch_ctr = epoch;
}

static int pass_width_for(int lg) { return lg <= 5 ? 8 : 2; }

static int pass_ncache(int lg, int njobs) { return njobs < pass_jobs_cached(lg) ? njobs : pass_jobs_cached(lg); }

static size_t pass_smem_bytes(int lg, int njobs)
{
    const int W = pass_width_for(lg);
    const size_t nup = (size_t)32 << lg, bufu = nup + (lg > 0 ? 32 : 0) + 1;
    return (W * bufu + PASS_SLOTS * nup) * sizeof(double2) + PASS_MAX_INS * sizeof(pde_pass_ins) +
           (size_t)pass_ncache(lg, njobs) * sizeof(pde_pass_job);
}

template <bool COL, int LG>
static int launch_pass(int njobs, int max_nseq, const pde_pass_job *dev_jobs, cudaStream_t st)
{
    constexpr int W = LG <= 5 ? 8 : 2;
    const size_t smem = pass_smem_bytes(LG, njobs);
    static PerDeviceSize attr;
    if (smem > 48 * 1024 && smem > attr.get()) {
        PDE_CUDA(cudaFuncSetAttribute(k_pass<COL, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.get() = smem;
    }
    // persistent CTAs: as many as are resident at once, never more than there are strips
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    while (per_sm > 1 && per_sm * 32 * W > 2048) --per_sm;
    if (per_sm > (LG <= 3 ? 2 : 1)) per_sm = LG <= 3 ? 2 : 1;        // register budget of the kernel (__launch_bounds__)
    long strips = (long)njobs * ceil_div(max_nseq, W);
    long grid = (long)sm_count() * per_sm;
    if (grid > strips) grid = strips;
    k_pass<COL, LG><<<(unsigned)grid, 32 * W, smem, st>>>(dev_jobs, njobs, pass_ncache(LG, njobs));
    return after_launch("pde_pass_run");
}

template <bool COL>
static int launch_pass_lg(int lg, int njobs, int max_nseq, const pde_pass_job *dev_jobs, cudaStream_t st)
{
    switch (lg) {
    case 0: return launch_pass<COL, 0>(njobs, max_nseq, dev_jobs, st);
    case 1: return launch_pass<COL, 1>(njobs, max_nseq, dev_jobs, st);
    case 2: return launch_pass<COL, 2>(njobs, max_nseq, dev_jobs, st);
    case 3: return launch_pass<COL, 3>(njobs, max_nseq, dev_jobs, st);
    case 4: return launch_pass<COL, 4>(njobs, max_nseq, dev_jobs, st);
    case 5: return launch_pass<COL, 5>(njobs, max_nseq, dev_jobs, st);
    default: return launch_pass<COL, 6>(njobs, max_nseq, dev_jobs, st);
    }
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_pass_width(int lg_segu) { return pass_width_for(lg_segu); }

int pde_pass_run(int layout, int lg_segu, int njobs, int max_nseq, const pde_pass_job *dev_jobs, void *stream)
{
    PDE_REQUIRE(layout == PDE_PASS_ROW || layout == PDE_PASS_COL, "layout");
    PDE_REQUIRE(lg_segu >= 0 && lg_segu <= 6, "0 <= log2(units per lane) <= 6 (sequences up to 4096)");
    PDE_REQUIRE(dev_jobs != nullptr, "null job list");
    if (njobs <= 0 || max_nseq <= 0) return PDE_OK;
    if (layout == PDE_PASS_ROW) return launch_pass_lg<false>(lg_segu, njobs, max_nseq, dev_jobs, as_stream(stream));
    return launch_pass_lg<true>(lg_segu, njobs, max_nseq, dev_jobs, as_stream(stream));
}

int pde_ipc_alloc(void **ptr, long bytes, void *handle64)
{
    PDE_REQUIRE(ptr && handle64 && bytes > 0, "arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PDE_CUDA(cudaMalloc(ptr, (size_t)bytes));
    PDE_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    PDE_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), *ptr));
    return PDE_OK;
}

int pde_ipc_open(const void *handle64, void **ptr)
{
    PDE_REQUIRE(ptr && handle64, "arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    PDE_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PDE_OK;
}

int pde_ipc_close(void *ptr)
{
    PDE_CUDA(cudaIpcCloseMemHandle(ptr));
    return PDE_OK;
}

int pde_ipc_free(void *ptr)
{
    PDE_CUDA(cudaFree(ptr));
    return PDE_OK;
}

int pde_peer_barrier(void *const *dev_peer_flags, void *dev_epoch, int rank, int nranks, int *dev_err, void *stream)
{
    PDE_REQUIRE(dev_peer_flags && dev_epoch && nranks >= 1 && nranks <= 32 && rank >= 0 && rank < nranks, "arguments");
    k_peer_barrier<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long *const *>(dev_peer_flags),
                                                    reinterpret_cast<unsigned long long *>(dev_epoch), rank, nranks, dev_err);
    return after_launch("pde_peer_barrier");
}

}  // extern "C"
