// Fused axis passes (sm_100a): one launch applies a whole chain of 1-D operators of the
// Chebyshev-Galerkin time step to every sequence of several 2-D arrays.
//
// SURVEY.md §8(d) describes an IMEX stage as ~9 axis passes: every operator of the step acts along ONE
// axis (stencil maps chebyshev.py:287-337, derivative recurrence differentiate_cheby.f90:28-53, banded
// products plans.py:54-74, offset-2 Thomas sweeps tdma.f90:55-106 / fdma.f90:1-98, per-column Poisson
// solves fdma.f90:146-195).  Round 1 ran each of them as its own kernel (39 launches per stage, every
// intermediate through HBM, recurrences as ONE thread per chain).  Here a CTA keeps PASS_W sequences
// (rows: ROW layout, or a strip of adjacent columns: COL layout) in shared memory, and the warp that owns
// a sequence interprets a small program of operators on it:
//
//   LOAD / STORE / AXPY   move the sequence between HBM and shared memory (cp.async, coalesced; a
//                         sequence may be split over several base pointers = the slabs of peer GPUs:
//                         the distributed transposes of the slab decomposition are the loads and
//                         stores of the row passes, straight over NVLink peer mappings);
//   POINT                 banded product / stencil map (taps at unit offsets -1..+2);
//   DIFF                  derivative recurrence (suffix sums);
//   REC1 / REC2           first / second order linear recurrences (the Thomas sweeps), CHAIN-SPLIT:
//                         every lane owns 1/32 of the chain, composes the affine map of its segment
//                         (2x2 for the two-term back substitution), the maps are combined with a
//                         warp scan, and a second walk over the segment writes the result.
//
// A sequence is stored as 16-byte units (x[2m], x[2m+1]): the two parity chains of the offset-2
// recurrences ride in the two halves of a double2, so one thread always advances two independent
// chains.  Lane l owns units [l*SEGU, (l+1)*SEGU); unit m lives at m + (m >> lg) (one pad unit per
// segment: lane-strided and consecutive accesses are both conflict-free).  Recurrence tables are
// prepared by the host in the matching "segment-transposed" order [j][lane] so that the 32 lanes read
// 512 contiguous bytes.  Arithmetic differs from the reference by rounding only (re-association across
// segments, reciprocal-scaled tables): the contract is 1e-12, checked against the oracle in
// tests/test_gpu_pass.py.
#include "common.cuh"
#include <cstring>

namespace pde {

constexpr int PASS_W = 4;                    // sequences (warps) per CTA

__device__ __forceinline__ void pcp16(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void pcp8(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void pcp_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double2 d2(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ double2 operator*(double2 a, double2 b) { return d2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return d2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 operator-(double2 a) { return d2(-a.x, -a.y); }
// a*b + c, a*b - c*d helpers (component-wise; the compiler contracts to DFMA)
__device__ __forceinline__ double2 fma2(double2 a, double2 b, double2 c) { return d2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ double2 nfma2(double2 a, double2 b, double2 c) { return d2(fma(-a.x, b.x, c.x), fma(-a.y, b.y, c.y)); }
__device__ __forceinline__ double2 shfl_up2(double2 v, int d)
{
    return d2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ double2 shfl_dn2(double2 v, int d)
{
    return d2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}

struct PassCtx {
    int lg, SEGU, NUP, BUFU;       // log2 units per lane segment, units per segment, padded units per sequence, pitch
    int lane, w, q0, q;            // lane, warp = sequence within the CTA, first / own sequence of the job
    bool live;                     // q < nseq
    int nseq, seq0;
    double2 *buf;                  // this warp's sequence
    double2 *all;                  // buffer 0 of the CTA
    __device__ __forceinline__ int pu(int m) const { return m + ((m >> lg) & -(lg > 0)); }
};

// segment of element i (ROW: along the sequence) -- nseg is small
__device__ __forceinline__ int seg_of(const pde_pass_ins &I, int i)
{
    int s = 0;
    while (s + 1 < I.nseg && i >= I.start[s + 1]) ++s;
    return s;
}

// address of element i of sequence q
template <bool COL>
__device__ __forceinline__ const double *elem_addr(const pde_pass_ins &I, int s, int q, int i)
{
    const double *base = reinterpret_cast<const double *>(I.p[s]);
    if (COL) return base + (long)(i - I.start[s]) * I.ld[s] + q;
    return base + (long)q * I.ld[s] + (i - I.start[s]);
}

// ------------------------------------------------------------------------------------------------
// LOAD: buffer <- elements [0, n) of the operand, zero beyond
// ------------------------------------------------------------------------------------------------
template <bool COL>
__device__ __forceinline__ void op_load(const PassCtx &c, const pde_pass_ins &I)
{
    const int n = I.n;
    if (!COL) {
        for (int m = c.lane; m < c.NUP; m += 32) {
            const int i = 2 * m;
            int bytes = 0;
            const double *src = reinterpret_cast<const double *>(I.p[0]);
            if (c.live && i < n) {
                const int s = seg_of(I, i);
                bytes = (i + 1 < n) ? 16 : 8;
                src = elem_addr<false>(I, s, c.q, i);
            }
            pcp16(&c.buf[c.pu(m)], src, bytes);
        }
    } else {
        // CTA-cooperative: thread (r0, col) walks rows r0, r0 + 32, ... of column q0 + col
        const int col = threadIdx.x & (PASS_W - 1), r0 = threadIdx.x / PASS_W;
        double *dst = reinterpret_cast<double *>(c.all + col * c.BUFU);
        const bool ok = c.q0 + col < c.nseq;
        for (int r = r0; r < 2 * c.NUP; r += 32) {
            int bytes = 0;
            const double *src = reinterpret_cast<const double *>(I.p[0]);
            if (ok && r < n) {
                const int s = seg_of(I, r);
                bytes = 8;
                src = elem_addr<true>(I, s, c.q0 + col, r);
            }
            pcp8(dst + 2 * c.pu(r >> 1) + (r & 1), src, bytes);
        }
    }
    pcp_wait_all();
}

// ------------------------------------------------------------------------------------------------
// STORE: elements [0, n) of the buffer -> operand.  flag ONLY_SEQ: only the sequence with global index off[0]
// ------------------------------------------------------------------------------------------------
template <bool COL>
__device__ __forceinline__ void op_store(const PassCtx &c, const pde_pass_ins &I)
{
    const int n = I.n;
    const bool only = (I.flags & PDE_PASS_F_ONLY_SEQ) != 0;
    if (!COL) {
        if (!c.live || (only && c.seq0 + c.q != I.off[0])) return;
        for (int m = c.lane; 2 * m < n; m += 32) {
            const int i = 2 * m;
            const int s = seg_of(I, i);
            double *dst = const_cast<double *>(elem_addr<false>(I, s, c.q, i));
            const double2 v = c.buf[c.pu(m)];
            if (i + 1 < n) *reinterpret_cast<double2 *>(dst) = v;
            else *dst = v.x;
        }
    } else {
        const int col = threadIdx.x & (PASS_W - 1), r0 = threadIdx.x / PASS_W;
        const double *srcb = reinterpret_cast<const double *>(c.all + col * c.BUFU);
        if (c.q0 + col >= c.nseq || (only && c.seq0 + c.q0 + col != I.off[0])) return;
        for (int r = r0; r < n; r += 32) {
            const int s = seg_of(I, r);
            double *dst = const_cast<double *>(elem_addr<true>(I, s, c.q0 + col, r));
            *dst = srcb[2 * c.pu(r >> 1) + (r & 1)];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// AXPY: buffer <- f1 * buffer + f0 * G   (f1 applied only with flag SCALED), G = elements [0, n) of the
// operand; flag STENCIL: G_i = g_i + st_i g_{i-2} with the element table st = p[7] (st_i = s_{i-2}, st_0 = st_1 = 0)
// ------------------------------------------------------------------------------------------------
template <bool COL>
__device__ __forceinline__ void op_axpy(const PassCtx &c, const pde_pass_ins &I)
{
    const int n = I.n;
    const double f0 = I.f0, f1 = (I.flags & PDE_PASS_F_SCALED) ? I.f1 : 1.0;
    const bool scaled = (I.flags & PDE_PASS_F_SCALED) != 0;
    const bool sten = (I.flags & PDE_PASS_F_STENCIL) != 0;
    const double *st = reinterpret_cast<const double *>(I.p[PDE_PASS_MAX_SEG - 1]);
    if (!COL) {
        if (!c.live) return;
        for (int m = c.lane; m < c.NUP; m += 32) {
            const int i = 2 * m;
            double2 g = d2(0.0, 0.0);
            if (i < n) {
                const int s = seg_of(I, i);
                const double *src = elem_addr<false>(I, s, c.q, i);
                if (i + 1 < n) g = __ldg(reinterpret_cast<const double2 *>(src));
                else g.x = __ldg(src);
                if (sten && i >= 2) {
                    const int s2 = seg_of(I, i - 2);
                    const double2 h = __ldg(reinterpret_cast<const double2 *>(elem_addr<false>(I, s2, c.q, i - 2)));
                    const double2 t = __ldg(reinterpret_cast<const double2 *>(st + i));
                    g.x = fma(t.x, h.x, g.x);
                    if (i + 1 < n) g.y = fma(t.y, h.y, g.y);
                }
            } else if (!scaled) {
                continue;
            }
            double2 &b = c.buf[c.pu(m)];
            b = scaled ? d2(fma(f0, g.x, f1 * b.x), fma(f0, g.y, f1 * b.y)) : d2(fma(f0, g.x, b.x), fma(f0, g.y, b.y));
        }
    } else {
        const int col = threadIdx.x & (PASS_W - 1), r0 = threadIdx.x / PASS_W;
        double *dstb = reinterpret_cast<double *>(c.all + col * c.BUFU);
        if (c.q0 + col >= c.nseq) return;
        const int lim = scaled ? 2 * c.NUP : n;
        for (int r = r0; r < lim; r += 32) {
            double g = 0.0;
            if (r < n) {
                const int s = seg_of(I, r);
                g = __ldg(elem_addr<true>(I, s, c.q0 + col, r));
                if (sten && r >= 2) {
                    const int s2 = seg_of(I, r - 2);
                    g = fma(__ldg(st + r), __ldg(elem_addr<true>(I, s2, c.q0 + col, r - 2)), g);
                }
            }
            double &b = dstb[2 * c.pu(r >> 1) + (r & 1)];
            b = scaled ? fma(f0, g, f1 * b) : fma(f0, g, b);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// POINT: y[m] = sum_t C_t[m] * x[m + off_t] (units), in place.  All offsets >= 0 (ascending walk) or all <= 0
// (descending walk).  C_t = p[t] (double2 per unit, NUP entries) or 1 when p[t] is null.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void op_point(const PassCtx &c, const pde_pass_ins &I)
{
    const int nt = I.n;
    bool desc = false;
    for (int t = 0; t < nt; ++t) desc |= I.off[t] < 0;
    constexpr int G = 4;                                 // chunks of 32 units per synchronisation
    const int ngroups = (c.NUP + 32 * G - 1) / (32 * G);
    for (int gi = 0; gi < ngroups; ++gi) {
        const int g = desc ? ngroups - 1 - gi : gi;
        double2 acc[G];
#pragma unroll
        for (int e = 0; e < G; ++e) {
            const int m = (g * G + e) * 32 + c.lane;
            acc[e] = d2(0.0, 0.0);
            if (m < c.NUP) {
                for (int t = 0; t < nt; ++t) {
                    const int mm = m + I.off[t];
                    double2 x = d2(0.0, 0.0);
                    if (mm >= 0 && mm < c.NUP) x = c.buf[c.pu(mm)];
                    const double2 *tab = reinterpret_cast<const double2 *>(I.p[t]);
                    if (tab) acc[e] = fma2(__ldg(tab + m), x, acc[e]);
                    else acc[e] = acc[e] + x;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < G; ++e) {
            const int m = (g * G + e) * 32 + c.lane;
            if (m < c.NUP) c.buf[c.pu(m)] = acc[e];
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// DIFF: dc_k = dc_{k+2} + 2 (k+1) c_{k+1}, dc_0 = dc_2 / 2 + c_1, result times f0 (1 / scale)
// (differentiate_cheby.f90:28-53).  Units: D[m] = D[m+1] + (2(2m+1) X[m].y, (4m+4) X[m+1].x).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void op_diff(const PassCtx &c, const pde_pass_ins &I)
{
    const int base = c.lane * (c.SEGU + (c.lg > 0));
    // first element of the next lane's segment (original value), before anybody writes
    double next_x = 0.0;
    if (c.lane < 31) next_x = c.buf[base + c.SEGU + (c.lg > 0)].x;
    __syncwarp();
    double2 acc = d2(0.0, 0.0);
#pragma unroll 4
    for (int j = c.SEGU - 1; j >= 0; --j) {
        const int m = c.lane * c.SEGU + j;
        const double2 cur = c.buf[base + j];
        acc.x = fma((double)(4 * m + 2), cur.y, acc.x);
        acc.y = fma((double)(4 * m + 4), next_x, acc.y);
        c.buf[base + j] = acc;
        next_x = cur.x;
    }
    // exclusive suffix sum of the segment totals over the lanes
    double2 v = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double2 t = shfl_dn2(v, d);
        if (c.lane + d < 32) v = v + t;
    }
    double2 carry = shfl_dn2(v, 1);
    if (c.lane == 31) carry = d2(0.0, 0.0);
    const double f0 = I.f0;
#pragma unroll 4
    for (int j = c.SEGU - 1; j >= 0; --j) {
        double2 t = c.buf[base + j] + carry;
        if (c.lane == 0 && j == 0) t.x *= 0.5;
        c.buf[base + j] = d2(t.x * f0, t.y * f0);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// REC1: y[m] = T0[m] b[m] - T1[m] y[m -+ 1]  (ascending, or descending with flag DESC), in place.
// Tables in segment-transposed order: entry of (lane, j) at [j * 32 + lane]; T0 may be null (= 1).
// flag PERSEQ: tables of sequence q start at p[t] + q * ld[t] doubles.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void op_rec1(const PassCtx &c, const pde_pass_ins &I)
{
    const bool desc = (I.flags & PDE_PASS_F_DESC) != 0;
    const bool perseq = (I.flags & PDE_PASS_F_PERSEQ) != 0;
    const long sq = (perseq && c.live) ? c.q : 0;
    const double2 *T0 = reinterpret_cast<const double2 *>(I.p[0] ? reinterpret_cast<const double *>(I.p[0]) + sq * I.ld[0] : nullptr);
    const double2 *T1 = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(I.p[1]) + sq * I.ld[1]);
    const int base = c.lane * (c.SEGU + (c.lg > 0));
    const int S = c.SEGU;
    const double2 one = d2(1.0, 1.0);
    // walk 1: affine map of the segment, x_out = A x_in + B
    double2 A = one, B = d2(0.0, 0.0);
#pragma unroll 4
    for (int jj = 0; jj < S; ++jj) {
        const int j = desc ? S - 1 - jj : jj;
        const double2 c1 = __ldg(T1 + j * 32 + c.lane);
        const double2 c0 = T0 ? __ldg(T0 + j * 32 + c.lane) : one;
        const double2 b = c.buf[base + j];
        B = nfma2(c1, B, c0 * b);
        A = -(c1 * A);
    }
    // inclusive scan of the maps in chain direction
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double2 Ap = desc ? shfl_dn2(A, d) : shfl_up2(A, d);
        const double2 Bp = desc ? shfl_dn2(B, d) : shfl_up2(B, d);
        const bool ok = desc ? (c.lane + d < 32) : (c.lane >= d);
        if (ok) {
            B = fma2(A, Bp, B);
            A = A * Ap;
        }
    }
    double2 y = desc ? shfl_dn2(B, 1) : shfl_up2(B, 1);
    if (desc ? (c.lane == 31) : (c.lane == 0)) y = d2(0.0, 0.0);
    // walk 2: the recurrence from the true incoming state
#pragma unroll 4
    for (int jj = 0; jj < S; ++jj) {
        const int j = desc ? S - 1 - jj : jj;
        const double2 c1 = __ldg(T1 + j * 32 + c.lane);
        const double2 c0 = T0 ? __ldg(T0 + j * 32 + c.lane) : one;
        const double2 b = c.buf[base + j];
        y = nfma2(c1, y, c0 * b);
        c.buf[base + j] = y;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// REC2 (descending): x[m] = T0[m] b[m] - T1[m] x[m+1] - T2[m] x[m+2], in place (back substitution of the
// 4-diagonal systems, fdma.f90:26-36, with reciprocal-scaled tables).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void op_rec2(const PassCtx &c, const pde_pass_ins &I)
{
    const bool perseq = (I.flags & PDE_PASS_F_PERSEQ) != 0;
    const long sq = (perseq && c.live) ? c.q : 0;
    const double2 *T0 = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(I.p[0]) + sq * I.ld[0]);
    const double2 *T1 = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(I.p[1]) + sq * I.ld[1]);
    const double2 *T2 = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(I.p[2]) + sq * I.ld[2]);
    const int base = c.lane * (c.SEGU + (c.lg > 0));
    const int S = c.SEGU;
    const double2 one = d2(1.0, 1.0), zero = d2(0.0, 0.0);
    // state (a, b) = (x[m+1], x[m+2]); after the segment: (a, b)_out = H (a, b)_in + p
    double2 h00 = one, h01 = zero, h10 = zero, h11 = one, p0 = zero, p1 = zero;
#pragma unroll 2
    for (int j = S - 1; j >= 0; --j) {
        const double2 c0 = __ldg(T0 + j * 32 + c.lane), c1 = __ldg(T1 + j * 32 + c.lane), c2 = __ldg(T2 + j * 32 + c.lane);
        const double2 b = c.buf[base + j];
        const double2 np = nfma2(c2, p1, nfma2(c1, p0, c0 * b));
        const double2 n0 = nfma2(c2, h10, -(c1 * h00));
        const double2 n1 = nfma2(c2, h11, -(c1 * h01));
        p1 = p0;
        p0 = np;
        h10 = h00;
        h11 = h01;
        h00 = n0;
        h01 = n1;
    }
    // inclusive suffix scan: this lane's map after the maps of the lanes above
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double2 g00 = shfl_dn2(h00, d), g01 = shfl_dn2(h01, d), g10 = shfl_dn2(h10, d), g11 = shfl_dn2(h11, d);
        const double2 r0 = shfl_dn2(p0, d), r1 = shfl_dn2(p1, d);
        if (c.lane + d < 32) {
            const double2 q0 = fma2(h00, r0, fma2(h01, r1, p0));
            const double2 q1 = fma2(h10, r0, fma2(h11, r1, p1));
            const double2 m00 = fma2(h00, g00, h01 * g10), m01 = fma2(h00, g01, h01 * g11);
            const double2 m10 = fma2(h10, g00, h11 * g10), m11 = fma2(h10, g01, h11 * g11);
            p0 = q0;
            p1 = q1;
            h00 = m00;
            h01 = m01;
            h10 = m10;
            h11 = m11;
        }
    }
    double2 a = shfl_dn2(p0, 1), b2 = shfl_dn2(p1, 1);
    if (c.lane == 31) a = b2 = zero;
#pragma unroll 2
    for (int j = S - 1; j >= 0; --j) {
        const double2 c0 = __ldg(T0 + j * 32 + c.lane), c1 = __ldg(T1 + j * 32 + c.lane), c2 = __ldg(T2 + j * 32 + c.lane);
        const double2 b = c.buf[base + j];
        const double2 x = nfma2(c2, b2, nfma2(c1, a, c0 * b));
        c.buf[base + j] = x;
        b2 = a;
        a = x;
    }
    __syncwarp();
}

template <bool COL>
__global__ void __launch_bounds__(32 * PASS_W) k_pass(const pde_pass_job *__restrict__ jobs, int lg)
{
    extern __shared__ __align__(16) double2 pass_smem[];
    const pde_pass_job job = jobs[blockIdx.y];
    PassCtx c;
    c.q0 = blockIdx.x * PASS_W;
    if (c.q0 >= job.nseq) return;
    c.lg = lg;
    c.SEGU = 1 << lg;
    c.NUP = 32 << lg;
    c.BUFU = c.NUP + (lg > 0 ? 32 : 0);
    c.lane = threadIdx.x & 31;
    c.w = threadIdx.x >> 5;
    c.q = c.q0 + c.w;
    c.nseq = job.nseq;
    c.seq0 = job.seq0;
    c.live = c.q < job.nseq;
    c.all = pass_smem;
    c.buf = pass_smem + c.w * c.BUFU;
    for (int ip = 0; ip < job.nins; ++ip) {
        const pde_pass_ins &I = job.prog[ip];
        switch (I.op) {
        case PDE_PASS_LOAD: op_load<COL>(c, I); break;
        case PDE_PASS_STORE: op_store<COL>(c, I); break;
        case PDE_PASS_AXPY: op_axpy<COL>(c, I); break;
        case PDE_PASS_SCALE: {
            const double f0 = I.f0;
            for (int m = c.lane; m < c.NUP; m += 32) {
                double2 &b = c.buf[c.pu(m)];
                b = d2(b.x * f0, b.y * f0);
            }
            break;
        }
        case PDE_PASS_SETZ0:
            if (c.live && c.lane == 0 && c.seq0 + c.q == I.off[0]) c.buf[0].x = 0.0;
            break;
        case PDE_PASS_POINT: op_point(c, I); break;
        case PDE_PASS_DIFF: op_diff(c, I); break;
        case PDE_PASS_REC1: op_rec1(c, I); break;
        case PDE_PASS_REC2: op_rec2(c, I); break;
        default: break;
        }
        __syncthreads();
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

}  // namespace pde

using namespace pde;

extern "C" {

int pde_pass_run(int layout, int lg_segu, int njobs, int max_nseq, const pde_pass_job *dev_jobs, void *stream)
{
    PDE_REQUIRE(layout == PDE_PASS_ROW || layout == PDE_PASS_COL, "layout");
    PDE_REQUIRE(lg_segu >= 0 && lg_segu <= 6, "0 <= log2(units per lane) <= 6 (sequences up to 4096)");
    PDE_REQUIRE(dev_jobs != nullptr, "null job list");
    if (njobs <= 0 || max_nseq <= 0) return PDE_OK;
    const int bufu = (32 << lg_segu) + (lg_segu > 0 ? 32 : 0);
    const size_t smem = (size_t)PASS_W * bufu * sizeof(double2);
    static PerDeviceSize attr_row, attr_col;
    size_t &have = (layout == PDE_PASS_ROW ? attr_row : attr_col).get();
    if (smem > 48 * 1024 && smem > have) {
        if (layout == PDE_PASS_ROW)
            PDE_CUDA(cudaFuncSetAttribute(k_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            PDE_CUDA(cudaFuncSetAttribute(k_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    dim3 grid(ceil_div(max_nseq, PASS_W), njobs);
    if (layout == PDE_PASS_ROW) k_pass<false><<<grid, 32 * PASS_W, smem, as_stream(stream)>>>(dev_jobs, lg_segu);
    else k_pass<true><<<grid, 32 * PASS_W, smem, as_stream(stream)>>>(dev_jobs, lg_segu);
    return after_launch("pde_pass_run");
}

int pde_pass_width(void) { return PASS_W; }

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
