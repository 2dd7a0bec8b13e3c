// Sequence-per-thread sweep kernels with a cp.async prefetch ring (sm_100a).
//
// Every sequential operator of the Chebyshev-Galerkin path (derivative recurrence,
// Thomas sweeps of the offset-2 tridiagonal and of the 4-diagonal systems) is one or
// two monotone passes over independent sequences with O(1) state.  One THREAD owns one
// sequence and walks it start to end (both parity chains interleaved -> ILP 2):
//   LC ("lane-coalesced", axis 0): element i of sequence q at base[i*ld + q]; a warp
//       touches one contiguous 256-byte row segment per step;
//   TS ("thread-sequential", axis 1): element i of sequence q at base[q*ld + i]; every
//       lane streams through its own row (32-byte sectors are reused over 4 steps via L1).
// The recurrences are latency chains, so the only way to keep HBM busy is memory-level
// parallelism: each thread keeps (SW_KC-1)*SW_C future elements of every input stream in
// flight with 8-byte cp.async copies into its private shared-memory ring (no registers,
// no barriers: a thread only ever reads what it copied itself); one group per chunk of
// SW_C steps so the chain steps of a chunk run back to back.  Several arrays (e.g.
// the U, V, T fields) are processed by one launch (blockIdx.y = job).
//
// Arithmetic is identical to the reference's Fortran, operation for operation (this
// header is only included from translation units compiled with --fmad=false).
#pragma once
#include "common.cuh"
#include <type_traits>
#include <cstdint>

namespace pde {

constexpr int SW_C = 8;          // steps per chunk (one cp.async group)
static_assert(true, "");
#ifndef PDE_SW_KC
#define PDE_SW_KC 4
#endif
constexpr int SW_KC = PDE_SW_KC;         // chunks in the ring (SW_KC-1 chunks = 24 steps in flight per thread)
constexpr int RING_K = SW_C * SW_KC;
// TS sweeps: every lane streams through its own row with 8-byte copies, so DRAM sees isolated 32-byte
// sectors of 32 different rows per warp step.  Each lane therefore asks L2 for the 128-byte window of a
// chunk SW_PF chunks ahead of its cp.async ring (one cp.async.bulk.prefetch.L2 per chunk and stream):
// DRAM is read in whole-line requests and the ring's copies hit L2.  0 = off.
#ifndef PDE_SW_PF
#define PDE_SW_PF 0
#endif
constexpr int SW_PF = PDE_SW_PF;
constexpr int SWEEP_MAX_JOBS = 8;
constexpr int SWEEP_MAX_IN = 5;

using SweepJob = pde_sweep_job;   // layout shared with the C ABI (include/pypde_b200.h)

struct SweepJobs {
    int njobs;
    int n;                     // sequence length (number of steps)
    SweepJob j[SWEEP_MAX_JOBS];
};

__device__ __forceinline__ void cp_async_8(double *smem, const double *gmem)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2_128(const double *gmem)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], 128;\n" ::"l"(gmem) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Correctly rounded a / d from the precomputed correctly rounded reciprocal rd = RN(1/d):
// q = a rd is within ~1 ulp, one FMA residual + correction makes it faithful to ~2^-105, the
// second one rounds it to RN(a/d) (Markstein's theorem).  5 dependent FP64 ops instead of the
// ~25-instruction generic division on the recurrence's critical path; same bits.
__device__ __forceinline__ double div_rn(double a, double d, const double *rdp, int i)
{
    if (rdp == nullptr) return a / d;
    const double rd = __ldg(rdp + i);
    double q = a * rd;
    double r = __fma_rn(-q, d, a);
    q = __fma_rn(r, rd, q);
    r = __fma_rn(-q, d, a);
    return __fma_rn(r, rd, q);
}
__device__ __forceinline__ double div_rn_v(double a, double d, double rd, bool have)
{
    if (!have) return a / d;
    double q = a * rd;
    double r = __fma_rn(-q, d, a);
    q = __fma_rn(r, rd, q);
    r = __fma_rn(-q, d, a);
    return __fma_rn(r, rd, q);
}

template <bool LC>
struct Writer {
    double *p0;        // element 0 of this thread's sequence
    unsigned ld;       // row pitch (LC: element stride of the sequence)
    double *pc;        // TS, interior chunks: element `ibase` (set once per chunk)
    int ibase;
    __device__ __forceinline__ Writer(double *base, long ld_, int q)
        : p0(LC ? base + q : base + (long)q * ld_), ld((unsigned)ld_), pc(p0), ibase(0) {}
    __device__ __forceinline__ void chunk(int i0)
    {
        if (!LC) {
            ibase = i0;
            pc = p0 + (unsigned)i0;
        }
    }
    // MID = interior chunk.  TS: i - ibase folds to a compile-time constant after inlining, so the store is
    // STG [pc + imm] (the plain p0[i] form spent 4 integer instructions per store on the sign-extended
    // 64-bit address: SASS of the first version).  LC: unsigned 32 x 32 -> 64 is ONE instruction
    // (IMAD.WIDE.U32); the signed / 64-bit forms cost 4-6.
    template <bool MID = false>
    __device__ __forceinline__ void st(int i, double v) const
    {
        if (LC) p0[(unsigned long long)(unsigned)i * ld] = v;
        else if (MID) pc[i - ibase] = v;
        else p0[i] = v;
    }
};

// Op interface:
//   static constexpr int NIN;  static constexpr bool ASC;
//   __host__ __device__ static int off(int s);                 index offset of stream s
//   __device__ static int len(int s, int n, job);     valid index range [0, len) of stream s
//   struct State;  __device__ static void init(State&, job, n, q);
//   template <class W> __device__ static void step(State&, job, n, i, const double *v, W &out);
template <class Op, bool LC, int BD>
__global__ void __launch_bounds__(BD) k_sweep(SweepJobs jobs)
{
    extern __shared__ double ring[];
    const SweepJob &job = jobs.j[blockIdx.y];
    const int q = blockIdx.x * BD + threadIdx.x;
    if (blockIdx.x * BD >= job.nseq) return;          // whole CTA beyond this job's sequences
    const int n = jobs.n;
    double *my = ring + threadIdx.x;
    constexpr int NIN = Op::NIN;
    constexpr int K = RING_K;
    constexpr int C = SW_C;
    // this thread walks the chain of parity par: i = par, par+2, ... (ASC) or top, top-2, ... (DESC)
    const int par = blockIdx.z;
    const int np = (n - par + 1) / 2;                 // chain length
    if (np <= 0) return;
    const int top = par + 2 * (np - 1);
    const bool active = q < job.nseq;                 // inactive lanes only help staging the tables

    // per-stream base pointer of this sequence and element stride (in doubles)
    const double *gq[NIN];
    long es[NIN];
    int slen[NIN];
#pragma unroll
    for (int s = 0; s < NIN; ++s) {
        es[s] = LC ? job.ldin[s] : 1;
        gq[s] = job.in[s] ? (LC ? job.in[s] + q : job.in[s] + (long)q * job.ldin[s]) : nullptr;
        slen[s] = job.in[s] ? Op::len(s, n, job) : 0;
    }
    Writer<LC> out(job.out, job.ldout, q);
    // L2 prefetch (TS only): rows must start on 16-byte boundaries
    bool pf_ok[NIN];
#pragma unroll
    for (int s = 0; s < NIN; ++s)
        pf_ok[s] = !LC && SW_PF > 0 && gq[s] != nullptr && ((unsigned long long)job.in[s] % 16 == 0) &&
                   (job.ldin[s] % 2 == 0);
    auto prefetch_chunk = [&](int c) {
        if (LC || SW_PF == 0) return;
        // elements [w0, w0 + 16) hold both parity chains' steps of chunk c (w0 even => 16-byte aligned)
        const int w0 = Op::ASC ? 2 * c * C : ((top - 2 * c * C - 2 * (C - 1)) & ~1);
#pragma unroll
        for (int s = 0; s < NIN; ++s)
            if (pf_ok[s] && w0 >= 0 && w0 + 2 * C <= slen[s]) prefetch_l2_128(gq[s] + w0);
    };

    // generic (checked) chunk issue: steps c*C .. c*C+C-1 into ring chunk slot c % KC
    auto issue_chunk = [&](int c) {
#pragma unroll
        for (int e = 0; e < C; ++e) {
            const int t = c * C + e;
            if (t >= np) break;
            const int i = Op::ASC ? par + 2 * t : top - 2 * t;
            const int slot = t & (K - 1);
#pragma unroll
            for (int s = 0; s < NIN; ++s) {
                const int ii = i + Op::off(s);
                if (ii >= 0 && ii < slen[s]) cp_async_8(my + (slot * NIN + s) * BD, gq[s] + (long)ii * es[s]);
            }
        }
    };
    // unchecked issue of a full interior chunk whose ring chunk slot CS is a compile-time constant
    auto issue_fast = [&](int c, auto cs_tag) {
        constexpr int CS = decltype(cs_tag)::value;
        const int i0 = Op::ASC ? par + 2 * c * C : top - 2 * c * C;
#pragma unroll
        for (int s = 0; s < NIN; ++s) {
            const long step = Op::ASC ? 2 * es[s] : -2 * es[s];
            const double *g = gq[s] + (long)(i0 + Op::off(s)) * es[s];
#pragma unroll
            for (int e = 0; e < C; ++e) {
                cp_async_8(my + ((CS * C + e) * NIN + s) * BD, g);
                g += step;
            }
        }
    };

    // Per-index coefficient tables: the parity half this CTA walks is staged once in shared memory
    // (entry t <-> index par + 2t).  __ldg table reads were the top stall of the first version (ncu:
    // long_scoreboard 9-13 per issue, L1 hit rate 4-50 % because the cp.async stream evicts them).
    constexpr int NT = Op::NT;
    const int H = (n + 1) / 2 + 2;
    double *tsm = ring + RING_K * NIN * BD;
    if (NT > 0) {
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const double *tp = job.tab[Op::tab_index(k)];
            const int tl = Op::tab_len(k, n);
            for (int t = threadIdx.x; t < H; t += BD) {
                const int i = par + 2 * t;
                tsm[t * NT + k] = (tp != nullptr && i < tl) ? __ldg(tp + i) : 0.0;
            }
        }
        __syncwarp();
    }
    if (!active) return;
    typename Op::State st;
    Op::init(st, job, n, q);
    st.tsm = tsm;
    st.set_chunk(par);
    const int nchunks = (np + C - 1) / C;
    // chunks [c_lo, c_hi) are "interior": full, every stream index valid, no edge logic in Op::step
    // (all their steps satisfy 8 <= i <= n-9)
    constexpr int c_lo_const = 1;
    const int c_lo = c_lo_const;
    const int c_hi = Op::ASC ? (n - 9 - par >= 0 ? ((n - 9 - par) / 2 + 1) / C : 0)
                             : (top - 8 >= 0 ? ((top - 8) / 2 + 1) / C : 0);
#pragma unroll 1
    for (int c = 0; c < SW_KC - 1; ++c) {
        if (c < nchunks) issue_chunk(c);
        cp_async_commit_group();
    }

    auto generic_chunk = [&](int c) {
        cp_async_wait_group<SW_KC - 2>();
        double v[C][NIN];
#pragma unroll
        for (int e = 0; e < C; ++e) {
            const int t = c * C + e;
            const int i = Op::ASC ? par + 2 * t : top - 2 * t;
            const int slot = t & (K - 1);
#pragma unroll
            for (int s = 0; s < NIN; ++s) {
                const int ii = i + Op::off(s);
                v[e][s] = (t < np && ii >= 0 && ii < slen[s]) ? my[(slot * NIN + s) * BD] : 0.0;
            }
        }
        if (c + SW_KC - 1 < nchunks) issue_chunk(c + SW_KC - 1);
        cp_async_commit_group();
        st.set_chunk(Op::ASC ? par + 2 * c * C : top - 2 * c * C);
#pragma unroll
        for (int e = 0; e < C; ++e) {
            const int t = c * C + e;
            if (t < np) Op::template step<false>(st, job, n, Op::ASC ? par + 2 * t : top - 2 * t, v[e], out);
        }
    };
    auto fast_chunk = [&](int c, auto cs_tag) {
        constexpr int CS = decltype(cs_tag)::value;                 // ring chunk slot of chunk c
        constexpr int NS = (CS + SW_KC - 1) % SW_KC;                // slot of chunk c + KC - 1
        cp_async_wait_group<SW_KC - 2>();
        double v[C][NIN];
#pragma unroll
        for (int e = 0; e < C; ++e)
#pragma unroll
            for (int s = 0; s < NIN; ++s) v[e][s] = my[((CS * C + e) * NIN + s) * BD];
        if (c + SW_KC - 1 < c_hi) issue_fast(c + SW_KC - 1, std::integral_constant<int, NS>{});
        else if (c + SW_KC - 1 < nchunks) issue_chunk(c + SW_KC - 1);
        cp_async_commit_group();
        prefetch_chunk(c + SW_KC - 1 + SW_PF);
        const int i0 = Op::ASC ? par + 2 * c * C : top - 2 * c * C;
        out.chunk(i0);
        st.set_chunk(i0);
#pragma unroll
        for (int e = 0; e < C; ++e) Op::template step<true>(st, job, n, Op::ASC ? i0 + 2 * e : i0 - 2 * e, v[e], out);
    };

    int c = 0;
#pragma unroll 1
    for (; c < nchunks && c < c_lo; ++c) generic_chunk(c);
    // steady state: SW_KC chunks per iteration so that ring offsets are compile-time constants
    // (c_lo = 1: chunk c lives in ring chunk slot c % SW_KC = (1 + u) % SW_KC)
    static_assert(c_lo_const == 1, "slot rotation below assumes the fast path starts at chunk 1");
#pragma unroll 1
    for (; c + SW_KC <= c_hi; c += SW_KC) {
        static_for<0, SW_KC>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            fast_chunk(c + u, std::integral_constant<int, (1 + u) % SW_KC>{});
        });
    }
#pragma unroll 1
    for (; c < nchunks; ++c) generic_chunk(c);
}

// ---------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------

// Every operator is templated on FULL: the batched stepper supplies every optional table
// (stencil, reciprocals, scale), so the FULL instantiation has no run-time feature tests; the
// generic C-ABI entry points use FULL = false.  A thread walks ONE parity chain, so the
// recurrence state is a single value.  Per-index tables are read from the shared-memory copy
// staged by k_sweep: table k, index i  ->  tsm[k*H + (i >> 1)]  (i has the chain's parity).
#define PDE_TB(st, k, i) ((st).tb((k), (i)))

// Where an operator finds its per-index tables:
//   TabParityT<NT> (k_sweep): the parity half this thread walks, staged per CTA: table k, index i -> tsm[(i >> 1)*NT + k]
//   TabTile   (k_sweep_tile): the slice [i0 - 2, i0 + TW + 2) of every table, staged per tile:  tt[k*TTP + (i - i0) + 2]
template <int NT>
struct TabParityT {
    // entry (table k, index i) of the parity half this thread walks: tsm[(i >> 1) * NT + k] (the tables interleaved, so
    // that one pointer per chunk -- pc, element `ibase` -- reaches every table entry of the chunk with a compile-time
    // offset; the first layout, tsm[k * H + (i >> 1)] with run-time H, cost two integer instructions per table read:
    // 14 IMAD + 5 LEA per step in the SASS of the 4-table back substitution)
    const double *tsm;
    const double *pc;
    int ibase;
    __device__ __forceinline__ void set_chunk(int i0)
    {
        ibase = i0;
        pc = tsm + (i0 >> 1) * NT;
    }
    __device__ __forceinline__ double tb(int k, int i) const { return pc[((i - ibase) >> 1) * NT + k]; }
};
using TabParity = TabParityT<0>;
constexpr int TILE_W = 64;               // elements of a sequence per tile
constexpr int TILE_P = TILE_W + 2;       // row pitch of a tile in shared memory (33 x 16 bytes: odd, conflict-free)
constexpr int TILE_TP = TILE_W + 4;      // table slice per tile
struct TabTile {
    const double *tt;
    int i0;
    __device__ __forceinline__ double tb(int k, int i) const { return tt[k * TILE_TP + (i - i0) + 2]; }
};

// differentiate_cheby.f90:28-53: dc[n-1] = 0, dc[n-2] = 2(n-1)c[n-1], dc[k] = dc[k+2] + 2(k+1)c[k+1],
// dc[0] = dc[2]/2 + c[1]; stored value divided by job.sc when job.flag (grad(): /= scale**deriv).
// POW2: every job's scale is a power of two (aspect 1: scale = 1/2), so x / sc == x * RN(1/sc) bit for bit
// and the 5-op division sequence becomes one multiplication.
template <bool FULL, bool POW2 = false, class TAB = TabParity>
struct DiffDesc {
    static constexpr int NIN = 1;
    static constexpr int NT = 0;
    static constexpr bool ASC = false;
    // k_sweep_tile: step index of output k is k + ISHIFT; HALO: reads input element i + 2 (or k + 1);
    // PARTIAL: some indices are not stored (in place only: the tile is pre-filled with the input)
    static constexpr int ISHIFT = 1;
    static constexpr bool HALO = true;
    static constexpr bool PARTIAL = false;
    __host__ __device__ static int off(int) { return 0; }
    __device__ static int len(int, int n, const SweepJob &) { return n; }
    __host__ __device__ static int tab_index(int) { return 0; }
    __host__ __device__ static int tab_len(int, int) { return 0; }
    struct State : TAB {
        double p;          // dc[k+2] of this chain
        double sc, rsc;    // scale and RN(1/scale): the division is a 5-op correctly rounded sequence
        bool div;
    };
    __device__ static void init(State &s, const SweepJob &job, int, int)
    {
        s.p = 0.0;
        s.div = (FULL ? true : job.flag != 0) && job.sc != 1.0;      // x / 1.0 == x: skip the division sequence
        s.sc = job.sc;
        s.rsc = 1.0 / job.sc;
    }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int n, int i, const double *v, W &out)
    {
        const int k = i - 1;
        double cur;
        if (MID) {
            cur = s.p + (double)(2 * i) * v[0];
        } else {
            if (i == n - 1) out.template st<false>(n - 1, 0.0);
            if (i == 0) return;
            if (i == n - 1) cur = (double)(2 * (n - 1)) * v[0];
            else if (k >= 1) cur = s.p + (double)(2 * i) * v[0];
            else cur = s.p / 2.0 + v[0];
        }
        s.p = cur;
        if (POW2) out.template st<MID>(k, cur * s.rsc);
        else out.template st<MID>(k, s.div ? div_rn_v(cur, s.sc, s.rsc, true) : cur);
    }
};

// tdma.f90:55-106, k = 2, forward part: g_i = (rhs_i - a_{i-2} g_{i-2}) / den_i with the fused
// S^T product rhs_i = u_i + s_i u_{i+2} (chebyshev.py:327) when tab[0] = s is given.
// job.tab: 0 = s (or null), 1 = a, 2 = den, 3 = w (back substitution), 4 = RN(1/den) (optional).
template <bool FULL, class TAB = TabParityT<4>>
struct TdmaFwd {
    static constexpr int NIN = 2;
    static constexpr int NT = 4;          // staged: 0 = s, 1 = a, 2 = den, 3 = rden
    static constexpr bool ASC = true;
    // k_sweep_tile: step index of output k is k + ISHIFT; HALO: reads input element i + 2 (or k + 1);
    // PARTIAL: some indices are not stored (in place only: the tile is pre-filled with the input)
    static constexpr int ISHIFT = 0;
    static constexpr bool HALO = true;
    static constexpr bool PARTIAL = false;
    __host__ __device__ static int off(int s) { return s == 0 ? 0 : 2; }
    __device__ static int len(int s, int n, const SweepJob &job)
    {
        if (s == 0) return n;
        return job.tab[0] ? n + 2 : 0;
    }
    __host__ __device__ static int tab_index(int k) { return k == 3 ? 4 : k; }
    __host__ __device__ static int tab_len(int k, int n) { return k == 1 ? n - 2 : n; }
    struct State : TAB {
        double g;
        bool has_s, has_r;
    };
    __device__ static void init(State &s, const SweepJob &job, int, int)
    {
        s.g = 0.0;
        s.has_s = FULL || job.tab[0] != nullptr;
        s.has_r = FULL || job.tab[4] != nullptr;
    }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int, int i, const double *v, W &out)
    {
        double rhs = v[0];
        if (s.has_s) rhs = v[0] + PDE_TB(s, 0, i) * v[1];
        const double den = PDE_TB(s, 2, i);
        double g;
        if (!MID && i < 2) g = rhs / den;
        else g = div_rn_v(rhs - PDE_TB(s, 1, i - 2) * s.g, den, PDE_TB(s, 3, i), s.has_r);
        s.g = g;
        out.template st<MID>(i, g);
    }
};

// back substitution x_i = g_i - w_i x_{i+2} (in place), job.tab[3] = w
template <bool FULL, class TAB = TabParityT<1>>
struct TdmaBwd {
    static constexpr int NIN = 1;
    static constexpr int NT = 1;
    static constexpr bool ASC = false;
    // k_sweep_tile: step index of output k is k + ISHIFT; HALO: reads input element i + 2 (or k + 1);
    // PARTIAL: some indices are not stored (in place only: the tile is pre-filled with the input)
    static constexpr int ISHIFT = 0;
    static constexpr bool HALO = false;
    static constexpr bool PARTIAL = true;
    __host__ __device__ static int off(int) { return 0; }
    __device__ static int len(int, int n, const SweepJob &) { return n; }
    __host__ __device__ static int tab_index(int) { return 3; }
    __host__ __device__ static int tab_len(int, int n) { return n - 2; }
    struct State : TAB {
        double x;
    };
    __device__ static void init(State &s, const SweepJob &, int, int) { s.x = 0.0; }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int n, int i, const double *v, W &out)
    {
        double x = v[0];
        if (MID || i < n - 2) {
            x = v[0] - PDE_TB(s, 0, i) * s.x;
            out.template st<MID>(i, x);
        }
        s.x = x;
    }
};

// fdma.f90:26-36: forward x_i -= l_{i-2} x_{i-2}; job.tab: 0 = l, 1 = d, 2 = u1, 3 = u2, 4 = RN(1/d) (optional)
template <bool FULL, class TAB = TabParityT<1>>
struct FdmaFwd {
    static constexpr int NIN = 1;
    static constexpr int NT = 1;
    static constexpr bool ASC = true;
    // k_sweep_tile: step index of output k is k + ISHIFT; HALO: reads input element i + 2 (or k + 1);
    // PARTIAL: some indices are not stored (in place only: the tile is pre-filled with the input)
    static constexpr int ISHIFT = 0;
    static constexpr bool HALO = false;
    static constexpr bool PARTIAL = true;
    __host__ __device__ static int off(int) { return 0; }
    __device__ static int len(int, int n, const SweepJob &) { return n; }
    __host__ __device__ static int tab_index(int) { return 0; }
    __host__ __device__ static int tab_len(int, int n) { return n - 2; }
    struct State : TAB {
        double p;
    };
    __device__ static void init(State &s, const SweepJob &, int, int) { s.p = 0.0; }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int, int i, const double *v, W &out)
    {
        double x = v[0];
        if (MID || i >= 2) {
            x = v[0] - PDE_TB(s, 0, i - 2) * s.p;
            out.template st<MID>(i, x);
        }
        s.p = x;
    }
};

template <bool FULL, class TAB = TabParityT<4>>
struct FdmaBwd {
    static constexpr int NIN = 1;
    static constexpr int NT = 4;          // staged: 0 = d, 1 = u1, 2 = u2, 3 = rd
    static constexpr bool ASC = false;
    // k_sweep_tile: step index of output k is k + ISHIFT; HALO: reads input element i + 2 (or k + 1);
    // PARTIAL: some indices are not stored (in place only: the tile is pre-filled with the input)
    static constexpr int ISHIFT = 0;
    static constexpr bool HALO = false;
    static constexpr bool PARTIAL = false;
    __host__ __device__ static int off(int) { return 0; }
    __device__ static int len(int, int n, const SweepJob &) { return n; }
    __host__ __device__ static int tab_index(int k) { return k + 1; }
    __host__ __device__ static int tab_len(int k, int n) { return k == 1 ? n - 2 : (k == 2 ? n - 4 : n); }
    struct State : TAB {
        double x2, x4;     // x_{i+2}, x_{i+4} of this chain
        bool has_r;
    };
    __device__ static void init(State &s, const SweepJob &job, int, int)
    {
        s.x2 = s.x4 = 0.0;
        s.has_r = FULL || job.tab[4] != nullptr;
    }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int n, int i, const double *v, W &out)
    {
        const double d = PDE_TB(s, 0, i);
        double x;
        if (!MID && i >= n - 2) x = v[0] / d;
        else if (!MID && i >= n - 4) x = (v[0] - PDE_TB(s, 1, i) * s.x2) / d;
        else x = div_rn_v(v[0] - PDE_TB(s, 1, i) * s.x2 - PDE_TB(s, 2, i) * s.x4, d, PDE_TB(s, 3, i), s.has_r);
        out.template st<MID>(i, x);
        s.x4 = s.x2;
        s.x2 = x;
    }
};

// twodma.f90:17-22; job.tab: 0 = d, 1 = u, 4 = RN(1/d) (optional)
template <bool FULL, class TAB = TabParityT<3>>
struct TwodmaBwd {
    static constexpr int NIN = 1;
    static constexpr int NT = 3;          // staged: 0 = d, 1 = u, 2 = rd
    static constexpr bool ASC = false;
    // k_sweep_tile: step index of output k is k + ISHIFT; HALO: reads input element i + 2 (or k + 1);
    // PARTIAL: some indices are not stored (in place only: the tile is pre-filled with the input)
    static constexpr int ISHIFT = 0;
    static constexpr bool HALO = false;
    static constexpr bool PARTIAL = false;
    __host__ __device__ static int off(int) { return 0; }
    __device__ static int len(int, int n, const SweepJob &) { return n; }
    __host__ __device__ static int tab_index(int k) { return k == 2 ? 4 : k; }
    __host__ __device__ static int tab_len(int k, int n) { return k == 1 ? n - 2 : n; }
    struct State : TAB {
        double x;
        bool has_r;
    };
    __device__ static void init(State &s, const SweepJob &job, int, int)
    {
        s.x = 0.0;
        s.has_r = job.tab[4] != nullptr;
    }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int n, int i, const double *v, W &out)
    {
        const double d = PDE_TB(s, 0, i);
        double x;
        if (!MID && i >= n - 2) x = v[0] / d;
        else x = div_rn_v(v[0] - PDE_TB(s, 1, i) * s.x, d, PDE_TB(s, 2, i), s.has_r);
        out.template st<MID>(i, x);
        s.x = x;
    }
};

// Poisson (A + lam_q C) columns with per-column LU tables (n x m arrays, same layout as x):
// streams: 0 = x, 1 = L (read at i-2);  itab[q] = 1 where the singular branch drops row/col 0
// (fdma.f90:173-185): that column's system starts at i = 1 and x[0] = 0.
template <bool FULL, class TAB = TabParityT<0>>
struct PoissonFwd {
    static constexpr int NIN = 2;
    static constexpr int NT = 0;
    static constexpr bool ASC = true;
    __host__ __device__ static int off(int s) { return s == 0 ? 0 : -2; }
    __device__ static int len(int, int n, const SweepJob &) { return n; }
    __host__ __device__ static int tab_index(int) { return 0; }
    __host__ __device__ static int tab_len(int, int) { return 0; }
    struct State : TAB {
        double p;
        int off;
    };
    __device__ static void init(State &s, const SweepJob &job, int, int q)
    {
        s.p = 0.0;
        s.off = job.itab[q];
    }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int, int i, const double *v, W &out)
    {
        if (!MID && i < s.off) {
            out.template st<MID>(i, 0.0);
            return;
        }
        double x = v[0];
        if (MID || i >= s.off + 2) {
            x = v[0] - v[1] * s.p;
            out.template st<MID>(i, x);
        }
        s.p = x;
    }
};

// streams: 0 = x, 1 = D, 2 = U1, 3 = U2, 4 = RN(1/D)
template <bool FULL, class TAB = TabParityT<0>>
struct PoissonBwd {
    static constexpr int NIN = 5;
    static constexpr int NT = 0;
    static constexpr bool ASC = false;
    __host__ __device__ static int off(int) { return 0; }
    __device__ static int len(int, int n, const SweepJob &) { return n; }
    __host__ __device__ static int tab_index(int) { return 0; }
    __host__ __device__ static int tab_len(int, int) { return 0; }
    struct State : TAB {
        double x2, x4;
        int off;
    };
    __device__ static void init(State &s, const SweepJob &job, int, int q)
    {
        s.x2 = s.x4 = 0.0;
        s.off = job.itab[q];
    }
    template <bool MID, class W>
    __device__ static void step(State &s, const SweepJob &, int n, int i, const double *v, W &out)
    {
        if (!MID && i < s.off) return;
        double x;
        if (!MID && i >= n - 2) x = v[0] / v[1];
        else if (!MID && i >= n - 4) x = (v[0] - v[2] * s.x2) / v[1];
        else x = div_rn_v(v[0] - v[2] * s.x2 - v[3] * s.x4, v[1], v[4], true);
        out.template st<MID>(i, x);
        s.x4 = s.x2;
        s.x2 = x;
    }
};


// ---------------------------------------------------------------------------
// TS sweeps, tiled: k_sweep_tile
// ---------------------------------------------------------------------------
// k_sweep<Op, TS> gives every lane its own row and streams it with 8-byte copies: DRAM sees isolated
// 32-byte sectors of 32 rows per warp step, each parity chain is a separate warp on another SM (both fetch
// every sector), and a chain step costs ~28 instructions of ONE warp (ncu: 1.3-1.7 TB/s, IPC 0.3-0.45).
// Here a warp owns 32 rows and moves them tile by tile (TILE_W = 64 elements of every row):
//   * load: 32 + 1 cp.async instructions, each one contiguous 512-byte row segment (lane l copies the
//     16-byte pair l), NST - 1 tiles in flight per warp; the per-index tables of the tile ride along;
//   * chain: the CTA is two warps, warp w walks the parity-w chain of row `lane` in shared memory (the first
//     version put both chains into one thread: half as many warps, each with twice the instruction stream --
//     slower, because these kernels are bound by the issue latency of the few resident warps);
//   * store: the output tile goes back as 32 contiguous 512-byte row segments.
// Same Op::step as k_sweep, so the arithmetic (and every bit of the result) is unchanged.
__device__ __forceinline__ void cp_async_16z(double *smem, const double *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}

struct TileWriter {
    double *row;       // this lane's row of the output tile
    int base;          // index of its first element
    template <bool MID = false>
    __device__ __forceinline__ void st(int i, double v) const { row[i - base] = v; }
};

template <class Op>
struct TileStage {
    static constexpr int value = 32 * TILE_P + ((Op::NT * TILE_TP + 1) & ~1);    // doubles: rows + table slices
};

template <class Op, int NST>
__global__ void __launch_bounds__(64) k_sweep_tile(SweepJobs jobs)
{
    extern __shared__ __align__(16) double tsm_[];
    constexpr int TW = TILE_W, TP = TILE_P, TTP = TILE_TP, NT = Op::NT;
    constexpr int STAGE = TileStage<Op>::value;
    constexpr int IS = Op::ISHIFT;
    const SweepJob &job = jobs.j[blockIdx.y];
    // two warps per CTA: warp w walks the chain of parity w of the CTA's 32 rows (lane = row) and moves rows
    // 16 w .. 16 w + 15 of every tile (lane = 16-byte pair of the row segment)
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, q0 = blockIdx.x * 32;
    if (q0 >= job.nseq) return;
    const int n = jobs.n;
    const int nrows = min(32, job.nseq - q0);
    const unsigned ldi = (unsigned)job.ldin[0], ldo = (unsigned)job.ldout;
    const int r_lo = 16 * wrp, r_hi = min(r_lo + 16, nrows);          // rows this warp copies
    const double *gin = job.in[0] + (long)q0 * ldi;
    double *gout = job.out + (long)q0 * ldo;
    const int in_len = (Op::HALO && IS == 0) ? (job.in[1] ? n + 2 : n) : n;     // TdmaFwd reads u_{i+2} of the n + 2 inputs
    const int ntiles = (n + TW - 1) / TW;
    double *out_tile = tsm_ + NST * STAGE;

    auto issue = [&](int u) {
        if (u >= ntiles) return;
        const int tbase = (Op::ASC ? u : ntiles - 1 - u) * TW;
        double *ti = tsm_ + (u % NST) * STAGE;
        const int col = tbase + 2 * lane;
        const int nb = col < in_len ? 16 : 0;
        const double *g = gin + (nb ? col : 0);
        double *d = ti + 2 * lane;
        if (r_hi - r_lo == 16) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
                cp_async_16z(d + (r_lo + r) * TP, g + (unsigned long long)((unsigned)(r_lo + r) * ldi), nb);
        } else {
            for (int r = r_lo; r < r_hi; ++r) cp_async_16z(d + r * TP, g + (unsigned long long)((unsigned)r * ldi), nb);
        }
        if (Op::HALO && wrp == 0 && lane < nrows) {
            const int ch = tbase + TW;
            cp_async_16z(ti + lane * TP + TW, gin + (unsigned long long)((unsigned)lane * ldi) + (ch < in_len ? ch : 0),
                         ch < in_len ? 16 : 0);
        }
        if (NT > 0 && wrp == 1) {
            double *tt = ti + 32 * TP;
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const double *tp = job.tab[Op::tab_index(k)];
                const int tl = tp ? Op::tab_len(k, n) : 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pr = lane + 32 * h;                 // pair of the slice [tbase - 2, tbase + TW + 2)
                    if (pr >= TTP / 2) break;
                    const int e0 = tbase - 2 + 2 * pr;
                    const int nbt = (e0 >= 0 && e0 + 1 < tl) ? 16 : ((e0 >= 0 && e0 < tl) ? 8 : 0);
                    cp_async_16z(tt + k * TTP + 2 * pr, nbt ? tp + e0 : gin, nbt);
                }
            }
        }
    };

    typename Op::State st;
    Op::init(st, job, n, q0 + lane);
    const bool active = lane < nrows;

#pragma unroll 1
    for (int u = 0; u < NST - 1; ++u) {
        issue(u);
        cp_async_commit_group();
    }
#pragma unroll 1
    for (int u = 0; u < ntiles; ++u) {
        cp_async_wait_group<NST - 2>();
        __syncthreads();         // tile u has landed for every thread; everybody is done with tile u - 1 (in and out)
        issue(u + NST - 1);
        cp_async_commit_group();
        const int tbase = (Op::ASC ? u : ntiles - 1 - u) * TW;
        const double *ti = tsm_ + (u % NST) * STAGE;
        const double *rin = ti + lane * TP;
        TileWriter out{out_tile + lane * TP, tbase};
        st.tt = ti + 32 * TP;
        st.i0 = tbase;
        // all steps of an interior tile satisfy 8 <= i <= n - 9: no edge logic (Op::step<true>)
        const bool mid = tbase >= 8 && tbase + TW + IS + 8 <= n;
        if (active) {
            if (Op::PARTIAL) {                                // unstored indices keep the input (in place)
                if (!mid) {
#pragma unroll 8
                    for (int j = wrp; j < TW; j += 2) out_tile[lane * TP + j] = rin[j];
                }
            }
            // output k = tbase + e has step index i = k + IS; this warp takes the i of its parity
            const int e0 = (wrp + IS) & 1;                    // first e of this warp's parity (tbase is even)
            auto block = [&](int b, auto mid_tag) {           // 8 of the 16 outputs e = 16 b .. 16 b + 15
                constexpr bool MID = decltype(mid_tag)::value;
                const double *rb = rin + 16 * b + e0 + IS;   // input element of step e = 16 b + e0
                const int ib = tbase + 16 * b + e0 + IS;
                double v[8][Op::NIN];
#pragma unroll
                for (int t = 0; t < 8; ++t)
#pragma unroll
                    for (int s = 0; s < Op::NIN; ++s) v[t][s] = rb[2 * t + Op::off(s)];
                static_for<0, 8>([&](auto tc) {
                    constexpr int t = Op::ASC ? decltype(tc)::value : 7 - decltype(tc)::value;
                    const int i = ib + 2 * t;
                    if (MID || (i >= 0 && i < n)) Op::template step<MID>(st, job, n, i, v[t], out);
                });
            };
#pragma unroll 1
            for (int bb = 0; bb < TW / 16; ++bb) {
                const int b = Op::ASC ? bb : TW / 16 - 1 - bb;
                if (mid) block(b, std::true_type{});
                else block(b, std::false_type{});
            }
        }
        __syncthreads();
        // store: row r of the tile = one contiguous segment of <= 512 bytes
        const int cnt = min(TW, n - tbase);
        if (2 * lane < cnt) {
            const double *src = out_tile + 2 * lane;
            double *dst = gout + tbase + 2 * lane;
            if (r_hi - r_lo == 16) {
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    *reinterpret_cast<double2 *>(dst + (unsigned long long)((unsigned)(r_lo + r) * ldo)) =
                        *reinterpret_cast<const double2 *>(src + (r_lo + r) * TP);
            } else {
                for (int r = r_lo; r < r_hi; ++r)
                    *reinterpret_cast<double2 *>(dst + (unsigned long long)((unsigned)r * ldo)) =
                        *reinterpret_cast<const double2 *>(src + r * TP);
            }
        }
    }
}

// eligibility of the tiled TS kernel: even lengths and pitches, 16-byte aligned rows and tables, one input array
template <class Op>
static bool sweep_tile_ok(const SweepJobs &jobs)
{
    auto al = [](const void *q) { return ((uintptr_t)q % 16) == 0; };
    if (jobs.n % 2 != 0 || jobs.n < 2 * TILE_W) return false;
    for (int j = 0; j < jobs.njobs; ++j) {
        const SweepJob &jb = jobs.j[j];
        if (!jb.in[0] || !jb.out || !al(jb.in[0]) || !al(jb.out) || jb.ldin[0] % 2 || jb.ldout % 2) return false;
        if (jb.ldin[0] >= (1L << 26) || jb.ldout >= (1L << 26)) return false;      // 32-bit row offsets
        for (int s = 1; s < Op::NIN; ++s)
            if (jb.in[s] && (jb.in[s] != jb.in[0] || jb.ldin[s] != jb.ldin[0])) return false;
        for (int k = 0; k < Op::NT; ++k)
            if (jb.tab[Op::tab_index(k)] && !al(jb.tab[Op::tab_index(k)])) return false;
        if (Op::PARTIAL && jb.out != jb.in[0]) return false;                  // unstored indices keep the input
        if (Op::HALO && !Op::ASC && jb.out == jb.in[0]) return false;         // the halo of a later tile would be overwritten
    }
    return true;
}

template <template <bool, class> class OpT, bool FULL>
static int launch_sweep_tile(const SweepJobs &jobs, cudaStream_t st, const char *what)
{
    using Op = OpT<FULL, TabTile>;
    constexpr int NST = 3;
    int maxseq = 0;
    for (int j = 0; j < jobs.njobs; ++j) maxseq = jobs.j[j].nseq > maxseq ? jobs.j[j].nseq : maxseq;
    if (maxseq <= 0) return PDE_OK;
    const size_t smem = ((size_t)NST * TileStage<Op>::value + 32 * TILE_P) * sizeof(double);
    static PerDeviceFlag attr;
    if (!attr.get()) {
        PDE_CUDA(cudaFuncSetAttribute(k_sweep_tile<Op, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.get() = true;
    }
    dim3 grid(ceil_div(maxseq, 32), jobs.njobs);
    k_sweep_tile<Op, NST><<<grid, 64, smem, st>>>(jobs);
    return after_launch(what);
}

template <class Op>
static int launch_sweep(const SweepJobs &jobs, int axis, cudaStream_t st, const char *what)
{
    if (jobs.njobs <= 0 || jobs.n <= 0) return PDE_OK;
    int maxseq = 0;
    for (int j = 0; j < jobs.njobs; ++j) maxseq = jobs.j[j].nseq > maxseq ? jobs.j[j].nseq : maxseq;
    if (maxseq <= 0) return PDE_OK;
    constexpr int BD = 32;      // one warp per CTA: spreads the few thousand chains over all SMs
    dim3 grid(ceil_div(maxseq, BD), jobs.njobs, 2);      // z = parity chain
    const int H = (jobs.n + 1) / 2 + 2;
    const size_t smem = ((size_t)RING_K * Op::NIN * BD + (size_t)Op::NT * H) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("%s: sequence length %d needs %zu bytes of shared memory", what, jobs.n, smem);
        return PDE_ERR_UNSUPPORTED;
    }
    static PerDeviceSize attr_lc_, attr_ts_;
    size_t &attr_lc = attr_lc_.get(), &attr_ts = attr_ts_.get();
    if (axis == 0) {
        if (smem > 48 * 1024 && smem > attr_lc) {
            PDE_CUDA(cudaFuncSetAttribute(k_sweep<Op, true, BD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_lc = 200 * 1024;
        }
        k_sweep<Op, true, BD><<<grid, BD, smem, st>>>(jobs);
    } else {
        if (smem > 48 * 1024 && smem > attr_ts) {
            PDE_CUDA(cudaFuncSetAttribute(k_sweep<Op, false, BD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_ts = 200 * 1024;
        }
        k_sweep<Op, false, BD><<<grid, BD, smem, st>>>(jobs);
    }
    return after_launch(what);
}

}  // namespace pde
