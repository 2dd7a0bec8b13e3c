// Compile-time specialised shared-memory FFT DCT-I (included by dct_fft.cu).
//
// ncu on the generic kernel (runtime P / radices) showed ~400 executed instructions per
// point with the fp64 pipe 9 % busy: integer divisions for the butterfly indices, table
// look-ups for the digit reversal, six radix-4/3 passes with a barrier each.  Here the
// transform length, the radix list, the sequences per CTA (S) and the block size (T) are
// template constants: all index arithmetic folds to shifts / multiply-high, the pass loop
// is unrolled, radix-16 / radix-8 butterflies stay in registers (3072 = 16*16*4*3 -> four
// shared-memory round trips instead of six) and the digit reversal is arithmetic.
// Same algorithm and data flow as k_dct_fft (see dct_fft.cu).
#pragma once

namespace pde {

__device__ __forceinline__ void cp_async8_fft(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}

template <>
__device__ __forceinline__ void dft<8>(double2 *a)
{
    // n = n0 + 2 n1 (n0 < 2, n1 < 4), k = k1 + 4 k0: 4-point DFTs over n1, twiddle W8^{n0 k1}, 2-point over n0
    const double h = 0.70710678118654752440;
    double2 e[4] = {a[0], a[2], a[4], a[6]}, o[4] = {a[1], a[3], a[5], a[7]};
    dft<4>(e);
    dft<4>(o);
    o[1] = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * W8^1 = (1 - i)/sqrt2
    o[2] = mul_mi(o[2]);                                                    // * W8^2 = -i
    o[3] = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * W8^3 = (-1 - i)/sqrt2
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = cadd(e[k], o[k]);
        a[k + 4] = csub(e[k], o[k]);
    }
}

template <>
__device__ __forceinline__ void dft<16>(double2 *a)
{
    // n = n0 + 4 n1, k = k1 + 4 k0:  T[n0][k1] = DFT4_{n1} a[n0 + 4 n1];  U = T * W16^{n0 k1};
    // X[k1 + 4 k0] = DFT4_{n0} U[n0][k1]
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;   // cos, sin(pi/8)
    const double h = 0.70710678118654752440;
    double2 t[4][4];
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0) {
        double2 v[4] = {a[n0], a[n0 + 4], a[n0 + 8], a[n0 + 12]};
        dft<4>(v);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) t[n0][k1] = v[k1];
    }
    // twiddles W16^m = exp(-i pi m / 8), m = n0 * k1
    auto w1 = [&](double2 z) { return make_double2(c1 * z.x + s1 * z.y, c1 * z.y - s1 * z.x); };   // m = 1
    auto w2 = [&](double2 z) { return make_double2(h * (z.x + z.y), h * (z.y - z.x)); };           // m = 2
    auto w3 = [&](double2 z) { return make_double2(s1 * z.x + c1 * z.y, s1 * z.y - c1 * z.x); };   // m = 3
    auto w6 = [&](double2 z) { return make_double2(h * (z.y - z.x), -h * (z.x + z.y)); };          // m = 6
    t[1][1] = w1(t[1][1]);
    t[1][2] = w2(t[1][2]);
    t[1][3] = w3(t[1][3]);
    t[2][1] = w2(t[2][1]);
    t[2][2] = mul_mi(t[2][2]);                                  // m = 4
    t[2][3] = w6(t[2][3]);
    t[3][1] = w3(t[3][1]);
    t[3][2] = w6(t[3][2]);
    {                                                           // m = 9: W16^9 = -W16^1
        double2 z = w1(t[3][3]);
        t[3][3] = make_double2(-z.x, -z.y);
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        double2 v[4] = {t[0][k1], t[1][k1], t[2][k1], t[3][k1]};
        dft<4>(v);
#pragma unroll
        for (int k0 = 0; k0 < 4; ++k0) a[k1 + 4 * k0] = v[k0];
    }
}

template <>
__device__ __forceinline__ void dft<12>(double2 *a)
{
    // n = n0 + 3 n1 (n0 < 3, n1 < 4), k = k1 + 4 k0:  T[n0][k1] = DFT4_{n1} a[n0 + 3 n1];  U = T * W12^{n0 k1};
    // X[k1 + 4 k0] = DFT3_{n0} U[n0][k1].  3072 = 16 * 16 * 12: the last pass (no twiddles) replaces the
    // radix-4 and radix-3 passes, three shared-memory round trips instead of four.
    const double c = 0.86602540378443864676, h = 0.5;                 // cos, sin(pi/6)
    double2 t[3][4];
#pragma unroll
    for (int n0 = 0; n0 < 3; ++n0) {
        double2 v[4] = {a[n0], a[n0 + 3], a[n0 + 6], a[n0 + 9]};
        dft<4>(v);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) t[n0][k1] = v[k1];
    }
    // W12^m = exp(-i pi m / 6):  m = 1: (c, -h);  2: (h, -c);  3: -i;  4: (-h, -c);  6: -1
    t[1][1] = make_double2(c * t[1][1].x + h * t[1][1].y, c * t[1][1].y - h * t[1][1].x);
    t[1][2] = make_double2(h * t[1][2].x + c * t[1][2].y, h * t[1][2].y - c * t[1][2].x);
    t[1][3] = mul_mi(t[1][3]);
    t[2][1] = make_double2(h * t[2][1].x + c * t[2][1].y, h * t[2][1].y - c * t[2][1].x);
    t[2][2] = make_double2(c * t[2][2].y - h * t[2][2].x, -(c * t[2][2].x + h * t[2][2].y));
    t[2][3] = make_double2(-t[2][3].x, -t[2][3].y);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        double2 v[3] = {t[0][k1], t[1][k1], t[2][k1]};
        dft<3>(v);
#pragma unroll
        for (int k0 = 0; k0 < 3; ++k0) a[k1 + 4 * k0] = v[k0];
    }
}

// Shared-memory layout of one sequence: logical index i lives at i + i / M1 (M1 = P / first radix):
// one pad element after each first-level block.  The split step reads the digit-reversed FFT output
// with stride M1 between consecutive k; M1 is a multiple of 8 elements (= all banks), so without the
// pad those reads were 8-way bank conflicts (ncu: 48 % of the shared wavefronts).
//
// SP extra elements follow each sequence.  An axis-0 CTA holds S adjacent columns and its load / split
// phases map a quarter-warp to (8 / S elements) x (S sequences): with a sequence pitch that is a multiple
// of 8 elements the S lanes of one element hit the same bank group (4-way conflicts on 2 of the ~11
// shared-memory sweeps of the kernel); a pitch of 8 / S (mod 8) spreads them over all eight.
template <int P, int M1, int SP = 0>
struct Pad {
    static constexpr int SEQ = P + P / M1 + SP;            // padded sequence length
    __device__ __forceinline__ static int phys(int i) { return i + i / M1; }
};
template <int P, int M1, int S, int AXIS>
struct SeqPad {
    static constexpr int value =
        (AXIS != 0 || S <= 1 || 8 % S != 0) ? 0 : ((8 / (S > 0 ? S : 1) - (P + P / M1) % 8) % 8 + 8) % 8;
};

// Register cap: what the CTAs that fit an SM by shared memory (at most 3) can share.  Registers are granted
// per warp in coarse units (ncu: a 106-register kernel = 3392 per warp was limited to TWO 6-warp CTAs, i.e. it
// was charged 4096), so the cap is rounded down to a multiple of 32 per thread: 96 is what lets three
// 192-thread CTAs run together.  Never below 96 (the radix-16 butterfly needs ~100): fewer CTAs instead.
constexpr int fft_regs(int ctas, int T)
{
    const int per_thread = 65536 / (ctas * (T / 32)) / 1024 * 1024 / 32;
    return per_thread > 168 ? 168 : per_thread;
}
template <int P, int S, int T, int R1>
struct FftRegs {
    static constexpr int SMEM = S * Pad<P, P / R1, 7>::SEQ * 16 + 1024;
    static constexpr int FIT = 227 * 1024 / SMEM;
    static constexpr int C3 = FIT < 1 ? 1 : (FIT > 3 ? 3 : FIT);
    static constexpr int CTAS = fft_regs(C3, T) >= 96 ? C3 : (C3 > 1 && fft_regs(C3 - 1, T) >= 96 ? C3 - 1 : 1);
    static constexpr int value = fft_regs(CTAS, T);
};

// Twiddles of a DIF pass.  WOFF < 0: the plain table W[t] = exp(-2 pi i t / P), read with stride (Bluestein
// kernels).  WOFF >= 0: per-pass tables laid out [r][j] (entry (r - 1) * M + j = W_NCUR^{j r}) starting at
// W + WOFF: consecutive lanes (consecutive j) read consecutive entries.  ncu on the strided version showed the
// L1TEX pipe 77 % busy -- a warp's request for W[j * r] touches up to 60 cache lines in the first pass, as
// many L1 cycles as all shared-memory traffic of the pass; same values, so results are bit-identical.
//
// Thread -> butterfly map of the passes after the first.  The butterflies of a pass are (block, j) with
// block < P / NCUR and j < M.  Walking j fastest (the obvious map) puts the 8 lanes of a quarter-warp on
// addresses blk * NCUR + j + r * M with M = 12, 3, 8, ... : 2- to 8-way bank conflicts (tools/sim_fft_banks.py:
// x1.33 / x2 on passes 2 / 3 of P = 3072, x8 on the radix-8 pass of P = 2048).  Instead consecutive
// butterflies walk the FIRST-LEVEL blocks (pitch M1 + 1 elements, odd, so 8 lanes hit 8 bank groups), then
// the blocks inside a first-level block, and j is the slowest index (its twiddles become warp broadcasts).
// Twiddle multiplication a_r *= w^r of one butterfly; ld(r) loads w^r from the pass table.  -DPDE_FFT_TWPOW: a radix-16
// butterfly loads only w^1, w^2, w^3, w^4, w^8, w^12 and forms the other nine as products w^{4q} w^b (9 complex
// multiplications instead of 9 16-byte L1 requests per thread: the kernels are L1TEX-bound, the fp64 pipe is 30 % busy).
// The products carry one more rounding (<= 1.5 ulp instead of 0.5 ulp per twiddle).
// Number of twiddles of a radix-16 butterfly that are requested BEFORE its inputs are read and transformed (the
// rest are requested while the first ones are multiplied).  ncu's source view of the 96-register row kernel showed
// every twiddle load issued a few instructions before its use: 30 % of all stall samples were long-scoreboard
// waits in the complex multiplications (the 46 KB first-pass table does not stay in the ~28 KB of L1 left next to
// 222 KB of shared memory).  Measured on the rbc2048 step (axis-1 DCT time per step): 0 early twiddles 2.23 ms,
// 3: 2.18, 4: 2.11, 5: 2.12, 6: 2.23 (spills).  Also tried: prefetch.global.L1 of all 15 lines while the row is in
// flight (2.19), and finishing / multiplying / storing the outputs in four groups with the next group's twiddles
// requested one group ahead (2.19).
#ifndef PDE_FFT_TWPRE
#define PDE_FFT_TWPRE 4
#endif
template <int R>
struct TwPre {
    static constexpr int N = (R == 16 && PDE_FFT_TWPRE > 0) ? PDE_FFT_TWPRE : 0;
};

// a_r *= w^r, r = 1 .. R-1; the first TwPre<R>::N twiddles come from `pre` (loaded by the caller ahead of the butterfly)
template <int R, int NP, class LD>
__device__ __forceinline__ void twiddle_mul_pre(double2 *a, const double2 *pre, LD ld)
{
#pragma unroll
    for (int r = 1; r < R; ++r) a[r] = cmul(a[r], r <= NP ? pre[r - 1] : ld(r));
}

template <int R, class LD>
__device__ __forceinline__ void twiddle_mul(double2 *a, LD ld)
{
#ifdef PDE_FFT_TWPOW
    if constexpr (R == 16) {
        const double2 w1 = ld(1), w2 = ld(2), w3 = ld(3);
        a[1] = cmul(a[1], w1);
        a[2] = cmul(a[2], w2);
        a[3] = cmul(a[3], w3);
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            const double2 wq = ld(4 * q);
            a[4 * q] = cmul(a[4 * q], wq);
            a[4 * q + 1] = cmul(a[4 * q + 1], cmul(wq, w1));
            a[4 * q + 2] = cmul(a[4 * q + 2], cmul(wq, w2));
            a[4 * q + 3] = cmul(a[4 * q + 3], cmul(wq, w3));
        }
        return;
    }
#endif
#pragma unroll
    for (int r = 1; r < R; ++r) a[r] = cmul(a[r], ld(r));
}

template <int WOFF, int P, int M1, int NCUR, int S, int T, int R, int SP = 0>
__device__ __forceinline__ void dif_pass_t(double2 *z, const double2 *__restrict__ W)
{
    constexpr int M = NCUR / R, PER_SEQ = P / R, TOTAL = S * PER_SEQ, TWS = P / NCUR;
    constexpr int PS = Pad<P, M1, SP>::SEQ;
    constexpr int RS = NCUR == P ? M + 1 : M;              // first pass: stride M1 + 1 (padded)
    constexpr int R1 = P / M1;                             // first radix = number of first-level blocks
    constexpr int NBLK = P / NCUR;                         // blocks of this level per sequence
    static_assert(NCUR == P || NBLK % R1 == 0, "levels nest inside the first-level blocks");
#ifdef PDE_FFT_NO_TW_REUSE
    constexpr bool TW_REUSE = false;
#else
    constexpr bool TW_REUSE = true;
#endif
    // First pass of a CTA that holds several sequences (axis 0): butterfly j of every sequence uses the same
    // R - 1 twiddles, and the [r][j] table of this pass is the large one (15 x 192 x 16 B = 46 KB, 4 L1 lines per
    // warp request).  ncu (r01_ncu_dct_fft_remap.csv): 0.6 M of the kernel's 1.0 M global load requests were
    // twiddles, ~17 % of the L1TEX cycles.  Keep them in registers across the thread's sequences.
    if constexpr (TW_REUSE && NCUR == P && S > 1 && WOFF >= 0 && M > 1 && T % PER_SEQ == 0 &&
                  S % (T / PER_SEQ) == 0) {
        constexpr int SPT = T / PER_SEQ;                   // sequences per sweep of the CTA
        const int j = (int)threadIdx.x % PER_SEQ, s0 = (int)threadIdx.x / PER_SEQ;
        double2 w[R - 1];
#pragma unroll
        for (int r = 1; r < R; ++r) w[r - 1] = __ldg(W + WOFF + (r - 1) * M + j);
#pragma unroll
        for (int sq = 0; sq < S / SPT; ++sq) {
            double2 *p = z + (s0 + sq * SPT) * PS + j;
            double2 a[R];
#pragma unroll
            for (int r = 0; r < R; ++r) a[r] = p[r * RS];
            dft<R>(a);
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(a[r], w[r - 1]);
#pragma unroll
            for (int r = 0; r < R; ++r) p[r * RS] = a[r];
        }
    } else {
#pragma unroll
    for (int b0 = 0; b0 < TOTAL; b0 += T) {
        const int b = b0 + threadIdx.x;
        if ((TOTAL % T) != 0 && b >= TOTAL) break;
        const int s = b / PER_SEQ;
        const int bb = b - s * PER_SEQ;
        int j, off;                                        // off = padded position of element (blk, j)
        if (NCUR == P) {
            j = bb;
            off = bb;
        } else {
            j = bb / NBLK;
            const int bi = bb - j * NBLK;
            const int first = bi % R1, inner = bi / R1;
            off = first * (M1 + 1) + inner * NCUR + j;
        }
        double2 *p = z + s * PS + off;
        double2 a[R];
        auto ldw = [&](int r) { return WOFF < 0 ? __ldg(W + j * (r * TWS)) : __ldg(W + WOFF + (r - 1) * M + j); };
        // (all 15 early in the 160-register axis-0 kernels: measured 2.06 ms instead of 2.03 per step -- no gain there)
        constexpr int NPRE = (M > 1 && WOFF >= 0) ? TwPre<R>::N : 0;
        double2 pre[NPRE > 0 ? NPRE : 1];
#pragma unroll
        for (int r = 1; r <= NPRE; ++r) pre[r - 1] = ldw(r);
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = p[r * RS];
        dft<R>(a);
        if (M > 1) {
            if constexpr (NPRE > 0) twiddle_mul_pre<R, NPRE>(a, pre, ldw);
            else twiddle_mul<R>(a, ldw);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * RS] = a[r];
    }
    }
}

template <int SP, int WOFF, int P, int M1, int NCUR, int S, int T, int R, int... Rest>
struct DifPassesW {
    __device__ __forceinline__ static void run(double2 *z, const double2 *__restrict__ W)
    {
        dif_pass_t<WOFF, P, M1, NCUR, S, T, R, SP>(z, W);
        __syncthreads();
        constexpr int NEXT = WOFF < 0 ? -1 : WOFF + (NCUR / R > 1 ? (R - 1) * (NCUR / R) : 0);
        if constexpr (sizeof...(Rest) > 0) DifPassesW<SP, NEXT, P, M1, NCUR / R, S, T, Rest...>::run(z, W);
    }
};

template <int P, int M1, int NCUR, int S, int T, int... RAD>
struct DifPasses : DifPassesW<0, -1, P, M1, NCUR, S, T, RAD...> {
};

// the passes after the first one (the fused load + first pass of k_dct_fft_t runs the first itself)
template <int R1, int... Rest>
struct TailPasses {
    template <int SP, int P, int M1, int S, int T>
    __device__ __forceinline__ static void run(double2 *z, const double2 *__restrict__ W)
    {
        if constexpr (sizeof...(Rest) > 0) DifPassesW<SP, (R1 - 1) * (P / R1), P, M1, P / R1, S, T, Rest...>::run(z, W);
    }
};

// host: the [r][j] tables of all passes, concatenated in pass order (offsets as in DifPassesW)
template <int R, int... Rest>
static void build_pass_tables(int P, int ncur, std::vector<double2> &out)
{
    const long double pi = 3.141592653589793238462643383279502884L;
    const int M = ncur / R, tws = P / ncur;
    if (M > 1)
        for (int r = 1; r < R; ++r)
            for (int j = 0; j < M; ++j) {
                const long double ang = 2.0L * pi * (long double)((long)j * r * tws) / (long double)P;
                out.push_back(make_double2((double)cosl(ang), (double)(-sinl(ang))));
            }
    if constexpr (sizeof...(Rest) > 0) build_pass_tables<Rest...>(P, M, out);
}

template <int R, int... Rest>
struct FirstRadix {
    static constexpr int value = R;
};

// position of output k after the in-place DIF passes (mixed-radix digit reversal)
template <int N, int R, int... Rest>
struct DigitRev {
    __device__ __forceinline__ static int pos(int k)
    {
        constexpr int n = N / R;
        if constexpr (sizeof...(Rest) > 0) return (k % R) * n + DigitRev<n, Rest...>::pos(k / R);
        else return (k % R) * n;
    }
};

template <int P, int S, int T, int AXIS, int... RAD>
__global__ void __maxnreg__((FftRegs<P, S, T, FirstRadix<RAD...>::value>::value))
k_dct_fft_t(const double2 *__restrict__ W, const double2 *__restrict__ CS, int mode, DctPtrs ptrs, long ldx,
            int n_in, long ldy, int n_out, int batch)
{
    extern __shared__ __align__(16) double2 zsm[];
    const double *__restrict__ x = ptrs.x[blockIdx.y];
    double *__restrict__ y = ptrs.y[blockIdx.y];
    const int q0 = blockIdx.x * S;
    const int ns = min(S, batch - q0);
    const bool bwd = mode == PDE_DCT_BWD;
    const double se = bwd ? 0.5 : 1.0, so = bwd ? -0.5 : 1.0;       // scale of even / odd interior inputs
    constexpr int H = P / 2;
    constexpr int M1 = P / FirstRadix<RAD...>::value;
    constexpr int SP = SeqPad<P, M1, S, AXIS>::value;
    using PD = Pad<P, M1, SP>;
    constexpr int PS = PD::SEQ;

    // ---- load: z_m = (e_2m, e_2m+1); m < P/2: (x_2m, x_2m+1); m >= P/2: (x_{2P-2m}, x_{2P-2m-1})
    auto load_one = [&](int idx, int nseq, double &v0, double &v1, int &dst) {
        int s, m;
        if (S == 1) {
            s = 0;
            m = idx;
        } else if (AXIS == 1) {
            s = idx / P;
            m = idx - s * P;
        } else {
            m = idx / nseq;
            s = idx - m * nseq;
        }
        const int n0 = m < H ? 2 * m : 2 * P - 2 * m;          // even index
        const int n1 = m < H ? 2 * m + 1 : 2 * P - 2 * m - 1;  // odd index
        const double *src = AXIS == 1 ? x + (long)(q0 + s) * ldx : x + q0 + s;
        const long es = AXIS == 1 ? 1 : ldx;
        v0 = n0 < n_in ? src[n0 * es] : 0.0;
        v1 = n1 < n_in ? src[n1 * es] : 0.0;
        v0 *= (n0 == 0 || n0 == P) ? 1.0 : se;
        v1 *= so;
        dst = s * PS + PD::phys(m);
    };
    // Axis 1, one sequence per CTA, one first-pass butterfly per thread: the butterfly's 16 inputs
    // z_m, m = j + r M1, are read straight from global memory (lanes j -> consecutive 16-byte pairs: coalesced; the
    // zero-padded part of a backward transform is never read), so the load phase's shared-memory store and the first
    // pass's shared-memory load -- 2 of the kernel's ~10 sweeps over the 48 KB sequence -- and one barrier disappear.
    // MEASURED (B200, 2048 x 3073, backward): 0.060 ms fused vs 0.059 ms unfused -- the exposed global latency of the
    // butterfly's 32 loads eats what the saved shared-memory traffic gives, so this stays opt-in (-DPDE_FFT_FUSE1);
    // parity tests pass with it (156 GPU tests).
#ifndef PDE_FFT_FUSE1
    constexpr bool FUSE1 = false;
#else
    constexpr bool FUSE1 = AXIS == 1 && S == 1 && FirstRadix<RAD...>::value == 16 && T == P / 16 && sizeof...(RAD) >= 2;
#endif
    if constexpr (FUSE1) {
        constexpr int R = 16, M = P / R;
        const int j = threadIdx.x;
        const double *src = x + (long)q0 * ldx;
        const bool vec = ((ldx & 1) == 0) && (((unsigned long long)x & 15) == 0);
        double2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int m = j + r * M;
            double v0, v1;
            if (r < R / 2) {                                   // m < P/2: (x_2m, x_2m+1)
                const int n0 = 2 * m;
                if (vec && n0 + 1 < n_in) {
                    const double2 t = __ldg(reinterpret_cast<const double2 *>(src + n0));
                    v0 = t.x;
                    v1 = t.y;
                } else {
                    v0 = n0 < n_in ? __ldg(src + n0) : 0.0;
                    v1 = n0 + 1 < n_in ? __ldg(src + n0 + 1) : 0.0;
                }
            } else {                                           // m >= P/2: (x_{2P-2m}, x_{2P-2m-1})
                const int n0 = 2 * P - 2 * m;
                v0 = n0 < n_in ? __ldg(src + n0) : 0.0;
                v1 = n0 - 1 < n_in ? __ldg(src + n0 - 1) : 0.0;
            }
            // the end points x_0 (m = 0) and x_P (m = P/2, even slot) are not scaled
            const bool edge = (r == 0 || r == R / 2) && j == 0;
            a[r] = make_double2(v0 * (edge ? 1.0 : se), v1 * so);
        }
        dft<R>(a);
#pragma unroll
        for (int r = 1; r < R; ++r) a[r] = cmul(a[r], __ldg(W + (r - 1) * M + j));
        double2 *p = zsm + j;
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * (M + 1)] = a[r];
        __syncthreads();
        TailPasses<RAD...>::template run<SP, P, M1, S, T>(zsm, W);
    } else {
    if (ns == S) {
        // Full CTA.  The iteration index is a compile-time constant and the thread index enters through
        // per-thread invariants, so the (sequence, element) split costs nothing per element (the generic
        // form below spent 44 instructions per element, 29 % of the kernel's instruction stream, on index
        // arithmetic).  Loads are issued in batches of UB iterations before their first use.
        static_assert((S * P) % T == 0, "S * P must be a multiple of T");
        constexpr int ITER = S * P / T;
                constexpr int UB = ITER < 16 ? ITER : (AXIS == 1 ? 16 : 8);      // (16 on axis 0: measured equal)
        static_assert(ITER % UB == 0, "load batches");
        static_assert(AXIS == 1 ? (P % T == 0 || T % P == 0) : (T % S == 0 && P % (T / S) == 0), "thread mapping");
        constexpr int C = AXIS == 1 ? (P % T == 0 ? T : P) : T / S;     // elements of one sequence per iteration
        constexpr int SPI = AXIS == 1 ? (P % T == 0 ? 0 : T / P) : 0;   // axis 1, short sequences: sequences per iteration
        const int ld32 = (int)ldx;
        // per-thread invariants
        const int tm = AXIS == 1 ? (int)threadIdx.x % C : (int)threadIdx.x / S;      // element offset within the iteration
        const int ts = AXIS == 1 ? (int)threadIdx.x / C : (int)threadIdx.x % S;      // sequence offset
        const double *tsrc = AXIS == 1 ? x + (long)(q0 + ts) * ld32 : x + q0 + ts;
        double2 *tz = zsm + ts * PS;
        static_for<0, ITER / UB>([&](auto bc) {
            constexpr int B0 = decltype(bc)::value * UB;
            double v0[UB], v1[UB];
            static_for<0, UB>([&](auto uc) {
                constexpr int u = decltype(uc)::value, it = B0 + u;
                // axis 1, P % T == 0: sequence it / (P / T), element (it % (P / T)) * T + tm
                // axis 1, T % P == 0: sequence it * SPI + ts,  element tm
                // axis 0:             sequence ts,             element it * C + tm
                constexpr int sc = AXIS == 1 ? (P % T == 0 ? it / (P / T) : it * SPI) : 0;
                constexpr int mc = AXIS == 1 ? (P % T == 0 ? (it % (P / T)) * T : 0) : it * C;
                const int m = mc + tm;
                const bool lo = (mc + C <= H) ? true : (mc >= H ? false : m < H);
                const int n0 = lo ? 2 * m : 2 * P - 2 * m;
                const int n1 = lo ? n0 + 1 : n0 - 1;
                const double *src = AXIS == 1 ? tsrc + (long)sc * ld32 : tsrc;
                if (AXIS == 1) {
                    v0[u] = n0 < n_in ? __ldg(src + n0) : 0.0;
                    v1[u] = n1 < n_in ? __ldg(src + n1) : 0.0;
                } else {
                    v0[u] = n0 < n_in ? __ldg(src + (long)n0 * ld32) : 0.0;
                    v1[u] = n1 < n_in ? __ldg(src + (long)n1 * ld32) : 0.0;
                }
            });
            static_for<0, UB>([&](auto uc) {
                constexpr int u = decltype(uc)::value, it = B0 + u;
                constexpr int sc = AXIS == 1 ? (P % T == 0 ? it / (P / T) : it * SPI) : 0;
                constexpr int mc = AXIS == 1 ? (P % T == 0 ? (it % (P / T)) * T : 0) : it * C;
                const int m = mc + tm;
                // the end points x_0 (m = 0, even slot) and x_P (m = H, even slot) are not scaled
                const bool edge = (mc == 0 && tm == 0) || (mc <= H && H < mc + C && m == H);
                tz[sc * PS + PD::phys(m)] = make_double2(v0[u] * (edge ? 1.0 : se), v1[u] * so);
            });
        });
    } else {
        for (int idx = threadIdx.x; idx < ns * P; idx += T) {
            double v0, v1;
            int dst;
            load_one(idx, ns, v0, v1, dst);
            zsm[dst] = make_double2(v0, v1);
        }
    }
    __syncthreads();
    DifPassesW<SP, 0, P, M1, P, S, T, RAD...>::run(zsm, W);      // W: per-pass [r][j] tables (p->Wp)
    }

    // ---- split + store
    const double fs = 1.0 / (2.0 * (double)P);
    const bool fwd = mode == PDE_DCT_FWD;
    constexpr int R1 = FirstRadix<RAD...>::value;
    // outputs k and P - k of sequence s (k <= H); y_k = (A_k + conj(A_{P-k}))/2 - i e^{-i pi k / P} (A_k - conj(A_{P-k}))/2
    auto split_one = [&](int sq, int k, const double2 *zq, double *dst) {
        const int k2 = P - k;
        const int kk = k == 0 ? 0 : k2;
        // digit-reversed position; its first-level block number is the lowest digit, so phys() needs no division
        const double2 a = zq[DigitRev<P, RAD...>::pos(k) + k % R1];
        const double2 b = zq[DigitRev<P, RAD...>::pos(kk) + kk % R1];
        const double2 cs = __ldg(CS + k);
        const double sr = a.x + b.x, dr = a.x - b.x, si = a.y + b.y;
        // forward scale: 1/(2P) at the ends, +-1/P inside (the sign is (-1)^k, and P - k has the parity of k)
        const double h = fwd ? (k == 0 ? 0.5 * fs : ((k & 1) ? -fs : fs)) : 0.5;
        const double t = cs.x * si - cs.y * dr;
        const double yk = h * (sr + t), yk2 = h * (sr - t);
        if (AXIS == 1) {
            if (k < n_out) dst[k] = yk;
            if (k2 != k && k2 < n_out) dst[k2] = yk2;
        } else {
            if (k < n_out) dst[(long)k * (int)ldy] = yk;
            if (k2 != k && k2 < n_out) dst[(long)k2 * (int)ldy] = yk2;
        }
        (void)sq;
    };
    if (ns == S) {
        static_assert((S * H) % T == 0, "S * P / 2 must be a multiple of T");
        static_assert(AXIS == 1 ? (H % T == 0 || T % H == 0) : (H % (T / S) == 0), "thread mapping (split)");
        constexpr int ITER = S * H / T;
        constexpr int C = AXIS == 1 ? (H % T == 0 ? T : H) : T / S;
        constexpr int SPI = AXIS == 1 ? (H % T == 0 ? 0 : T / H) : 0;
        const int tk = AXIS == 1 ? (int)threadIdx.x % C : (int)threadIdx.x / S;
        const int ts = AXIS == 1 ? (int)threadIdx.x / C : (int)threadIdx.x % S;
        const double2 *tz = zsm + ts * PS;
        double *tdst = AXIS == 1 ? y + (long)(q0 + ts) * (int)ldy : y + q0 + ts;
        static_for<0, ITER>([&](auto ic) {
            constexpr int it = decltype(ic)::value;
            constexpr int sc = AXIS == 1 ? (H % T == 0 ? it / (H / T) : it * SPI) : 0;
            constexpr int kc = AXIS == 1 ? (H % T == 0 ? (it % (H / T)) * T : 0) : it * C;
            split_one(sc, kc + tk, tz + sc * PS, AXIS == 1 ? tdst + (long)sc * (int)ldy : tdst);
        });
        if (threadIdx.x < S) {                      // k = H (its partner is itself)
            const int sq = threadIdx.x;
            split_one(sq, H, zsm + sq * PS, AXIS == 1 ? y + (long)(q0 + sq) * (int)ldy : y + q0 + sq);
        }
    } else {
        for (int idx = threadIdx.x; idx < ns * (H + 1); idx += T) {
            int sq, k;
            if (AXIS == 1) {
                sq = idx / (H + 1);
                k = idx - sq * (H + 1);
            } else {
                k = idx / ns;
                sq = idx - k * ns;
            }
            split_one(sq, k, zsm + sq * PS, AXIS == 1 ? y + (long)(q0 + sq) * (int)ldy : y + q0 + sq);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Axis 1 (contiguous rows), one sequence per CTA pass, PERSISTENT CTAs with a TMA-prefetched input row.
//
// ncu on k_dct_fft_t<3072, 1, 192, 1, ...> (profiles/r02_ncu_stage.csv): L1TEX 83 % busy, ~5500 wavefronts per
// sequence -- 3400 shared-memory wavefronts of the four z sweeps, the rest global loads (input row, twiddles) that
// share the same pipe -- and 4.4 long-scoreboard stalls per issue: every CTA waits for its own row before it can
// start.  Here a CTA loops over rows; while it runs passes 2.. and the split of row i, the copy engine
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx: no LSU wavefronts, no registers) brings row i+1 into
// a 24 KB staging buffer.  The first radix-16 pass reads its 16 inputs straight from that buffer (real data: half the
// bytes of the packed z, and the zero-padded part of a backward transform is never read), so the load phase's
// z store and the first pass's z load disappear: 6.5 instead of 8 shared sweeps, no exposed global latency.
// Needs 16-byte aligned rows (even ldx); 49.4 + 24.6 KB of shared memory per CTA = 3 CTAs per SM as before.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem, const void *gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int P, int T, int... RAD>
__global__ void __maxnreg__((FftRegs<P, 1, T, 16>::value))
k_dct_row_tma(const double2 *__restrict__ W, const double2 *__restrict__ CS, int mode, const __grid_constant__ DctPtrs ptrs,
              long ldx, int n_in,
              long ldy, int n_out, int batch, int nitems)
{
    static_assert(FirstRadix<RAD...>::value == 16 && T == P / 16, "one radix-16 first-pass butterfly per thread");
    extern __shared__ __align__(16) double2 zsm[];
    constexpr int R = 16, M = P / R, H = P / 2;
    using PD = Pad<P, M, 0>;
    constexpr int PS = PD::SEQ, RAWN = P + 2;
    double *raw = reinterpret_cast<double *>(zsm + PS);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(raw + RAWN);
    const bool bwd = mode == PDE_DCT_BWD, fwd = mode == PDE_DCT_FWD;
    const double se = bwd ? 0.5 : 1.0, so = bwd ? -0.5 : 1.0;
    const int j = threadIdx.x;
    const unsigned even_bytes = (unsigned)(n_in & ~1) * 8u;

    auto issue = [&](int item) {              // thread 0: start the copy of one row
        const int job = item / batch, q = item - job * batch;
        const double *src = ptrs.x[job] + (long)q * ldx;
        mbar_expect_tx(bar, even_bytes);
        if (even_bytes) bulk_load(raw, src, even_bytes, bar);
        if (n_in & 1) {                        // odd tail element: the bulk copy moves multiples of 16 bytes
            cp_async8_fft(raw + n_in - 1, src + n_in - 1);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if (j == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        if ((int)blockIdx.x < nitems) issue(blockIdx.x);
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    }
    __syncthreads();

    unsigned phase = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, phase ^= 1u) {
        const int job = item / batch, q = item - job * batch;
        double *__restrict__ yrow = ptrs.y[job] + (long)q * ldy;
        auto ldw1 = [&](int r) { return __ldg(W + (r - 1) * M + j); };
        double2 pre1[TwPre<R>::N > 0 ? TwPre<R>::N : 1];       // requested while the row is still in flight
#pragma unroll
        for (int r = 1; r <= TwPre<R>::N; ++r) pre1[r - 1] = ldw1(r);
        mbar_wait(bar, phase);
        // ---- first pass: z_m = (e_2m, e_2m+1), m = j + r M; m < H: (x_2m, x_2m+1); m >= H: (x_{2P-2m}, x_{2P-2m-1})
        double2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int m = j + r * M;
            double v0, v1;
            if (r < R / 2) {
                const int n0 = 2 * m;
                if (n0 + 1 < n_in) {
                    const double2 t = *reinterpret_cast<const double2 *>(raw + n0);
                    v0 = t.x;
                    v1 = t.y;
                } else {
                    v0 = n0 < n_in ? raw[n0] : 0.0;
                    v1 = 0.0;
                }
            } else {
                const int n0 = 2 * P - 2 * m;
                v0 = n0 < n_in ? raw[n0] : 0.0;
                v1 = n0 - 1 < n_in ? raw[n0 - 1] : 0.0;
            }
            const bool edge = (r == 0 || r == R / 2) && j == 0;        // x_0 and x_P are not scaled
            a[r] = make_double2(v0 * (edge ? 1.0 : se), v1 * so);
        }
        dft<R>(a);
        if constexpr (TwPre<R>::N > 0) twiddle_mul_pre<R, TwPre<R>::N>(a, pre1, ldw1);
        else twiddle_mul<R>(a, ldw1);
        {
            double2 *p = zsm + j;
#pragma unroll
            for (int r = 0; r < R; ++r) p[r * (M + 1)] = a[r];
        }
        __syncthreads();                       // z complete; the staging buffer is free
        const int next = item + gridDim.x;
        if (j == 0 && next < nitems) issue(next);
        TailPasses<RAD...>::template run<0, P, M, 1, T>(zsm, W);
        // ---- split + store: y_k = (A_k + conj(A_{P-k}))/2 - i e^{-i pi k / P} (A_k - conj(A_{P-k}))/2
        const double fs = 1.0 / (2.0 * (double)P);
        auto split_one = [&](int k) {
            const int k2 = P - k;
            const int kk = k == 0 ? 0 : k2;
            const double2 av = zsm[DigitRev<P, RAD...>::pos(k) + k % R];
            const double2 bv = zsm[DigitRev<P, RAD...>::pos(kk) + kk % R];
            const double2 cs = __ldg(CS + k);
            const double sr = av.x + bv.x, dr = av.x - bv.x, si = av.y + bv.y;
            const double h = fwd ? (k == 0 ? 0.5 * fs : ((k & 1) ? -fs : fs)) : 0.5;
            const double t = cs.x * si - cs.y * dr;
            if (k < n_out) yrow[k] = h * (sr + t);
            if (k2 != k && k2 < n_out) yrow[k2] = h * (sr - t);
        };
        static_assert(H % T == 0, "split mapping");
#pragma unroll
        for (int it = 0; it < H / T; ++it) split_one(it * T + j);
        if (j == 0) {
            split_one(H);
            asm volatile("cp.async.wait_all;\n" ::: "memory");      // tail element of the next row (issued long ago)
        }
        __syncthreads();                       // z free for the next row
    }
}

template <int P, int T, int... RAD>
static int launch_row_tma(const FftDctPlan *p, int mode, int njobs, const DctPtrs &ptrs, long ldx, int n_in, long ldy,
                          int n_out, int batch, cudaStream_t st)
{
    auto kern = k_dct_row_tma<P, T, RAD...>;
    constexpr int M1 = P / 16;
    constexpr size_t smem = (size_t)Pad<P, M1, 0>::SEQ * 16 + (size_t)(P + 2) * 8 + 16;
    if (!p->Wp) {
        std::vector<double2> tab;
        build_pass_tables<RAD...>(P, P, tab);
        double2 *d = nullptr;
        PDE_CUDA(cudaMalloc(&d, sizeof(double2) * tab.size()));
        PDE_CUDA(cudaMemcpy(d, tab.data(), sizeof(double2) * tab.size(), cudaMemcpyHostToDevice));
        const_cast<FftDctPlan *>(p)->Wp = d;
    }
    constexpr int per_cta = (int)smem + 1024;
    constexpr int fit = 227 * 1024 / per_cta;
    constexpr int ctas = fit < 3 ? (fit < 1 ? 1 : fit) : 3;
    static PerDeviceFlag attr;
    if (!attr.get()) {
        // all of the 228 KB as shared memory: three 74 KB CTAs.  (Measured: a shorter staging buffer for the
        // zero-padded backward transforms, which fits the 196 KB configuration and leaves 60 KB of L1 for the
        // twiddle tables, runs at the same speed -- and a carve-out hint of 85 % silently drops to two CTAs.)
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_dct_row_tma): %s", cudaGetErrorString(e));
            return PDE_ERR_CUDA;
        }
        attr.get() = true;
    }
    const long nitems = (long)njobs * batch;
    const int grid = (int)std::min<long>(nitems, (long)ctas * sm_count());
    kern<<<grid, T, smem, st>>>(p->Wp, p->CS, mode, ptrs, ldx, n_in, ldy, n_out, batch, (int)nitems);
    return after_launch("pde_dct1(fft, persistent rows)");
}

// rows through the persistent TMA kernel: 16-byte aligned rows, a specialised length; PDE_DCT_TMA=0 disables
static int dispatch_row_tma(const FftDctPlan *p, int mode, int njobs, const DctPtrs &ptrs, long ldx, int n_in, long ldy,
                            int n_out, int batch, cudaStream_t st)
{
    static const bool enabled = !(getenv("PDE_DCT_TMA") && atoi(getenv("PDE_DCT_TMA")) == 0);
    if (!enabled || (ldx & 1) || n_in > p->P + 1) return -1;
    for (int jn = 0; jn < njobs; ++jn)
        if ((unsigned long long)ptrs.x[jn] & 15) return -1;
    switch (p->P) {
    case 3072: return launch_row_tma<3072, 192, 16, 16, 12>(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
    case 2048: return launch_row_tma<2048, 128, 16, 16, 8>(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
    case 4096: return launch_row_tma<4096, 256, 16, 16, 16>(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
    default: return -1;
    }
}

template <int P, int S, int T, int AXIS, int... RAD>
static int launch_fft_t(const FftDctPlan *p, int mode, int njobs, const DctPtrs &ptrs, long ldx, int n_in, long ldy,
                        int n_out, int batch, cudaStream_t st)
{
    auto kern = k_dct_fft_t<P, S, T, AXIS, RAD...>;
    constexpr int M1 = P / FirstRadix<RAD...>::value;
    constexpr size_t smem = (size_t)S * Pad<P, M1, SeqPad<P, M1, S, AXIS>::value>::SEQ * 16;
    if (!p->Wp) {
        std::vector<double2> tab;
        build_pass_tables<RAD...>(P, P, tab);
        double2 *d = nullptr;
        PDE_CUDA(cudaMalloc(&d, sizeof(double2) * tab.size()));
        PDE_CUDA(cudaMemcpy(d, tab.data(), sizeof(double2) * tab.size(), cudaMemcpyHostToDevice));
        const_cast<FftDctPlan *>(p)->Wp = d;
    }
    static PerDeviceFlag attr;
    if (!attr.get()) {
        // Shared-memory carve-out: room for up to 3 CTAs, and never more than needed -- what is left of the
        // 256 KB is L1, which has to hold the twiddle tables (16 P bytes).  The S = 4 axis-0 CTAs need
        // 193 KB + 1 KB: asking for 100 % left a 28 KB L1 that thrashed (0.31 ms); the 196 KB configuration
        // leaves 60 KB.
        {
            const int per_cta = (int)smem + 1024;
            const int fit = 227 * 1024 / per_cta;
            const int want = per_cta * (fit < 3 ? fit : 3);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, want * 100 / (228 * 1024));
        }
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_dct_fft_t): %s", cudaGetErrorString(e));
            return PDE_ERR_CUDA;
        }
        attr.get() = true;
    }
    // (Tried: launching the axis-0 CTAs as clusters of 2/4/8 so that the four 32-byte sectors of a line are
    // requested close in time, and other shared-memory carve-outs: no effect once the row pitch of the arrays
    // is not a multiple of 2 KB.  With a 16 KB pitch the axis-0 time is erratic, 0.2 .. 0.66 ms.)
    dim3 grid(ceil_div(batch, S), njobs);
    kern<<<grid, T, smem, st>>>(p->Wp, p->CS, mode, ptrs, ldx, n_in, ldy, n_out, batch);
    return after_launch("pde_dct1(fft, specialised)");
}

// Dispatch of the specialised sizes (P = 3 * 2^k: the FFT-friendly 3/2-rule grids of N = 2^k,
// and powers of two).  S * P = 3072 points per CTA (48 KB), 192 threads: one radix-16
// butterfly per thread and pass.  Returns -1 when P has no specialisation.
template <int AXIS>
static int dispatch_fft_t(const FftDctPlan *p, int mode, int njobs, const DctPtrs &ptrs, long ldx, int n_in,
                          long ldy, int n_out, int batch, cudaStream_t st)
{
#define PDE_FFT_CASE(PP, SS, TT, ...)                                                                     \
    case PP: return launch_fft_t<PP, SS, TT, AXIS, __VA_ARGS__>(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
    // 3 * 2^k with a final radix-12 pass (three shared-memory round trips instead of four); PDE_FFT_RADIX12=0
    // selects the radix-4 + radix-3 tail for comparison
    static const bool radix12 = !(getenv("PDE_FFT_RADIX12") && atoi(getenv("PDE_FFT_RADIX12")) == 0);
    if (radix12) {
        // PDE_FFT_AXIS0_S=2: two columns per CTA (99 KB, 2 CTAs per SM, 16-byte row pieces) instead of four (197 KB, 1 CTA)
        static const int s0 = getenv("PDE_FFT_AXIS0_S") ? atoi(getenv("PDE_FFT_AXIS0_S")) : 4;
        if (AXIS == 0 && s0 == 2 && p->P == 3072)
            return launch_fft_t<3072, 2, 192, AXIS, 16, 16, 12>(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
        if (AXIS == 0) {
            switch (p->P) {
                PDE_FFT_CASE(1536, 4, 384, 16, 8, 12)
                PDE_FFT_CASE(3072, 4, 384, 16, 16, 12)
            default: break;
            }
        }
        switch (p->P) {
            PDE_FFT_CASE(1536, 2, 192, 16, 8, 12)
            PDE_FFT_CASE(3072, 1, 192, 16, 16, 12)
        default: break;
        }
    }
    if (AXIS == 0) {
        // strided sequences: >= 4 adjacent columns per CTA so that every global access is a full
        // 32-byte sector (S = 1 measured 0.25 ms vs 0.14 ms for the contiguous axis at P = 3072)
        switch (p->P) {
            PDE_FFT_CASE(1536, 4, 384, 16, 16, 2, 3)
            PDE_FFT_CASE(3072, 4, 384, 16, 16, 4, 3)
            PDE_FFT_CASE(6144, 2, 384, 16, 16, 8, 3)
            PDE_FFT_CASE(2048, 4, 256, 16, 16, 8)
            PDE_FFT_CASE(4096, 2, 256, 16, 16, 16)
        default: break;
        }
    }
    switch (p->P) {
        PDE_FFT_CASE(96, 32, 192, 16, 2, 3)
        PDE_FFT_CASE(192, 16, 192, 16, 4, 3)
        PDE_FFT_CASE(384, 8, 192, 16, 8, 3)
        PDE_FFT_CASE(768, 4, 192, 16, 16, 3)
        PDE_FFT_CASE(1536, 2, 192, 16, 16, 2, 3)
        PDE_FFT_CASE(3072, 1, 192, 16, 16, 4, 3)
        PDE_FFT_CASE(6144, 1, 384, 16, 16, 8, 3)
        PDE_FFT_CASE(256, 8, 128, 16, 16)
        PDE_FFT_CASE(512, 4, 128, 16, 16, 2)
        PDE_FFT_CASE(1024, 2, 128, 16, 16, 4)
        PDE_FFT_CASE(2048, 1, 128, 16, 16, 8)
        PDE_FFT_CASE(4096, 1, 256, 16, 16, 16)
    default: return -1;
    }
#undef PDE_FFT_CASE
}

}  // namespace pde
