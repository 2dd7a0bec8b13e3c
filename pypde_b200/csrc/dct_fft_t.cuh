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

// Shared-memory layout of one sequence: logical index i lives at i + i / M1 (M1 = P / first radix):
// one pad element after each first-level block.  The split step reads the digit-reversed FFT output
// with stride M1 between consecutive k; M1 is a multiple of 8 elements (= all banks), so without the
// pad those reads were 8-way bank conflicts (ncu: 48 % of the shared wavefronts).
template <int P, int M1>
struct Pad {
    static constexpr int SEQ = P + P / M1;                 // padded sequence length
    __device__ __forceinline__ static int phys(int i) { return i + i / M1; }
};

template <int P, int M1, int NCUR, int S, int T, int R>
__device__ __forceinline__ void dif_pass_t(double2 *z, const double2 *__restrict__ W)
{
    constexpr int M = NCUR / R, PER_SEQ = P / R, TOTAL = S * PER_SEQ, TWS = P / NCUR;
    constexpr int PS = Pad<P, M1>::SEQ;
    constexpr int RS = NCUR == P ? M + 1 : M;              // first pass: stride M1 + 1 (padded)
#pragma unroll
    for (int b0 = 0; b0 < TOTAL; b0 += T) {
        const int b = b0 + threadIdx.x;
        if ((TOTAL % T) != 0 && b >= TOTAL) break;
        const int s = b / PER_SEQ;
        const int bb = b - s * PER_SEQ;
        const int blk = bb / M;
        const int j = bb - blk * M;
        const int i0 = blk * NCUR + j;
        double2 *p = z + s * PS + (NCUR == P ? i0 : Pad<P, M1>::phys(i0));
        double2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = p[r * RS];
        dft<R>(a);
        if (M > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(a[r], __ldg(W + j * (r * TWS)));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * RS] = a[r];
    }
}

template <int P, int M1, int NCUR, int S, int T, int R, int... Rest>
struct DifPasses {
    __device__ __forceinline__ static void run(double2 *z, const double2 *__restrict__ W)
    {
        dif_pass_t<P, M1, NCUR, S, T, R>(z, W);
        __syncthreads();
        if constexpr (sizeof...(Rest) > 0) DifPasses<P, M1, NCUR / R, S, T, Rest...>::run(z, W);
    }
};

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
__global__ void __launch_bounds__(T)
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
    using PD = Pad<P, M1>;
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
    if (ns == S) {
        // full CTA: compile-time trip count, all global loads of a batch issued before the first use
        constexpr int ITER = (S * P + T - 1) / T;
        constexpr int UB = AXIS == 1 ? 16 : 8;                  // loads in flight per thread: 2 * UB (16 thrashes on axis 0)
#pragma unroll 1
        for (int it0 = 0; it0 < ITER; it0 += UB) {
            double v0[UB], v1[UB];
            int dst[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int idx = (it0 + u) * T + threadIdx.x;
                dst[u] = -1;
                if (it0 + u < ITER && idx < S * P) load_one(idx, S, v0[u], v1[u], dst[u]);
            }
#pragma unroll
            for (int u = 0; u < UB; ++u)
                if (dst[u] >= 0) zsm[dst[u]] = make_double2(v0[u], v1[u]);
        }
    } else {
        for (int idx = threadIdx.x; idx < ns * P; idx += T) {
            double v0, v1;
            int dst;
            load_one(idx, ns, v0, v1, dst);
            zsm[dst] = make_double2(v0, v1);
        }
    }
    __syncthreads();
    DifPasses<P, M1, P, S, T, RAD...>::run(zsm, W);

    // ---- split + store
    const double fs = 1.0 / (2.0 * (double)P);
    const bool fwd = mode == PDE_DCT_FWD;
    for (int idx = threadIdx.x; idx < ns * (H + 1); idx += T) {
        int s, k;
        if (S == 1) {
            s = 0;
            k = idx;
        } else if (AXIS == 1) {
            s = idx / (H + 1);
            k = idx - s * (H + 1);
        } else {
            k = idx / ns;
            s = idx - k * ns;
        }
        const int k2 = P - k;
        const double2 a = zsm[s * PS + PD::phys(DigitRev<P, RAD...>::pos(k))];
        const double2 b = zsm[s * PS + PD::phys(DigitRev<P, RAD...>::pos(k == 0 ? 0 : k2))];
        const double2 cs = __ldg(CS + k);
        const double sr = a.x + b.x, dr = a.x - b.x, si = a.y + b.y;
        double yk = 0.5 * (sr + cs.x * si - cs.y * dr);
        double yk2 = 0.5 * (sr - cs.x * si + cs.y * dr);
        if (fwd) {
            yk *= (k == 0 ? fs : ((k & 1) ? -2.0 * fs : 2.0 * fs));
            yk2 *= (k2 == P ? fs : ((k2 & 1) ? -2.0 * fs : 2.0 * fs));
        }
        double *dst = AXIS == 1 ? y + (long)(q0 + s) * ldy : y + q0 + s;
        const long ds = AXIS == 1 ? 1 : ldy;
        if (k < n_out) dst[k * ds] = yk;
        if (k2 != k && k2 < n_out) dst[k2 * ds] = yk2;
    }
}

template <int P, int S, int T, int AXIS, int... RAD>
static int launch_fft_t(const FftDctPlan *p, int mode, int njobs, const DctPtrs &ptrs, long ldx, int n_in, long ldy,
                        int n_out, int batch, cudaStream_t st)
{
    auto kern = k_dct_fft_t<P, S, T, AXIS, RAD...>;
    constexpr int M1 = P / FirstRadix<RAD...>::value;
    constexpr size_t smem = (size_t)S * Pad<P, M1>::SEQ * 16;
    static bool attr = false;
    if (!attr) {
        // enough shared memory for 4 CTAs (or as many as fit), the rest stays L1 for the twiddle table
        {
            const int want = (int)(smem + 1024) * 3;
            const int pct = want >= 227 * 1024 ? 100 : (want * 100 + 227 * 1024 - 1) / (227 * 1024);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_dct_fft_t): %s", cudaGetErrorString(e));
            return PDE_ERR_CUDA;
        }
        attr = true;
    }
    dim3 grid(ceil_div(batch, S), njobs);
    kern<<<grid, T, smem, st>>>(p->W, p->CS, mode, ptrs, ldx, n_in, ldy, n_out, batch);
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
