// Bluestein (chirp-z) DCT-I for lengths L = P+1 whose P is odd or has large prime factors
// (N = 2^k grid points: P = 2047 = 23*89, 4095, 127 prime, ...).  Included by dct_fft.cu.
//
// The complex DFT of length P that the DCT-I needs (z_m = e_2m + i e_2m+1, see dct_fft.cu) is written
// as a convolution, Y_k = conj(c_k) sum_m (z_m conj(c_m)) c_{k-m}, c_j = exp(i pi j^2 / P), and
// evaluated with power-of-two FFTs of length M >= 2P-1 held in shared memory:
//     a = z conj(c) zero padded -> forward DIF FFT_M (digit-reversed output)
//       -> times Bhat (FFT of the wrapped chirp, pre-permuted to digit-reversed order, 1/M folded in)
//       -> inverse DIT passes (exact inverse of the DIF passes: natural-order output)
//       -> times conj(c_k) -> real split of the even extension -> DCT-I.
// Chirp phases use j^2 mod 2P in exact integer arithmetic and long-double sin/cos on the host.
#pragma once

namespace pde {

// inverse of dif_pass_t: x = IDFT_R( conj(W^{j r}) y_r ) (the 1/R factors are folded into Bhat)
template <int WOFF, int P, int M1, int NCUR, int S, int T, int R>
__device__ __forceinline__ void dit_pass_t(double2 *z, const double2 *__restrict__ W)
{
    constexpr int M = NCUR / R, PER_SEQ = P / R, TOTAL = S * PER_SEQ, TWS = P / NCUR;
    constexpr int PS = Pad<P, M1>::SEQ;
    constexpr int RS = NCUR == P ? M + 1 : M;
#pragma unroll
    for (int b0 = 0; b0 < TOTAL; b0 += T) {
        const int b = b0 + threadIdx.x;
        if ((TOTAL % T) != 0 && b >= TOTAL) break;
        const int s = b / PER_SEQ;
        const int bb = b - s * PER_SEQ;
        // same conflict-free thread -> butterfly map as dif_pass_t (dct_fft_t.cuh): consecutive butterflies walk
        // the first-level blocks (pitch M1 + 1), j is the slow index.  With j fastest the last-level pass
        // (NCUR = 16: addresses 16 b + r) was an 8-way bank conflict (tools/sim_fft_banks.py).
        constexpr int R1 = P / M1, NBLK = P / NCUR;
        int j, off;
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
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = p[r * RS];
        if (M > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                // per-pass [r][j] tables (same entries as W[j r TWS]; coalesced / broadcast instead of strided)
                const double2 w = WOFF < 0 ? __ldg(W + j * (r * TWS)) : __ldg(W + WOFF + (r - 1) * M + j);
                a[r] = make_double2(a[r].x * w.x + a[r].y * w.y, a[r].y * w.x - a[r].x * w.y);   // * conj(w)
            }
        }
        // IDFT = swap(DFT(swap(.)))
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = make_double2(a[r].y, a[r].x);
        dft<R>(a);
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * RS] = make_double2(a[r].y, a[r].x);
    }
}

// passes in reverse order of DifPasses<P, M1, P, S, T, RAD...>
template <int WOFF, int P, int M1, int NCUR, int S, int T, int R, int... Rest>
struct DitPasses {
    __device__ __forceinline__ static void run(double2 *z, const double2 *__restrict__ W)
    {
        constexpr int NEXT = WOFF < 0 ? -1 : WOFF + (NCUR / R > 1 ? (R - 1) * (NCUR / R) : 0);   // as DifPassesW
        if constexpr (sizeof...(Rest) > 0) DitPasses<NEXT, P, M1, NCUR / R, S, T, Rest...>::run(z, W);
        dit_pass_t<WOFF, P, M1, NCUR, S, T, R>(z, W);
        __syncthreads();
    }
};

struct BluesteinTables {
    const double2 *W;       // exp(-2 pi i j / M), j < M
    const double2 *chirp;   // c_m = exp(i pi m^2 / P), m < P
    const double2 *Bhat;    // FFT_M of the wrapped chirp / M, in digit-reversed (DIF output) order
    const double2 *CS;      // (cos, sin)(pi k / P), k <= (P+1)/2
    const double2 *Wp;      // per-pass [r][j] twiddle tables of the length-M FFT (built at the first launch)
};

template <int M, int S, int T, int AXIS, int... RAD>
__global__ void __launch_bounds__(T)
k_dct_bluestein(BluesteinTables tb, int P, int mode, DctPtrs ptrs, long ldx, int n_in, long ldy, int n_out, int batch)
{
    extern __shared__ __align__(16) double2 zsm[];
    const double *__restrict__ x = ptrs.x[blockIdx.y];
    double *__restrict__ y = ptrs.y[blockIdx.y];
    const int q0 = blockIdx.x * S;
    const int ns = min(S, batch - q0);
    const bool bwd = mode == PDE_DCT_BWD;
    const double se = bwd ? 0.5 : 1.0, so = bwd ? -0.5 : 1.0;
    const double sP = (bwd && (P & 1)) ? -1.0 : 1.0;     // (-1)^P c_P: the last coefficient's sign for odd P
    constexpr int M1 = M / FirstRadix<RAD...>::value;
    using PD = Pad<M, M1>;
    constexpr int PS = PD::SEQ;

    // ---- load: a_m = z_m conj(c_m), z_m = (e_2m, e_2m+1) of the even extension (period 2P); zero pad to M
    for (int idx = threadIdx.x; idx < ns * M; idx += T) {
        int s, m;
        if (AXIS == 1) {
            s = idx / M;
            m = idx - s * M;
        } else {
            m = idx / ns;
            s = idx - m * ns;
        }
        double2 a = make_double2(0.0, 0.0);
        if (m < P) {
            const int j0 = 2 * m, j1 = 2 * m + 1;
            const int n0 = j0 <= P ? j0 : 2 * P - j0;
            const int n1 = j1 <= P ? j1 : 2 * P - j1;
            const double *src = AXIS == 1 ? x + (long)(q0 + s) * ldx : x + q0 + s;
            const long es = AXIS == 1 ? 1 : ldx;
            double v0 = n0 < n_in ? src[n0 * es] : 0.0;
            double v1 = n1 < n_in ? src[n1 * es] : 0.0;
            // input scaling of the backward Chebyshev transform: ends x1, interior 0.5 (-1)^n
            v0 *= n0 == 0 ? 1.0 : (n0 == P ? sP : ((n0 & 1) ? so : se));
            v1 *= n1 == 0 ? 1.0 : (n1 == P ? sP : ((n1 & 1) ? so : se));
            const double2 c = __ldg(tb.chirp + m);
            a = make_double2(v0 * c.x + v1 * c.y, v1 * c.x - v0 * c.y);          // z * conj(c)
        }
        zsm[s * PS + PD::phys(m)] = a;
    }
    __syncthreads();
    DifPassesW<0, 0, M, M1, M, S, T, RAD...>::run(zsm, tb.Wp);
    // ---- pointwise product with the chirp spectrum (digit-reversed order on both sides)
    for (int idx = threadIdx.x; idx < S * M; idx += T) {
        const int s = idx / M, i = idx - s * M;
        const double2 b = __ldg(tb.Bhat + i);
        double2 *p = zsm + s * PS + PD::phys(i);
        *p = cmul(*p, b);
    }
    __syncthreads();
    DitPasses<0, M, M1, M, S, T, RAD...>::run(zsm, tb.Wp);

    // ---- Z_k = conv_k conj(c_k); real split of the even extension; store
    const int half = (P + 1) / 2;                 // pairs (k, P-k), k = 0 .. half (k = 0 pairs with P)
    const double fs = 1.0 / (2.0 * (double)P);
    const bool fwd = mode == PDE_DCT_FWD;
    auto Z = [&](int s, int k) {
        const double2 v = zsm[s * PS + PD::phys(k)];
        const double2 c = __ldg(tb.chirp + k);
        return make_double2(v.x * c.x + v.y * c.y, v.y * c.x - v.x * c.y);
    };
    for (int idx = threadIdx.x; idx < ns * (half + 1); idx += T) {
        int s, k;
        if (AXIS == 1) {
            s = idx / (half + 1);
            k = idx - s * (half + 1);
        } else {
            k = idx / ns;
            s = idx - k * ns;
        }
        const int k2 = P - k;
        if (k > k2) continue;
        const double2 a = Z(s, k);
        const double2 b = Z(s, k == 0 ? 0 : k2);
        const double2 cs = __ldg(tb.CS + k);
        const double sr = a.x + b.x, dr = a.x - b.x, si = a.y + b.y;
        double yk = 0.5 * (sr + cs.x * si - cs.y * dr);
        double yk2 = 0.5 * (sr - cs.x * si + cs.y * dr);
        if (fwd) {
            yk *= (k == 0 ? fs : ((k & 1) ? -2.0 * fs : 2.0 * fs));
            yk2 *= (k2 == P ? ((P & 1) ? -fs : fs) : ((k2 & 1) ? -2.0 * fs : 2.0 * fs));
        }
        double *dst = AXIS == 1 ? y + (long)(q0 + s) * ldy : y + q0 + s;
        const long ds = AXIS == 1 ? 1 : ldy;
        if (k < n_out) dst[k * ds] = yk;
        if (k2 != k && k2 < n_out) dst[k2 * ds] = yk2;
    }
}

template <int M, int S, int T, int AXIS, int... RAD>
static int launch_bluestein(const BluesteinTables &tb, int P, int mode, int njobs, const DctPtrs &ptrs, long ldx,
                            int n_in, long ldy, int n_out, int batch, cudaStream_t st)
{
    auto kern = k_dct_bluestein<M, S, T, AXIS, RAD...>;
    constexpr int M1 = M / FirstRadix<RAD...>::value;
    constexpr size_t smem = (size_t)S * Pad<M, M1>::SEQ * 16;
    static PerDeviceFlag attr;
    if (!attr.get()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_dct_bluestein): %s", cudaGetErrorString(e));
            return PDE_ERR_CUDA;
        }
        attr.get() = true;
    }
    if (!tb.Wp) {
        std::vector<double2> tab;
        build_pass_tables<RAD...>(M, M, tab);
        double2 *d = nullptr;
        PDE_CUDA(cudaMalloc(&d, sizeof(double2) * tab.size()));
        PDE_CUDA(cudaMemcpy(d, tab.data(), sizeof(double2) * tab.size(), cudaMemcpyHostToDevice));
        const_cast<BluesteinTables &>(tb).Wp = d;
    }
    dim3 grid(ceil_div(batch, S), njobs);
    kern<<<grid, T, smem, st>>>(tb, P, mode, ptrs, ldx, n_in, ldy, n_out, batch);
    return after_launch("pde_dct1(bluestein)");
}

// convolution lengths: M = 256 .. 8192 (P <= 4096)
template <int AXIS>
static int dispatch_bluestein(int M, const BluesteinTables &tb, int P, int mode, int njobs, const DctPtrs &ptrs,
                              long ldx, int n_in, long ldy, int n_out, int batch, cudaStream_t st)
{
#define PDE_BS_CASE(MM, SS, TT, ...)                                                                        \
    case MM: return launch_bluestein<MM, SS, TT, AXIS, __VA_ARGS__>(tb, P, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
    constexpr int S0 = AXIS == 0 ? 2 : 1;      // strided axis: two adjacent columns per CTA where they fit
    switch (M) {
        PDE_BS_CASE(256, 8, 128, 16, 16)
        PDE_BS_CASE(512, 4, 128, 16, 16, 2)
        PDE_BS_CASE(1024, 2, 128, 16, 16, 4)
        PDE_BS_CASE(2048, S0, 128 * S0, 16, 16, 8)
        PDE_BS_CASE(4096, S0, 256 * S0, 16, 16, 16)
        PDE_BS_CASE(8192, 1, 512, 16, 16, 8, 4)
    default: return -1;
    }
#undef PDE_BS_CASE
}

}  // namespace pde
