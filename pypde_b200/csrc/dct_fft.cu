// Shared-memory FFT DCT-I for sm_100a.
//
// DCT-I of length L = P+1 (P even, P = 2^a 3^b 5^c) through the real FFT of the even
// extension e (length 2P), computed with ONE complex FFT of length P per sequence
// (z_m = e_2m + i e_2m+1) and the usual real-FFT split, of which only the real part
// survives because e is even:
//     y_k = 1/2 [ (Zr_k + Zr_{P-k}) + cos(pi k/P) (Zi_k + Zi_{P-k}) - sin(pi k/P) (Zr_k - Zr_{P-k}) ]
// This is the numerically stable route (pocketfft / FFTW use the same 2(L-1) real FFT;
// the length-(L-1) "FFTPACK cost" shortcut loses O(L) digits).
//
// One CTA owns S whole sequences in shared memory (16 P bytes each).  The FFT is an
// in-place decimation-in-frequency mixed-radix transform (radix 2/3/4/5 butterflies in
// registers, one __syncthreads per pass, twiddles from an accurate host-built table);
// its output stays in digit-reversed order and the split step reads it through a
// permutation table, writing results straight to global memory:
//   axis 1: a sequence is a contiguous row    -> fully coalesced loads/stores; 16-byte aligned rows of the hot
//           lengths go through k_dct_row_tma (dct_fft_t.cuh): persistent CTAs, the next row staged by the
//           bulk-copy engine (cp.async.bulk + mbarrier) while the current one is transformed;
//   axis 0: a sequence is a strided column    -> the CTA takes S adjacent columns, so every
//           access is an S*8-byte segment (32 B sectors fully used for S = 4).
// Zero padding (n_in < L), truncation (n_out < L) and the Chebyshev scale/sign/mass
// factors are folded into the load / store phases (no extra passes over HBM).
#include "common.cuh"
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstdlib>

namespace pde {

struct FftDctPlan {
    int L = 0, P = 0;
    int npass = 0;
    int radix[24];
    double2 *W = nullptr;     // W[j]  = exp(-2 pi i j / P), j < P   (Bluestein: j < M)
    double2 *Wp = nullptr;    // per-pass [r][j] twiddle tables of the specialised kernels (built on first use)
    double2 *CS = nullptr;    // CS[k] = (cos(pi k/P), sin(pi k/P)), k <= P/2
    int *pos = nullptr;       // digit-reversed position of output k
    int M = 0;                // Bluestein convolution length (0: direct FFT of length P)
    double2 *chirp = nullptr, *Bhat = nullptr;
};

struct PassList {
    int npass;
    int radix[24];
};

struct DctPtrs {                 // blockIdx.y selects the array (pde_dct1_multi)
    const double *x[PDE_MAX_JOBS];
    double *y[PDE_MAX_JOBS];
};

// ---- small DFTs (forward, exp(-2 pi i qr/R)) -------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)

template <int R>
__device__ __forceinline__ void dft(double2 *a);

template <>
__device__ __forceinline__ void dft<2>(double2 *a)
{
    double2 t = a[0];
    a[0] = cadd(t, a[1]);
    a[1] = csub(t, a[1]);
}

template <>
__device__ __forceinline__ void dft<3>(double2 *a)
{
    const double q = 0.86602540378443864676;     // sqrt(3)/2
    double2 s = cadd(a[1], a[2]), d = csub(a[1], a[2]);
    double2 m = make_double2(a[0].x - 0.5 * s.x, a[0].y - 0.5 * s.y);
    double2 qd = make_double2(q * d.y, -q * d.x);          // -i q d
    a[0] = cadd(a[0], s);
    a[1] = cadd(m, qd);
    a[2] = csub(m, qd);
}

template <>
__device__ __forceinline__ void dft<4>(double2 *a)
{
    double2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
    double2 t2 = cadd(a[1], a[3]), t3 = mul_mi(csub(a[1], a[3]));
    a[0] = cadd(t0, t2);
    a[2] = csub(t0, t2);
    a[1] = cadd(t1, t3);
    a[3] = csub(t1, t3);
}

template <>
__device__ __forceinline__ void dft<5>(double2 *a)
{
    const double c1 = 0.30901699437494742410;    // cos(2pi/5)
    const double c2 = -0.80901699437494742410;   // cos(4pi/5)
    const double s1 = 0.95105651629515357212;    // sin(2pi/5)
    const double s2 = 0.58778525229247312917;    // sin(4pi/5)
    double2 p1 = cadd(a[1], a[4]), m1 = csub(a[1], a[4]);
    double2 p2 = cadd(a[2], a[3]), m2 = csub(a[2], a[3]);
    double2 r1 = make_double2(a[0].x + c1 * p1.x + c2 * p2.x, a[0].y + c1 * p1.y + c2 * p2.y);
    double2 r2 = make_double2(a[0].x + c2 * p1.x + c1 * p2.x, a[0].y + c2 * p1.y + c1 * p2.y);
    // -i (s1 m1 + s2 m2), -i (s2 m1 - s1 m2)
    double2 i1 = make_double2(s1 * m1.y + s2 * m2.y, -(s1 * m1.x + s2 * m2.x));
    double2 i2 = make_double2(s2 * m1.y - s1 * m2.y, -(s2 * m1.x - s1 * m2.x));
    a[0] = make_double2(a[0].x + p1.x + p2.x, a[0].y + p1.y + p2.y);
    a[1] = cadd(r1, i1);
    a[4] = csub(r1, i1);
    a[2] = cadd(r2, i2);
    a[3] = csub(r2, i2);
}

// One in-place DIF pass over S sequences of length P held in shared memory.
template <int R>
__device__ __forceinline__ void dif_pass(double2 *z, int P, int n_cur, int S, const double2 *__restrict__ W)
{
    const int m = n_cur / R;
    const int per_seq = P / R;
    const int tws = P / n_cur;
    for (int idx = threadIdx.x; idx < S * per_seq; idx += blockDim.x) {
        const int s = idx / per_seq;
        const int b = idx - s * per_seq;
        const int blk = b / m;
        const int j = b - blk * m;
        double2 *p = z + (long)s * P + blk * n_cur + j;
        double2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = p[r * m];
        dft<R>(a);
        if (j != 0) {
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(a[r], __ldg(W + j * r * tws));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * m] = a[r];
    }
}

__device__ __forceinline__ void fft_inplace(double2 *z, int P, int S, const PassList &pl,
                                            const double2 *__restrict__ W)
{
    int n_cur = P;
    for (int i = 0; i < pl.npass; ++i) {
        const int R = pl.radix[i];
        if (R == 4) dif_pass<4>(z, P, n_cur, S, W);
        else if (R == 2) dif_pass<2>(z, P, n_cur, S, W);
        else if (R == 3) dif_pass<3>(z, P, n_cur, S, W);
        else dif_pass<5>(z, P, n_cur, S, W);
        n_cur /= R;
        __syncthreads();
    }
}

// input scaling of the three modes (see dct.cu for the conventions)
__device__ __forceinline__ double in_scale(int mode, int n, int P, double v)
{
    if (mode != PDE_DCT_BWD) return v;
    if (n == 0 || n == P) return v;                 // ends doubled, times 0.5
    return (n & 1) ? -0.5 * v : 0.5 * v;
}

// AXIS = 1: sequence q is row q (contiguous); AXIS = 0: sequence q is column q (stride ld).
template <int AXIS>
__global__ void __launch_bounds__(1024)
k_dct_fft(PassList pl, const double2 *__restrict__ W, const double2 *__restrict__ CS,
          const int *__restrict__ pos, int P, int mode, DctPtrs ptrs, long ldx, int n_in,
          long ldy, int n_out, int batch, int S)
{
    extern __shared__ __align__(16) double2 zsm[];
    const double *__restrict__ x = ptrs.x[blockIdx.y];
    double *__restrict__ y = ptrs.y[blockIdx.y];
    const int q0 = blockIdx.x * S;
    const int ns = min(S, batch - q0);
    // ---- load: z_m = e_{2m} + i e_{2m+1}, e = even extension of the (scaled, zero padded) input
    if (AXIS == 1) {
        for (int idx = threadIdx.x; idx < ns * P; idx += blockDim.x) {
            const int s = idx / P, m = idx - s * P;
            const double *row = x + (long)(q0 + s) * ldx;
            int j0 = 2 * m, j1 = 2 * m + 1;
            const int n0 = j0 <= P ? j0 : 2 * P - j0;
            const int n1 = j1 <= P ? j1 : 2 * P - j1;
            const double v0 = n0 < n_in ? in_scale(mode, n0, P, row[n0]) : 0.0;
            const double v1 = n1 < n_in ? in_scale(mode, n1, P, row[n1]) : 0.0;
            zsm[(long)s * P + m] = make_double2(v0, v1);
        }
    } else {
        for (int idx = threadIdx.x; idx < ns * P; idx += blockDim.x) {
            const int m = idx / ns, s = idx - m * ns;      // s fastest: S-wide global segments
            const double *col = x + q0 + s;
            int j0 = 2 * m, j1 = 2 * m + 1;
            const int n0 = j0 <= P ? j0 : 2 * P - j0;
            const int n1 = j1 <= P ? j1 : 2 * P - j1;
            const double v0 = n0 < n_in ? in_scale(mode, n0, P, col[(long)n0 * ldx]) : 0.0;
            const double v1 = n1 < n_in ? in_scale(mode, n1, P, col[(long)n1 * ldx]) : 0.0;
            zsm[(long)s * P + m] = make_double2(v0, v1);
        }
    }
    __syncthreads();
    fft_inplace(zsm, P, ns, pl, W);
    // ---- split + store: outputs k and P-k from Z_k, Z_{P-k}
    const int half = P / 2 + 1;
    const double fscale = 1.0 / (2.0 * (double)P);
    for (int idx = threadIdx.x; idx < ns * half; idx += blockDim.x) {
        int s, k;
        if (AXIS == 1) {
            s = idx / half;
            k = idx - s * half;
        } else {
            k = idx / ns;
            s = idx - k * ns;
        }
        const int k2 = P - k;
        const double2 a = zsm[(long)s * P + __ldg(pos + k)];
        const double2 b = zsm[(long)s * P + __ldg(pos + (k == 0 ? 0 : k2))];
        const double2 cs = __ldg(CS + k);
        const double sr = a.x + b.x, dr = a.x - b.x, si = a.y + b.y;
        double yk = 0.5 * (sr + cs.x * si - cs.y * dr);
        double yk2 = 0.5 * (sr - cs.x * si + cs.y * dr);
        if (mode == PDE_DCT_FWD) {
            // c_k = m_k (-1)^k y_k / (2P), m = [1,2,...,2,1]
            yk *= (k == 0 ? fscale : ((k & 1) ? -2.0 * fscale : 2.0 * fscale));
            yk2 *= (k2 == P ? fscale : ((k2 & 1) ? -2.0 * fscale : 2.0 * fscale));
        }
        if (AXIS == 1) {
            double *row = y + (long)(q0 + s) * ldy;
            if (k < n_out) row[k] = yk;
            if (k2 != k && k2 < n_out) row[k2] = yk2;
        } else {
            double *col = y + q0 + s;
            if (k < n_out) col[(long)k * ldy] = yk;
            if (k2 != k && k2 < n_out) col[(long)k2 * ldy] = yk2;
        }
    }
}

}  // namespace pde
#include "dct_fft_t.cuh"
#include "dct_bluestein.cuh"
namespace pde {

// ---------------------------------------------------------------------------------------
static bool factor(int P, std::vector<int> &radices)
{
    radices.clear();
    int n = P;
    // Power-of-two passes first (large strides: lanes read consecutive addresses), odd radices
    // last: the final small-stride passes then step through shared memory with strides 3 / 5
    // (co-prime to the 8 x 16-byte bank groups) instead of 4 / 16, which removes the 4-way bank
    // conflicts ncu showed for a radix-4 tail (profiles/r01_dct_fft_v1.txt).
    int n2 = 0;
    while (n % 2 == 0) { ++n2; n /= 2; }
    if (n2 & 1) radices.push_back(2);
    for (int i = 0; i < n2 / 2; ++i) radices.push_back(4);
    while (n % 3 == 0) { radices.push_back(3); n /= 3; }
    while (n % 5 == 0) { radices.push_back(5); n /= 5; }
    return n == 1;
}

static int bluestein_length(int P)
{
    int M = 256;
    while (M < 2 * P - 1) M <<= 1;
    return M <= 8192 ? M : 0;
}

int fft_dct_supported(int L)
{
    const int P = L - 1;
    if (P < 2) return 0;
    std::vector<int> r;
    if (!(P & 1) && factor(P, r) && (size_t)P * 16 <= 200 * 1024) return 2;   // direct FFT of length P
    if (P >= 64 && bluestein_length(P)) return 3;                              // chirp-z on a 2^k FFT
    return 0;
}

// digit-reversed position of output k after DIF passes with the given radices
static int digit_rev(int k, int n, const std::vector<int> &rad)
{
    int q = 0;
    for (int r : rad) {
        n /= r;
        q += (k % r) * n;
        k /= r;
    }
    return q;
}

static int bluestein_create(FftDctPlan **out, int L)
{
    const int P = L - 1;
    const int M = bluestein_length(P);
    typedef long double ld;
    const ld pi = 3.141592653589793238462643383279502884L;
    FftDctPlan *p = new FftDctPlan();
    p->L = L;
    p->P = P;
    p->M = M;
    std::vector<int> rad;
    switch (M) {      // must match dispatch_bluestein
    case 256: rad = {16, 16}; break;
    case 512: rad = {16, 16, 2}; break;
    case 1024: rad = {16, 16, 4}; break;
    case 2048: rad = {16, 16, 8}; break;
    case 4096: rad = {16, 16, 16}; break;
    default: rad = {16, 16, 8, 4}; break;
    }
    std::vector<double2> W(M), chirp(P), CS((P + 1) / 2 + 1), Bh(M);
    for (int j = 0; j < M; ++j) {
        ld ang = 2.0L * pi * (ld)j / (ld)M;
        W[j] = make_double2((double)cosl(ang), (double)(-sinl(ang)));
    }
    std::vector<ld> cr(P), ci(P);
    for (long long m = 0; m < P; ++m) {
        const long long q = (m * m) % (2LL * P);          // exact phase index: c_m = exp(i pi q / P)
        ld ang = pi * (ld)q / (ld)P;
        cr[m] = cosl(ang);
        ci[m] = sinl(ang);
        chirp[m] = make_double2((double)cr[m], (double)ci[m]);
    }
    for (int k = 0; k <= (P + 1) / 2; ++k) {
        ld ang = pi * (ld)k / (ld)P;
        CS[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    // Bhat = FFT_M(wrapped chirp) / M in long double (O(M^2) would be too slow: radix-2 FFT on the host)
    std::vector<ld> br(M, 0.0L), bi(M, 0.0L);
    for (int j = 0; j < P; ++j) {
        br[j] = cr[j];
        bi[j] = ci[j];
        if (j) {
            br[M - j] = cr[j];
            bi[M - j] = ci[j];
        }
    }
    {   // iterative radix-2 DIT FFT, forward sign
        int lg = 0;
        while ((1 << lg) < M) ++lg;
        for (int i = 0; i < M; ++i) {
            int j = 0;
            for (int b = 0; b < lg; ++b)
                if (i & (1 << b)) j |= 1 << (lg - 1 - b);
            if (j > i) {
                std::swap(br[i], br[j]);
                std::swap(bi[i], bi[j]);
            }
        }
        for (int len = 2; len <= M; len <<= 1) {
            for (int i = 0; i < M; i += len) {
                for (int k = 0; k < len / 2; ++k) {
                    ld ang = -2.0L * pi * (ld)k / (ld)len;
                    ld wr = cosl(ang), wi = sinl(ang);
                    ld xr = br[i + k + len / 2] * wr - bi[i + k + len / 2] * wi;
                    ld xi = br[i + k + len / 2] * wi + bi[i + k + len / 2] * wr;
                    br[i + k + len / 2] = br[i + k] - xr;
                    bi[i + k + len / 2] = bi[i + k] - xi;
                    br[i + k] += xr;
                    bi[i + k] += xi;
                }
            }
        }
    }
    for (int k = 0; k < M; ++k)
        Bh[digit_rev(k, M, rad)] = make_double2((double)(br[k] / (ld)M), (double)(bi[k] / (ld)M));
    PDE_CUDA(cudaMalloc(&p->W, sizeof(double2) * M));
    PDE_CUDA(cudaMalloc(&p->chirp, sizeof(double2) * P));
    PDE_CUDA(cudaMalloc(&p->CS, sizeof(double2) * CS.size()));
    PDE_CUDA(cudaMalloc(&p->Bhat, sizeof(double2) * M));
    PDE_CUDA(cudaMemcpy(p->W, W.data(), sizeof(double2) * M, cudaMemcpyHostToDevice));
    PDE_CUDA(cudaMemcpy(p->chirp, chirp.data(), sizeof(double2) * P, cudaMemcpyHostToDevice));
    PDE_CUDA(cudaMemcpy(p->CS, CS.data(), sizeof(double2) * CS.size(), cudaMemcpyHostToDevice));
    PDE_CUDA(cudaMemcpy(p->Bhat, Bh.data(), sizeof(double2) * M, cudaMemcpyHostToDevice));
    *out = p;
    return PDE_OK;
}

int fft_dct_create(FftDctPlan **out, int L)
{
    const int P = L - 1;
    std::vector<int> r;
    if (fft_dct_supported(L) == 3) return bluestein_create(out, L);
    if ((P & 1) || !factor(P, r) || r.size() > 24) {
        set_error("fft_dct_create: L-1 = %d is not an even 2^a 3^b 5^c", P);
        return PDE_ERR_UNSUPPORTED;
    }
    FftDctPlan *p = new FftDctPlan();
    p->L = L;
    p->P = P;
    p->npass = (int)r.size();
    for (int i = 0; i < p->npass; ++i) p->radix[i] = r[i];
    const long double pi = 3.141592653589793238462643383279502884L;
    std::vector<double2> W(P), CS(P / 2 + 1);
    for (int j = 0; j < P; ++j) {
        // exact octant reduction keeps the table accurate to long-double rounding
        long double ang = 2.0L * pi * (long double)j / (long double)P;
        W[j] = make_double2((double)cosl(ang), (double)(-sinl(ang)));
    }
    for (int k = 0; k <= P / 2; ++k) {
        long double ang = pi * (long double)k / (long double)P;
        CS[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    std::vector<int> pos(P);
    for (int k = 0; k < P; ++k) {
        int n = P, kk = k, q = 0;
        for (int i = 0; i < p->npass; ++i) {
            n /= r[i];
            q += (kk % r[i]) * n;
            kk /= r[i];
        }
        pos[k] = q;
    }
    PDE_CUDA(cudaMalloc(&p->W, sizeof(double2) * P));
    PDE_CUDA(cudaMalloc(&p->CS, sizeof(double2) * (P / 2 + 1)));
    PDE_CUDA(cudaMalloc(&p->pos, sizeof(int) * P));
    PDE_CUDA(cudaMemcpy(p->W, W.data(), sizeof(double2) * P, cudaMemcpyHostToDevice));
    PDE_CUDA(cudaMemcpy(p->CS, CS.data(), sizeof(double2) * (P / 2 + 1), cudaMemcpyHostToDevice));
    PDE_CUDA(cudaMemcpy(p->pos, pos.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
    static PerDeviceFlag attr;
    if (!attr.get()) {
        PDE_CUDA(cudaFuncSetAttribute(k_dct_fft<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        PDE_CUDA(cudaFuncSetAttribute(k_dct_fft<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        attr.get() = true;
    }
    *out = p;
    return PDE_OK;
}

void fft_dct_destroy(FftDctPlan *p)
{
    if (!p) return;
    cudaFree(p->W);
    cudaFree(p->Wp);
    cudaFree(p->CS);
    cudaFree(p->pos);
    cudaFree(p->chirp);
    cudaFree(p->Bhat);
    delete p;
}

int fft_dct_exec(FftDctPlan *p, int mode, int njobs, const double *const *xs, long ldx, int n_in,
                 double *const *ys, long ldy, int n_out, int batch, int axis, cudaStream_t st)
{
    DctPtrs ptrs{};
    for (int j = 0; j < njobs; ++j) {
        ptrs.x[j] = xs[j];
        ptrs.y[j] = ys[j];
    }
    const int P = p->P;
    if (p->M) {
        BluesteinTables tb{p->W, p->chirp, p->Bhat, p->CS, p->Wp};
        const int rc = axis == 1 ? dispatch_bluestein<1>(p->M, tb, P, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st)
                                 : dispatch_bluestein<0>(p->M, tb, P, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
        p->Wp = const_cast<double2 *>(tb.Wp);          // per-pass tables, built by the first launch
        if (rc < 0) {
            set_error("pde_dct1: no Bluestein kernel for M = %d", p->M);
            return PDE_ERR_UNSUPPORTED;
        }
        return rc;
    }
    if (axis == 1) {   // contiguous rows, 16-byte aligned: persistent CTAs with TMA-prefetched rows
        const int rc = dispatch_row_tma(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
        if (rc >= 0) return rc;
    }
    {   // compile-time specialised kernels for the hot lengths
        const int rc = axis == 1 ? dispatch_fft_t<1>(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st)
                                 : dispatch_fft_t<0>(p, mode, njobs, ptrs, ldx, n_in, ldy, n_out, batch, st);
        if (rc >= 0) return rc;
    }
    const size_t per_seq = (size_t)P * 16;
    // sequences per CTA: aim at <= ~100 KB (2 CTAs/SM); axis 0 wants >= 4 columns for full sectors
    int S = (int)((100 * 1024) / per_seq);
    if (S < 1) S = 1;
    if (axis == 0 && S < 4) S = (int)((200 * 1024) / per_seq) >= 4 ? 4 : (int)((200 * 1024) / per_seq);
    if (S > 16) S = 16;
    // keep the grid at >= ~2 waves when the batch allows it
    while (S > (axis == 0 ? 4 : 1) && ceil_div(batch, S) * njobs < 2 * sm_count()) S >>= 1;
    if (S < 1) S = 1;
    PassList pl;
    pl.npass = p->npass;
    for (int i = 0; i < p->npass; ++i) pl.radix[i] = p->radix[i];
    const size_t smem = per_seq * S;
    const dim3 grid(ceil_div(batch, S), njobs);
    // >= 2 radix-4 butterflies per thread and pass; more warps hide the twiddle / smem latency
    int T = 128;
    while (T < 1024 && (long)T * 8 < (long)S * P) T <<= 1;
    if (axis == 1)
        k_dct_fft<1><<<grid, T, smem, st>>>(pl, p->W, p->CS, p->pos, P, mode, ptrs, ldx, n_in, ldy, n_out, batch, S);
    else
        k_dct_fft<0><<<grid, T, smem, st>>>(pl, p->W, p->CS, p->pos, P, mode, ptrs, ldx, n_in, ldy, n_out, batch, S);
    return after_launch("pde_dct1(fft)");
}

}  // namespace pde
