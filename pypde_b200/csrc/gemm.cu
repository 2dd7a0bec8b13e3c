// Dense fp64 contraction on the fp64 tensor pipe (DMMA m8n8k4) for sm_100a.
//
// C(m x n) = A(m x k) * B,  B stored (k x n) [TB = false] or (n x k) [TB = true].
// Used for the two dense projections of the eigen-diagonalised Poisson solve
// (Hy, Qy: pypde/templates/poisson.py:93-108) and for the dense-matrix DCT-I of
// small transform lengths.
//
// CTA tile 128 x 64 x 16, 8 warps (4 x 2), warp tile 32 x 32 = 4 x 4 DMMA tiles,
// 3-stage cp.async ring in shared memory; smem rows are padded so that the
// 64-bit fragment loads of a half-warp hit 16 distinct bank pairs.
// fp64 is the only tensor-core precision this path may use (parity <= 1e-12);
// on B200 the DMMA pipe issues 64 FMA/clk/SM, the same peak as the DFMA pipe, so
// the kernel is issue-bound on DMMA long before shared memory matters.
#include "common.cuh"

namespace pde {

constexpr int BN = 64, BK = 16, STAGES = 3;   // BM = 32 * MI (MI m-tiles of 8 rows per warp, 4 warps along M)
constexpr int APITCH = BK + 4;           // doubles per smem row of an (rows x BK) tile
constexpr int BPITCH_NN = BN + 4;        // doubles per smem row of the (BK x BN) tile
constexpr int A_TILE_MAX = 128 * APITCH;      // doubles
constexpr int B_TILE = (BN * APITCH > BK * BPITCH_NN) ? BN * APITCH : BK * BPITCH_NN;
constexpr int GEMM_SMEM = STAGES * (A_TILE_MAX + B_TILE) * 8;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// rows x BK tile of a row-major (k contiguous) matrix -> smem [rows][APITCH]
template <int ROWS, bool VEC>
__device__ __forceinline__ void load_tile_kmajor(double *sm, const double *g, long ld, int row0, int nrows,
                                                 int k0, int K)
{
    if (VEC) {
        constexpr int CHUNKS = ROWS * (BK / 2);
#pragma unroll
        for (int c = threadIdx.x; c < CHUNKS; c += 256) {
            const int r = c / (BK / 2), kc = (c % (BK / 2)) * 2;
            const int gr = row0 + r, gk = k0 + kc;
            int bytes = 0;
            if (gr < nrows) bytes = gk + 1 < K ? 16 : (gk < K ? 8 : 0);
            const double *src = bytes ? g + (long)gr * ld + gk : g;
            cp_async16(sm + r * APITCH + kc, src, bytes);
        }
    } else {
        constexpr int ELEMS = ROWS * BK;
#pragma unroll
        for (int c = threadIdx.x; c < ELEMS; c += 256) {
            const int r = c / BK, kc = c % BK;
            const int gr = row0 + r, gk = k0 + kc;
            const int bytes = (gr < nrows && gk < K) ? 8 : 0;
            const double *src = bytes ? g + (long)gr * ld + gk : g;
            cp_async8(sm + r * APITCH + kc, src, bytes);
        }
    }
}

// BK x BN tile of a row-major (n contiguous) matrix -> smem [BK][BPITCH_NN]
template <bool VEC>
__device__ __forceinline__ void load_tile_nmajor(double *sm, const double *g, long ld, int k0, int K, int n0,
                                                 int N)
{
    if (VEC) {
        constexpr int CHUNKS = BK * (BN / 2);
#pragma unroll
        for (int c = threadIdx.x; c < CHUNKS; c += 256) {
            const int kr = c / (BN / 2), nc = (c % (BN / 2)) * 2;
            const int gk = k0 + kr, gn = n0 + nc;
            int bytes = 0;
            if (gk < K) bytes = gn + 1 < N ? 16 : (gn < N ? 8 : 0);
            const double *src = bytes ? g + (long)gk * ld + gn : g;
            cp_async16(sm + kr * BPITCH_NN + nc, src, bytes);
        }
    } else {
        constexpr int ELEMS = BK * BN;
#pragma unroll
        for (int c = threadIdx.x; c < ELEMS; c += 256) {
            const int kr = c / BN, nc = c % BN;
            const int gk = k0 + kr, gn = n0 + nc;
            const int bytes = (gk < K && gn < N) ? 8 : 0;
            const double *src = bytes ? g + (long)gk * ld + gn : g;
            cp_async8(sm + kr * BPITCH_NN + nc, src, bytes);
        }
    }
}

template <bool TB, bool VEC, int MI>
__global__ void __launch_bounds__(256, 2)
k_gemm_f64(const double *__restrict__ A, long lda, const double *__restrict__ B, long ldb,
           double *__restrict__ C, long ldc, int M, int N, int K)
{
    constexpr int BM = 32 * MI;
    constexpr int A_TILE = BM * APITCH;
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * A_TILE;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 3, wn = warp >> 2;
    const int lr = lane >> 2, lc = lane & 3;

    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int KT = (K + BK - 1) / BK;
    auto load = [&](int kt, int stage) {
        load_tile_kmajor<BM, VEC>(As + stage * A_TILE, A, lda, m0, M, kt * BK, K);
        if (TB) load_tile_kmajor<BN, VEC>(Bs + stage * B_TILE, B, ldb, n0, N, kt * BK, K);
        else load_tile_nmajor<VEC>(Bs + stage * B_TILE, B, ldb, kt * BK, K, n0, N);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        if (nk < KT) load(nk, nk % STAGES);
        cp_async_commit();
        const double *as = As + (kt % STAGES) * A_TILE + (wm * 8 * MI + lr) * APITCH + lc;
        const double *bs = TB ? Bs + (kt % STAGES) * B_TILE + (wn * 32 + lr) * APITCH + lc
                              : Bs + (kt % STAGES) * B_TILE + lc * BPITCH_NN + wn * 32 + lr;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MI], b[4];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * APITCH + kk * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = TB ? bs[j * 8 * APITCH + kk * 4] : bs[kk * 4 * BPITCH_NN + j * 8];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm * 8 * MI + i * 8 + lr;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + lc * 2;
            double *cp = C + (long)row * ldc + col;
            if (VEC && col + 1 < N) {
                *reinterpret_cast<double2 *>(cp) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                if (col < N) cp[0] = acc[i][j][0];
                if (col + 1 < N) cp[1] = acc[i][j][1];
            }
        }
    }
}

int gemm_f64(bool transB, const double *A, long lda, const double *B, long ldb, double *C, long ldc, int m,
             int n, int k, cudaStream_t st)
{
    if (m <= 0 || n <= 0) return PDE_OK;
    const bool vec = (lda % 2 == 0) && (ldb % 2 == 0) && (ldc % 2 == 0) &&
                     ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0);
    // CTA tile 128 x 64 when that fills the machine, else 64 x 64 / 32 x 64 (row slabs of the
    // multi-GPU path and small grids have few rows)
    const int sms = sm_count();
    int MI = 4;
    if ((long)ceil_div(n, BN) * ceil_div(m, 128) < 2L * sms) MI = 2;
    if ((long)ceil_div(n, BN) * ceil_div(m, 64) < 2L * sms) MI = 1;
    dim3 grid(ceil_div(n, BN), ceil_div(m, 32 * MI));
    static bool attr = false;
    auto set_attr = [&](auto kern) {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    };
    if (!attr) {
#define PDE_GEMM_ATTR(TBV, VECV)                                  \
        PDE_CUDA(set_attr(k_gemm_f64<TBV, VECV, 4>));               \
        PDE_CUDA(set_attr(k_gemm_f64<TBV, VECV, 2>));               \
        PDE_CUDA(set_attr(k_gemm_f64<TBV, VECV, 1>));
        PDE_GEMM_ATTR(false, false)
        PDE_GEMM_ATTR(false, true)
        PDE_GEMM_ATTR(true, false)
        PDE_GEMM_ATTR(true, true)
#undef PDE_GEMM_ATTR
        attr = true;
    }
#define PDE_GEMM_LAUNCH(TBV, VECV)                                                                              \
    do {                                                                                                        \
        if (MI == 4) k_gemm_f64<TBV, VECV, 4><<<grid, 256, GEMM_SMEM, st>>>(A, lda, B, ldb, C, ldc, m, n, k);   \
        else if (MI == 2) k_gemm_f64<TBV, VECV, 2><<<grid, 256, GEMM_SMEM, st>>>(A, lda, B, ldb, C, ldc, m, n, k); \
        else k_gemm_f64<TBV, VECV, 1><<<grid, 256, GEMM_SMEM, st>>>(A, lda, B, ldb, C, ldc, m, n, k);           \
    } while (0)
    if (transB) {
        if (vec) PDE_GEMM_LAUNCH(true, true);
        else PDE_GEMM_LAUNCH(true, false);
    } else {
        if (vec) PDE_GEMM_LAUNCH(false, true);
        else PDE_GEMM_LAUNCH(false, false);
    }
#undef PDE_GEMM_LAUNCH
    return after_launch("pde_gemm_f64");
}

}  // namespace pde

extern "C" int pde_gemm_f64(int transB, const double *A, long lda, const double *B, long ldb, double *C,
                            long ldc, int m, int n, int k, void *stream)
{
    PDE_REQUIRE(A && B && C, "null pointer");
    PDE_REQUIRE(k >= 1, "k >= 1");
    PDE_REQUIRE(lda >= k && ldc >= n && ldb >= (transB ? k : n), "leading dimensions");
    return pde::gemm_f64(transB != 0, A, lda, B, ldb, C, ldc, m, n, k, pde::as_stream(stream));
}
