// Dense fp64 contraction on the fp64 tensor pipe (DMMA m8n8k4) for sm_100a.
//
// C(m x n) = A(m x k) * B,  B stored (k x n) [TB = false] or (n x k) [TB = true].
// Used for the two dense projections of the eigen-diagonalised Poisson solve
// (Hy, Qy: pypde/templates/poisson.py:93-108) and for the dense-matrix DCT-I of
// small transform lengths.
//
// CTA tile (8 MI WM) x (8 NI WN) x 16 with 8 warps laid out WM x WN, warp tile = MI x NI DMMA tiles
// (128 x 64 = 4 x 2 warps of 4 x 4 tiles; 128 x 56 = 8 x 1 warps of 2 x 7 tiles; 64 x 64; 32 x 64),
// cp.async ring in shared memory (K depth 32 x 2 stages or 16 x 3, chosen per launch); smem rows are padded so that the
// 64-bit fragment loads of a half-warp hit 16 distinct bank pairs.  The tile shape is chosen per
// problem to fill whole waves of 2 CTAs per SM: the Poisson projections of rbc2048 (2046 x 2046) are
// 512 tiles of 128 x 64 = 1.73 waves on 148 SMs, but 592 tiles of 128 x 56 = exactly 2 waves (-12 %).
// fp64 is the only tensor-core precision this path may use (parity <= 1e-12);
// on B200 the DMMA pipe issues 64 FMA/clk/SM, the same peak as the DFMA pipe, so
// the kernel is issue-bound on DMMA long before shared memory matters.
#include "common.cuh"
#include <cstdlib>

namespace pde {

// K depth BK of a shared-memory tile and depth of the cp.async ring: template parameters, two configurations.
// Two CTAs per SM must fit (<= 113 KB each).  Measured on the 2046 x 2046 x 2048 projections of rbc2048 (ms per
// step, 6 products): BK 16 x 3 stages 3.677, 8 x 4: 3.53, 8 x 5: 3.56, 32 x 2: 3.476 (half the barriers per flop) -- but
// on the row slabs of the multi-GPU path (1024 / 512 rows: ONE wave of tiles, all CTAs start together) the deeper
// ring wins: 16 x 3 1.95 / 1.15 ms, 32 x 2 2.16 / 1.20 ms.  So: 32 x 2 for launches of two or more waves, else 16 x 3.
// (256 x 56 tiles, one 8-warp CTA per SM with 202 registers, fewer fragment loads per DMMA: 3.93 -- 16 warps per SM matter.)
constexpr int BN = 64;                         // BM = 32 * MI (MI m-tiles of 8 rows per warp, 4 warps along M)
constexpr int BPITCH_NN = BN + 4;              // doubles per smem row of the (BK x BN) tile
template <int BK>
struct KTile {
    static constexpr int APITCH = BK + 4;      // doubles per smem row of an (rows x BK) tile
    static constexpr int A_TILE_MAX = 128 * APITCH;
    static constexpr int B_TILE = (BN * APITCH > BK * BPITCH_NN) ? BN * APITCH : BK * BPITCH_NN;
};
template <int BK, int STAGES>
constexpr int gemm_smem() { return STAGES * (KTile<BK>::A_TILE_MAX + KTile<BK>::B_TILE) * 8; }

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
template <int ROWS, bool VEC, int BK>
__device__ __forceinline__ void load_tile_kmajor(double *sm, const double *g, long ld, int row0, int nrows,
                                                 int k0, int K)
{
    constexpr int APITCH = KTile<BK>::APITCH;
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

// BK x BNT tile of a row-major (n contiguous) matrix -> smem [BK][BPITCH_NN]
template <bool VEC, int BNT, int BK>
__device__ __forceinline__ void load_tile_nmajor(double *sm, const double *g, long ld, int k0, int K, int n0,
                                                 int N)
{
    if (VEC) {
        constexpr int CHUNKS = BK * (BNT / 2);
#pragma unroll
        for (int c = threadIdx.x; c < CHUNKS; c += 256) {
            const int kr = c / (BNT / 2), nc = (c % (BNT / 2)) * 2;
            const int gk = k0 + kr, gn = n0 + nc;
            int bytes = 0;
            if (gk < K) bytes = gn + 1 < N ? 16 : (gn < N ? 8 : 0);
            const double *src = bytes ? g + (long)gk * ld + gn : g;
            cp_async16(sm + kr * BPITCH_NN + nc, src, bytes);
        }
    } else {
        constexpr int ELEMS = BK * BNT;
#pragma unroll
        for (int c = threadIdx.x; c < ELEMS; c += 256) {
            const int kr = c / BNT, nc = c % BNT;
            const int gk = k0 + kr, gn = n0 + nc;
            const int bytes = (gk < K && gn < N) ? 8 : 0;
            const double *src = bytes ? g + (long)gk * ld + gn : g;
            cp_async8(sm + kr * BPITCH_NN + nc, src, bytes);
        }
    }
}

// batched form: blockIdx.z picks the operands of one problem out of DEVICE pointer arrays (a null array = the
// operand is shared by all problems): the ensemble runs the dense DCT / projection products of all its members
// in one launch
struct GemmBatch {
    const double *const *A;
    const double *const *B;
    double *const *C;
};

template <bool TB, bool VEC, int MI, int NI, int WN, int BK, int STAGES>
__global__ void __launch_bounds__(256, 2)
k_gemm_f64(const double *__restrict__ A, long lda, const double *__restrict__ B, long ldb,
           double *__restrict__ C, long ldc, int M, int N, int K, GemmBatch bt)
{
    if (bt.A) A = bt.A[blockIdx.z];
    if (bt.B) B = bt.B[blockIdx.z];
    if (bt.C) C = bt.C[blockIdx.z];
    constexpr int WM = 8 / WN;
    constexpr int BM = WM * 8 * MI, BNT = WN * 8 * NI;
    static_assert(BM <= 128 && BNT <= BN, "tile exceeds the shared-memory layout");
    constexpr int APITCH = KTile<BK>::APITCH, B_TILE = KTile<BK>::B_TILE;
    constexpr int A_TILE = BM * APITCH;
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * A_TILE;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BNT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp % WM, wn = warp / WM;
    const int lr = lane >> 2, lc = lane & 3;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int KT = (K + BK - 1) / BK;
    auto load = [&](int kt, int stage) {
        load_tile_kmajor<BM, VEC, BK>(As + stage * A_TILE, A, lda, m0, M, kt * BK, K);
        if (TB) load_tile_kmajor<BNT, VEC, BK>(Bs + stage * B_TILE, B, ldb, n0, N, kt * BK, K);
        else load_tile_nmajor<VEC, BNT, BK>(Bs + stage * B_TILE, B, ldb, kt * BK, K, n0, N);
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
        const double *bs = TB ? Bs + (kt % STAGES) * B_TILE + (wn * 8 * NI + lr) * APITCH + lc
                              : Bs + (kt % STAGES) * B_TILE + lc * BPITCH_NN + wn * 8 * NI + lr;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * APITCH + kk * 4];
#pragma unroll
            for (int j = 0; j < NI; ++j)
                b[j] = TB ? bs[j * 8 * APITCH + kk * 4] : bs[kk * 4 * BPITCH_NN + j * 8];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm * 8 * MI + i * 8 + lr;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = n0 + wn * 8 * NI + j * 8 + lc * 2;
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

// tile shapes: {BM, BN, MI, NI, WN}
struct GemmShape {
    int bm, bn;
};
static const GemmShape GEMM_SHAPES[] = {{128, 64}, {128, 56}, {128, 48}, {64, 64}, {32, 64}};

template <bool TB, bool VEC, int MI, int NI, int WN, int BK, int STAGES>
static int gemm_launch_k(const double *A, long lda, const double *B, long ldb, double *C, long ldc, int m, int n,
                         int k, cudaStream_t st, GemmBatch bt, int nbatch)
{
    auto kern = k_gemm_f64<TB, VEC, MI, NI, WN, BK, STAGES>;
    constexpr int SMEM = gemm_smem<BK, STAGES>();
    static PerDeviceFlag attr;
    if (!attr.get()) {
        PDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr.get() = true;
    }
    constexpr int BM = (8 / WN) * 8 * MI, BNT = WN * 8 * NI;
    dim3 grid(ceil_div(n, BNT), ceil_div(m, BM), nbatch);
    kern<<<grid, 256, SMEM, st>>>(A, lda, B, ldb, C, ldc, m, n, k, bt);
    return after_launch("pde_gemm_f64");
}

template <bool TB, bool VEC, int MI, int NI, int WN>
static int gemm_launch(const double *A, long lda, const double *B, long ldb, double *C, long ldc, int m, int n,
                       int k, cudaStream_t st, GemmBatch bt = GemmBatch{nullptr, nullptr, nullptr}, int nbatch = 1)
{
    // two or more waves of CTAs (2 per SM): 32-deep K tiles, 2-stage ring; a single wave: 16-deep, 3 stages
    constexpr int BM = (8 / WN) * 8 * MI, BNT = WN * 8 * NI;
    const long tiles = (long)ceil_div(n, BNT) * ceil_div(m, BM) * nbatch;
    static const int forced = getenv("PDE_GEMM_BK") ? atoi(getenv("PDE_GEMM_BK")) : 0;
    const bool deep = forced ? forced == 32 : tiles >= 4L * sm_count();
    if (deep) return gemm_launch_k<TB, VEC, MI, NI, WN, 32, 2>(A, lda, B, ldb, C, ldc, m, n, k, st, bt, nbatch);
    return gemm_launch_k<TB, VEC, MI, NI, WN, 16, 3>(A, lda, B, ldb, C, ldc, m, n, k, st, bt, nbatch);
}

template <bool TB, bool VEC>
static int gemm_shape(int shape, const double *A, long lda, const double *B, long ldb, double *C, long ldc, int m,
                      int n, int k, cudaStream_t st, GemmBatch bt = GemmBatch{nullptr, nullptr, nullptr}, int nbatch = 1)
{
    switch (shape) {
    case 0: return gemm_launch<TB, VEC, 4, 4, 2>(A, lda, B, ldb, C, ldc, m, n, k, st, bt, nbatch);   // 128 x 64
    case 1: return gemm_launch<TB, VEC, 2, 7, 1>(A, lda, B, ldb, C, ldc, m, n, k, st, bt, nbatch);   // 128 x 56
    case 2: return gemm_launch<TB, VEC, 2, 6, 1>(A, lda, B, ldb, C, ldc, m, n, k, st, bt, nbatch);   // 128 x 48
    case 3: return gemm_launch<TB, VEC, 2, 4, 2>(A, lda, B, ldb, C, ldc, m, n, k, st, bt, nbatch);   // 64 x 64
    default: return gemm_launch<TB, VEC, 1, 4, 2>(A, lda, B, ldb, C, ldc, m, n, k, st, bt, nbatch);  // 32 x 64
    }
}

int gemm_f64(bool transB, const double *A, long lda, const double *B, long ldb, double *C, long ldc, int m,
             int n, int k, cudaStream_t st)
{
    if (m <= 0 || n <= 0) return PDE_OK;
    const bool vec = (lda % 2 == 0) && (ldb % 2 == 0) && (ldc % 2 == 0) &&
                     ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0);
    // Tile shape: the one that needs the least (waves x tile area) with 2 CTAs per SM; 64- and 32-row
    // tiles only when the 128-row ones cannot fill the machine (row slabs of the multi-GPU path, small grids).
    const long slots = 2L * sm_count();
    int shape = 0;
    static const int forced = getenv("PDE_GEMM_SHAPE") ? atoi(getenv("PDE_GEMM_SHAPE")) : -1;
    if (forced >= 0 && forced <= 4) {
        shape = forced;
    } else if ((long)ceil_div(n, 64) * ceil_div(m, 64) < slots) {
        shape = 4;
    } else if ((long)ceil_div(n, 64) * ceil_div(m, 128) < slots) {
        shape = 3;
    } else {
        double best = 0.0;
        for (int sidx = 0; sidx < 3; ++sidx) {
            const GemmShape &g = GEMM_SHAPES[sidx];
            const long tiles = (long)ceil_div(n, g.bn) * ceil_div(m, g.bm);
            // narrower tiles re-read A more often and issue more fragment loads per DMMA: 3 % handicap per step
            const double cost = (double)ceil_div(tiles, slots) * g.bm * g.bn * (1.0 + 0.03 * sidx);
            if (sidx == 0 || cost < best) {
                best = cost;
                shape = sidx;
            }
        }
    }
    if (transB) return vec ? gemm_shape<true, true>(shape, A, lda, B, ldb, C, ldc, m, n, k, st)
                           : gemm_shape<true, false>(shape, A, lda, B, ldb, C, ldc, m, n, k, st);
    return vec ? gemm_shape<false, true>(shape, A, lda, B, ldb, C, ldc, m, n, k, st)
               : gemm_shape<false, false>(shape, A, lda, B, ldb, C, ldc, m, n, k, st);
}

// nbatch problems of one shape; operands from device pointer arrays (null array: shared operand A / B).
// `vec`: the caller guarantees 16-byte aligned operands with even leading dimensions.
int gemm_f64_batched(bool transB, const double *A, const double *const *Aarr, long lda, const double *B,
                     const double *const *Barr, long ldb, double *const *Carr, long ldc, int m, int n, int k,
                     int nbatch, bool vec, cudaStream_t st)
{
    if (m <= 0 || n <= 0 || nbatch <= 0) return PDE_OK;
    const long slots = 2L * sm_count();
    int shape = 0;
    if ((long)ceil_div(n, 64) * ceil_div(m, 64) * nbatch < slots) shape = 4;
    else if ((long)ceil_div(n, 64) * ceil_div(m, 128) * nbatch < slots || m <= 64) shape = m <= 32 ? 4 : 3;
    GemmBatch bt{Aarr, Barr, Carr};
    if (transB) return vec ? gemm_shape<true, true>(shape, A, lda, B, ldb, nullptr, ldc, m, n, k, st, bt, nbatch)
                           : gemm_shape<true, false>(shape, A, lda, B, ldb, nullptr, ldc, m, n, k, st, bt, nbatch);
    return vec ? gemm_shape<false, true>(shape, A, lda, B, ldb, nullptr, ldc, m, n, k, st, bt, nbatch)
               : gemm_shape<false, false>(shape, A, lda, B, ldb, nullptr, ldc, m, n, k, st, bt, nbatch);
}

}  // namespace pde

extern "C" int pde_gemm_f64_batched(int transB, const double *A, const double *const *dev_A, long lda,
                                    const double *B, const double *const *dev_B, long ldb, double *const *dev_C,
                                    long ldc, int m, int n, int k, int nbatch, int aligned, void *stream)
{
    PDE_REQUIRE((A != nullptr) != (dev_A != nullptr), "exactly one of A / dev_A");
    PDE_REQUIRE((B != nullptr) != (dev_B != nullptr), "exactly one of B / dev_B");
    PDE_REQUIRE(dev_C != nullptr && k >= 1 && nbatch >= 1 && nbatch <= 65535, "arguments");
    PDE_REQUIRE(lda >= k && ldc >= n && ldb >= (transB ? k : n), "leading dimensions");
    const bool vec = aligned && (lda % 2 == 0) && (ldb % 2 == 0) && (ldc % 2 == 0) &&
                     (!A || (uintptr_t)A % 16 == 0) && (!B || (uintptr_t)B % 16 == 0);
    return pde::gemm_f64_batched(transB != 0, A, dev_A, lda, B, dev_B, ldb, dev_C, ldc, m, n, k, nbatch, vec,
                                 pde::as_stream(stream));
}

extern "C" int pde_gemm_f64(int transB, const double *A, long lda, const double *B, long ldb, double *C,
                            long ldc, int m, int n, int k, void *stream)
{
    PDE_REQUIRE(A && B && C, "null pointer");
    PDE_REQUIRE(k >= 1, "k >= 1");
    PDE_REQUIRE(lda >= k && ldc >= n && ldb >= (transB ? k : n), "leading dimensions");
    return pde::gemm_f64(transB != 0, A, lda, B, ldb, C, ldc, m, n, k, pde::as_stream(stream));
}
