// DCT-I plans (sm_100a): replaces scipy.fftpack.dctn(type=1) and the scale /
// sign / mass passes of Chebyshev.forward_fft / backward_fft
// (pypde/bases/chebyshev.py:67-93, :143-149).
//
// Algorithms
//   DENSE (1): y = C x with the L x L cosine matrix (all scale factors folded in,
//              entries computed on the host in long double with exact integer
//              argument reduction) on the fp64 tensor pipe (gemm.cu).  Used for
//              short transforms, and as the reference algorithm for the others.
//   FFT   (2): shared-memory complex FFT of length P = L-1 (P even, factors
//              2/3/5) of the even extension, see dct_fft.cu.
//   BLUESTEIN (3): chirp-z on a power-of-two FFT for every other length.
#include "common.cuh"
#include <cmath>
#include <vector>

namespace pde {

int gemm_f64_batched(bool transB, const double *A, const double *const *Aarr, long lda, const double *B,
                     const double *const *Barr, long ldb, double *const *Carr, long ldc, int m, int n, int k,
                     int nbatch, bool vec, cudaStream_t st);
int gemm_f64(bool transB, const double *A, long lda, const double *B, long ldb, double *C, long ldc, int m,
             int n, int k, cudaStream_t st);

// dct_fft.cu
struct FftDctPlan;
int fft_dct_supported(int L);
int fft_dct_create(FftDctPlan **p, int L);
void fft_dct_destroy(FftDctPlan *p);
int fft_dct_exec(FftDctPlan *p, int mode, int njobs, const double *const *xs, long ldx, int n_in,
                 double *const *ys, long ldy, int n_out, int batch, int axis, cudaStream_t st);

// cos(pi * r / P) for integers 0 <= r, P > 0, accurate to long-double rounding
static long double cos_pi_frac(long long r, long long P)
{
    r %= 2 * P;
    if (r > P) r = 2 * P - r;              // cos(2 pi - t) = cos t
    long double sgn = 1.0L;
    if (2 * r > P) {                        // cos(pi - t) = -cos t
        r = P - r;
        sgn = -1.0L;
    }
    const long double pi = 3.141592653589793238462643383279502884L;
    if (4 * r > P)                          // t > pi/4: cos t = sin(pi/2 - t)
        return sgn * sinl(pi * (long double)(P - 2 * r) / (long double)(2 * P));
    return sgn * cosl(pi * (long double)r / (long double)P);
}

}  // namespace pde

using namespace pde;

struct pde_dct_plan_s {
    int L = 0, algo = 0;
    long ldm = 0;                      // leading dimension of the dense matrices (even)
    double *mat[3] = {nullptr, nullptr, nullptr};
    FftDctPlan *fft = nullptr;
};

static int build_dense(pde_dct_plan_s *p, int mode)
{
    if (p->mat[mode]) return PDE_OK;
    const int L = p->L;
    const long long P = L - 1;
    const long ld = p->ldm;
    std::vector<double> h((size_t)L * ld, 0.0);
    for (long long k = 0; k < L; ++k) {
        for (long long n = 0; n < L; ++n) {
            const long double c = cos_pi_frac(k * n, P);
            long double v;
            if (mode == PDE_DCT_RAW) {
                v = ((n == 0 || n == P) ? 1.0L : 2.0L) * c;
            } else if (mode == PDE_DCT_FWD) {
                // c_k = m_k (-1)^k (0.5 * DCT1(f)_k) / (N-1), m = [1,2,...,2,1]
                const long double wn = (n == 0 || n == P) ? 1.0L : 2.0L;
                const long double mk = (k == 0 || k == P) ? 1.0L : 2.0L;
                v = ((k & 1) ? -1.0L : 1.0L) * mk * wn * c / (2.0L * (long double)P);
            } else {
                // f_j = 0.5 * DCT1(g)_j, g_n = (-1)^n c_n, ends doubled  =>  sum_n (-1)^n c_n cos(pi j n / P)
                v = ((n & 1) ? -1.0L : 1.0L) * c;
            }
            h[(size_t)k * ld + n] = (double)v;
        }
    }
    double *d = nullptr;
    PDE_CUDA(cudaMalloc(&d, sizeof(double) * h.size()));
    PDE_CUDA(cudaMemcpy(d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    p->mat[mode] = d;
    return PDE_OK;
}

extern "C" {

int pde_dct_plan_create(pde_dct_plan_t *plan, int L, int algo)
{
    PDE_REQUIRE(plan, "null plan pointer");
    PDE_REQUIRE(L >= 2, "DCT-I needs L >= 2");
    PDE_REQUIRE(algo >= 0 && algo <= 3, "algo in 0..3");
    if (algo == 0) {
        // FFT whenever L-1 is even and smooth (specialised kernels exist from P = 96 on); the dense
        // DMMA matrix for very short or awkward lengths
        // (3: Bluestein pays off from L ~ 512 on; below, the dense matrix is faster)
        const int sup = fft_dct_supported(L);
        // measured (profiles/r01_dct_sweep.json): dense 1.8 TB/s at L = 128, 0.97 TB/s at 256; FFT ~0.85 TB/s
        algo = (sup == 2 && L >= 320) ? 2 : ((sup == 3 && L >= 384) ? 3 : 1);
    }
    pde_dct_plan_s *p = new pde_dct_plan_s();
    p->L = L;
    p->algo = algo;
    p->ldm = (L + 1) & ~1L;
    if (algo == 2 || algo == 3) {
        if (fft_dct_supported(L) != algo && !(algo == 3 && fft_dct_supported(L) == 2)) {
            delete p;
            set_error("pde_dct_plan_create: algo %d does not support L = %d", algo, L);
            return PDE_ERR_UNSUPPORTED;
        }
        int rc = fft_dct_create(&p->fft, L);
        if (rc != PDE_OK) {
            delete p;
            return rc;
        }
    }
    *plan = p;
    return PDE_OK;
}

int pde_dct_plan_destroy(pde_dct_plan_t p)
{
    if (!p) return PDE_OK;
    for (int i = 0; i < 3; ++i) cudaFree(p->mat[i]);
    if (p->fft) fft_dct_destroy(p->fft);
    delete p;
    return PDE_OK;
}

int pde_dct_plan_algo(pde_dct_plan_t p) { return p ? p->algo : 0; }

int pde_dct1_multi(pde_dct_plan_t p, int mode, int njobs, const double *const *x, long ldx, int n_in,
                   double *const *y, long ldy, int n_out, int batch, int axis, void *stream)
{
    PDE_REQUIRE(p && x && y, "null pointer");
    PDE_REQUIRE(njobs >= 1 && njobs <= PDE_MAX_JOBS, "1..8 jobs");
    PDE_REQUIRE(mode >= 0 && mode <= 2, "mode");
    PDE_REQUIRE(axis == 0 || axis == 1, "axis");
    PDE_REQUIRE(n_in >= 1 && n_in <= p->L && n_out >= 1 && n_out <= p->L, "n_in / n_out in 1..L");
    for (int j = 0; j < njobs; ++j) PDE_REQUIRE(x[j] && y[j] && x[j] != y[j], "null / in-place transform");
    if (batch <= 0) return PDE_OK;
    cudaStream_t st = as_stream(stream);
    if (p->algo != 1)
        return fft_dct_exec(p->fft, mode, njobs, x, ldx, n_in, y, ldy, n_out, batch, axis, st);
    int rc = build_dense(p, mode);
    for (int j = 0; j < njobs && rc == PDE_OK; ++j) {
        if (axis == 0)   // Y(n_out x batch) = Mat(n_out x n_in) X(n_in x batch)
            rc = gemm_f64(false, p->mat[mode], p->ldm, x[j], ldx, y[j], ldy, n_out, batch, n_in, st);
        else             // Y(batch x n_out) = X(batch x n_in) Mat(n_out x n_in)^T
            rc = gemm_f64(true, x[j], ldx, p->mat[mode], p->ldm, y[j], ldy, batch, n_out, n_in, st);
    }
    return rc;
}

int pde_dct1_batched(pde_dct_plan_t p, int mode, int nbatch, const double *const *dev_x, long ldx, int n_in,
                     double *const *dev_y, long ldy, int n_out, int batch, int axis, int aligned, void *stream)
{
    PDE_REQUIRE(p && dev_x && dev_y, "null pointer");
    PDE_REQUIRE(mode >= 0 && mode <= 2 && (axis == 0 || axis == 1) && nbatch >= 1, "arguments");
    PDE_REQUIRE(n_in >= 1 && n_in <= p->L && n_out >= 1 && n_out <= p->L, "n_in / n_out in 1..L");
    if (p->algo != 1) {
        set_error("pde_dct1_batched: plan of length %d uses the FFT path; the batched form is the dense-matrix path", p->L);
        return PDE_ERR_UNSUPPORTED;
    }
    if (batch <= 0) return PDE_OK;
    int rc = build_dense(p, mode);
    if (rc != PDE_OK) return rc;
    const bool vec = aligned && ldx % 2 == 0 && ldy % 2 == 0 && p->ldm % 2 == 0;
    if (axis == 0)   // Y(n_out x batch) = Mat(n_out x n_in) X(n_in x batch)
        return gemm_f64_batched(false, p->mat[mode], nullptr, p->ldm, nullptr, dev_x, ldx, dev_y, ldy, n_out, batch, n_in,
                                nbatch, vec, as_stream(stream));
    // Y(batch x n_out) = X(batch x n_in) Mat(n_out x n_in)^T
    return gemm_f64_batched(true, nullptr, dev_x, ldx, p->mat[mode], nullptr, p->ldm, dev_y, ldy, batch, n_out, n_in, nbatch,
                            vec, as_stream(stream));
}

int pde_dct1(pde_dct_plan_t p, int mode, const double *x, long ldx, int n_in, double *y, long ldy,
             int n_out, int batch, int axis, void *stream)
{
    return pde_dct1_multi(p, mode, 1, &x, ldx, n_in, &y, ldy, n_out, batch, axis, stream);
}

}  // extern "C"
