// placeholder: shared-memory FFT DCT-I (filled in next)
#include "common.cuh"
namespace pde {
struct FftDctPlan { int L; };
int fft_dct_supported(int) { return 0; }
int fft_dct_create(FftDctPlan **, int) { set_error("fft dct not built"); return PDE_ERR_UNSUPPORTED; }
void fft_dct_destroy(FftDctPlan *) {}
int fft_dct_exec(FftDctPlan *, int, const double *, long, int, double *, long, int, int, int, cudaStream_t)
{ set_error("fft dct not built"); return PDE_ERR_UNSUPPORTED; }
}
