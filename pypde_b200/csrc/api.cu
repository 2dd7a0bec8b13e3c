// Library-level entry points of the C ABI: error string, version, device info,
// launch counter.
#include "common.cuh"

namespace pde {

static thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace pde

extern "C" {

const char *pde_last_error(void) { return pde::g_err; }

int pde_version(void) { return 100; }

int pde_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    PDE_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) {
        PDE_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (cc_major) {
        PDE_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        PDE_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return PDE_OK;
}

long pde_launch_count(void) { return pde::g_launches.load(); }
void pde_launch_count_reset(void) { pde::g_launches.store(0); }

}  // extern "C"
