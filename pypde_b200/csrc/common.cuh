// Shared helpers for the pypde_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <atomic>
#include <type_traits>

#include "pypde_b200.h"

namespace pde {

void set_error(const char *fmt, ...);
extern std::atomic<long> g_launches;

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Every kernel launch in the library goes through this check (counts launches,
// turns launch errors into a PDE_ERR_CUDA return).
inline int after_launch(const char *what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PDE_ERR_CUDA;
    }
    return PDE_OK;
}

#define PDE_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            pde::set_error("%s failed: %s", #call, cudaGetErrorString(e_));             \
            return PDE_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define PDE_REQUIRE(cond, msg)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            pde::set_error("%s: requirement failed: %s (%s)", __func__, #cond, msg);    \
            return PDE_ERR_ARG;                                                         \
        }                                                                               \
    } while (0)

int sm_count();

// "done once" flags for per-device state (cudaFuncSetAttribute applies to the current device only)
struct PerDeviceFlag {
    bool done[64] = {};
    bool &get()
    {
        int d = 0;
        cudaGetDevice(&d);
        return done[d & 63];
    }
};
struct PerDeviceSize {
    size_t v[64] = {};
    size_t &get()
    {
        int d = 0;
        cudaGetDevice(&d);
        return v[d & 63];
    }
};

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// compile-time loop: f(std::integral_constant<int, I>) for I in [I0, N)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
#endif

}  // namespace pde
