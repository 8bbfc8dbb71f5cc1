// common.cuh -- error plumbing and small device helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/shgo_b200.h"

namespace shgo_b200 {

constexpr int kMaxDim = 32;        // fused builtin evaluation (tables live in shared memory)
constexpr int kSobolBits = 30;     // scipy's generator width
constexpr int kEvalThreads = 256;  // one Sobol "low" block = 2^8 consecutive indices

char *err_buf();
int fail(int code, const char *fmt, ...);

#define SHGO_CUDA_TRY(expr)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ::shgo_b200::fail(SHGO_B200_ERR_CUDA, "%s failed: %s", #expr,             \
                                     cudaGetErrorString(e__));                               \
    } while (0)

#define SHGO_REQUIRE(cond, ...)                                                              \
    do {                                                                                     \
        if (!(cond)) return ::shgo_b200::fail(SHGO_B200_ERR_ARG, __VA_ARGS__);               \
    } while (0)

// sm count of the current device (cached per device); persistent grids are sized from it
int sm_count();

// ordered compaction of the flags equal to `match` (graph.cu); tmp: shgo_b200_compact_workspace(n)
int compact_flags(const uint8_t *flags, uint64_t n, int match, int64_t base, int64_t *idx_out,
                  uint64_t cap, uint64_t *count, void *tmp, cudaStream_t st);

// stable LSD radix sort of (key, id) pairs by 32-bit key, 8 bits per pass (reorder.cu).  Input and
// result live in (k0, i0); (k1, i1) and hist (radix_hist_bytes(n) bytes) are scratch.
size_t radix_hist_bytes(uint64_t n);
int radix_sort_pairs(uint32_t *k0, int32_t *i0, uint32_t *k1, int32_t *i1, uint64_t *hist, uint64_t n,
                     cudaStream_t st);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ double ld_nc_f64(const double *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

}  // namespace shgo_b200
