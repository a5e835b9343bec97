// Shared helpers for libpdb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/pdb200.h"

namespace pdb {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

// Call right after a <<<>>> launch.
inline int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
    return PDB_OK;
}

#define PDB_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) return pdb::fail(PDB_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define PDB_TRY(expr)               \
    do {                            \
        int _rc = (expr);           \
        if (_rc != PDB_OK) return _rc; \
    } while (0)

constexpr int kNumSMs = 148;   // B200
constexpr int kMaxLevels = 8;

struct LevelTable {
    int h[kMaxLevels];
    int w[kMaxLevels];
    int start[kMaxLevels];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 16-byte vector reduction into global memory (sm_90+): one L2 atomic transaction per 4 floats.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace pdb
