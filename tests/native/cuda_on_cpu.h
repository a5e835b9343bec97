// TEST INFRASTRUCTURE ONLY.  A minimal "CUDA on the CPU" shim: enough of the CUDA C++ surface to compile the simple
// byte / bit kernels of partdistillation_b200/csrc/postprocess_kernels.cuh with g++ and execute them with CUDA's
// semantics — one OS thread per CUDA thread of a block (a pool per launch), blocks run one after the other, __syncthreads and the warp
// collectives (__ballot_sync, __reduce_add_sync, __syncthreads_or) implemented with std::barrier so that divergent
// exits behave as on the device (a thread that returns drops out of its warp's and its block's barriers).
// Used by tests/test_postprocess_host_cpu.py in the build container, which has no GPU.  Never part of the product.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
// bf16 pairs as the kernels store them (round to nearest even; NaN payloads are not needed by the tests)
struct alignas(4) __nv_bfloat162 { unsigned short x, y; };
inline unsigned short host_f2bf16_rn(float f) {
    unsigned u;
    __builtin_memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40);
    return (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{host_f2bf16_rn(a), host_f2bf16_rn(b)}; }

#define __global__
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static          // blocks run one at a time, so one static instance = one block's shared memory

namespace cpu_cuda {

struct Warp {
    std::unique_ptr<std::barrier<>> bar;
    unsigned pred[32];
    int val[32];
    unsigned long long xchg[32];        // __shfl_*_sync payloads (up to 8 bytes per lane)
    unsigned sub_arrived = 0;           // rendezvous of a SUBSET of the lanes (collectives with a partial member mask)
    unsigned sub_gen = 0;
};

struct Block {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<Warp> warps;
    std::atomic<int> or_flag[2];
};

inline void* g_dyn_smem = nullptr;           // dynamic shared memory of the running launch (one block at a time)
inline thread_local Block* t_block = nullptr;
inline thread_local Warp* t_warp = nullptr;
inline thread_local int t_lane = 0;
inline thread_local int t_or_phase = 0;

}  // namespace cpu_cuda

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void __syncthreads() { cpu_cuda::t_block->bar->arrive_and_wait(); }

inline int __syncthreads_or(int pred) {
    auto* b = cpu_cuda::t_block;
    const int ph = cpu_cuda::t_or_phase;
    cpu_cuda::t_or_phase ^= 1;
    if (pred) b->or_flag[ph].store(1);
    b->bar->arrive_and_wait();
    const int r = b->or_flag[ph].load();
    b->bar->arrive_and_wait();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) b->or_flag[ph].store(0);   // next use of this phase is two barriers away
    return r;
}

inline unsigned __ballot_sync(unsigned, int pred) {
    auto* w = cpu_cuda::t_warp;
    w->pred[cpu_cuda::t_lane] = pred ? 1u : 0u;
    w->bar->arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= w->pred[i] << i;
    w->bar->arrive_and_wait();
    return r;
}

inline int __reduce_add_sync(unsigned, int v) {
    auto* w = cpu_cuda::t_warp;
    w->val[cpu_cuda::t_lane] = v;
    w->bar->arrive_and_wait();
    int r = 0;
    for (int i = 0; i < 32; ++i) r += w->val[i];
    w->bar->arrive_and_wait();
    return r;
}

namespace cpu_cuda {
// Convergence point of a warp collective.  Full mask: every non-exited lane takes part (std::barrier with drop-on-exit).
// Partial mask: exactly the named lanes meet — lanes outside the mask are somewhere else (e.g. already waiting at a
// __syncthreads), as CUDA's *_sync semantics require; naming a lane that never arrives deadlocks here as it does on the GPU.
inline void warp_converge(unsigned mask) {
    Warp* w = t_warp;
    if (mask == 0xffffffffu) {
        w->bar->arrive_and_wait();
        return;
    }
    const unsigned me = 1u << t_lane;
    const unsigned gen = __atomic_load_n(&w->sub_gen, __ATOMIC_ACQUIRE);
    const unsigned seen = __atomic_fetch_or(&w->sub_arrived, me, __ATOMIC_ACQ_REL) | me;
    if (seen == mask) {
        __atomic_store_n(&w->sub_arrived, 0u, __ATOMIC_RELEASE);
        __atomic_fetch_add(&w->sub_gen, 1u, __ATOMIC_ACQ_REL);
    } else {
        while (__atomic_load_n(&w->sub_gen, __ATOMIC_ACQUIRE) == gen) std::this_thread::yield();
    }
}
}  // namespace cpu_cuda

inline void __syncwarp(unsigned mask = 0xffffffffu) { cpu_cuda::warp_converge(mask); }

// value held by lane (lane ^ lane_mask); every lane of the warp takes part (full-mask use only)
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    static_assert(sizeof(T) <= sizeof(unsigned long long), "shuffle payload");
    auto* w = cpu_cuda::t_warp;
    unsigned long long raw = 0;
    __builtin_memcpy(&raw, &v, sizeof(T));
    w->xchg[cpu_cuda::t_lane] = raw;
    cpu_cuda::warp_converge(mask);
    raw = w->xchg[(cpu_cuda::t_lane ^ lane_mask) & 31];
    cpu_cuda::warp_converge(mask);
    T r;
    __builtin_memcpy(&r, &raw, sizeof(T));
    return r;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v) {
    float old = *p, next;
    do { next = old + v; } while (!__atomic_compare_exchange(p, &old, &next, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
inline double atomicAdd(double* p, double v) {
    double old = *p, next;
    do { next = old + v; } while (!__atomic_compare_exchange(p, &old, &next, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
// round-to-nearest single operations that the compiler must not contract (build with -ffp-contract=off as well)
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
using std::floor;
inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
inline double rsqrt(double v) { return 1.0 / sqrt(v); }
inline float __expf(float v) { return expf(v); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float v) { return 1.0f / v; }
inline float exp2f_(float v) { return exp2f(v); }
inline float __int_as_float(int v) { float r; __builtin_memcpy(&r, &v, 4); return r; }
inline int __float_as_int(float v) { int r; __builtin_memcpy(&r, &v, 4); return r; }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
using std::min;
using std::max;

namespace cpu_cuda {

// kernel<<<grid, block>>>(args...)  ==  launch(grid, block, [&] { kernel(args...); })
// One pool of blockDim OS threads per launch walks over the blocks in order; a launch-wide barrier separates two
// blocks (the `static` shared memory of a kernel belongs to one block at a time).
template <typename F>
void launch(dim3 grid, dim3 block, F body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    blockDim = block;
    gridDim = grid;
    const int64_t nblocks = (int64_t)grid.x * grid.y * grid.z;
    std::barrier<> between_blocks(nthreads);
    Block blk[2];                                   // block i uses blk[i & 1]; re-armed by thread 0 while nobody uses it
    auto arm = [&](Block& b) {
        b.bar = std::make_unique<std::barrier<>>(nthreads);
        b.or_flag[0] = 0;
        b.or_flag[1] = 0;
        b.warps.clear();
        b.warps.resize(nwarps);
        for (int wi = 0; wi < nwarps; ++wi) {
            b.warps[wi].bar = std::make_unique<std::barrier<>>(std::min(32, nthreads - 32 * wi));
            std::fill(b.warps[wi].pred, b.warps[wi].pred + 32, 0u);
            std::fill(b.warps[wi].val, b.warps[wi].val + 32, 0);
        }
    };
    arm(blk[0]);
    std::vector<std::thread> threads;
    threads.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        threads.emplace_back([&, t] {
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (int64_t i = 0; i < nblocks; ++i) {
                Block& b = blk[i & 1];
                if (t == 0 && i + 1 < nblocks) arm(blk[(i + 1) & 1]);      // idle since the barrier that ended block i - 1
                blockIdx = dim3((unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((int64_t)grid.x * grid.y)));
                t_block = &b;
                t_warp = &b.warps[t / 32];
                t_lane = t % 32;
                t_or_phase = 0;
                body();
                // the thread has exited the kernel: it no longer takes part in any collective of this block
                t_warp->pred[t_lane] = 0;
                t_warp->val[t_lane] = 0;
                t_warp->bar->arrive_and_drop();
                b.bar->arrive_and_drop();
                between_blocks.arrive_and_wait();
            }
        });
    for (auto& th : threads) th.join();
}

}  // namespace cpu_cuda

namespace cpu_cuda {
// kernel<<<grid, block, smem_bytes>>>(args...) for kernels that declare PDB_DYNAMIC_SMEM(type, name)
template <typename F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    std::vector<float> smem(smem_bytes / sizeof(float) + 4);
    g_dyn_smem = smem.data();
    launch(grid, block, body);
    g_dyn_smem = nullptr;
}
}  // namespace cpu_cuda

#define PDB_DYNAMIC_SMEM(type, name) type* name = reinterpret_cast<type*>(cpu_cuda::g_dyn_smem)
