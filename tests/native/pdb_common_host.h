// TEST INFRASTRUCTURE ONLY.  Host restatement of the few device helpers of partdistillation_b200/csrc/common.cuh that the
// kernel files use (common.cuh itself includes the CUDA runtime and inline PTX, so the host builds cannot include it).
#pragma once
#include "cuda_on_cpu.h"

#define PDB_OK 0
#define PDB_REQUIRE(cond, ...) do { if (!(cond)) return -1; } while (0)
#define PDB_TRY(expr) do { int _rc = (expr); if (_rc != PDB_OK) return _rc; } while (0)

namespace pdb {
constexpr int kNumSMs = 148;
constexpr int kMaxLevels = 8;
struct LevelTable { int h[kMaxLevels]; int w[kMaxLevels]; int start[kMaxLevels]; };
inline float warp_sum(float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
inline float warp_max(float v) { for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
// red.global.add.v4.f32: four float reductions into global memory
inline void red_add_v4(float* addr, float a, float b, float c, float d) {
    atomicAdd(addr, a); atomicAdd(addr + 1, b); atomicAdd(addr + 2, c); atomicAdd(addr + 3, d);
}
}  // namespace pdb
