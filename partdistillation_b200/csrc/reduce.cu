// Column sums of a short, wide matrix: the bias gradients of the decoder's Linear layers (db = sum over the B * Q = 200 rows
// of dy; mask2former_transformer_decoder.py:148-208 via autograd) — 133 launches of ATen's generic reduce_kernel at ~10 us each
// in the C2 step (profiles/r02_step_profile_c2_v2.txt).  One CTA per 32 columns, 8 row lanes x 32 columns, rows strided by 8,
// 8 partial sums combined through shared memory: every warp load is one 128-byte line.
#include "common.cuh"

namespace pdb {

__global__ void __launch_bounds__(256)
col_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int N, int accumulate) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r0 = threadIdx.x >> 5;
    float a0 = 0.f, a1 = 0.f;
    if (c < N) {
        int r = r0;
        for (; r + 8 < rows; r += 16) {            // two independent chains
            a0 += __ldg(x + (int64_t)r * N + c);
            a1 += __ldg(x + (int64_t)(r + 8) * N + c);
        }
        if (r < rows) a0 += __ldg(x + (int64_t)r * N + c);
    }
    part[r0][threadIdx.x & 31] = a0 + a1;
    __syncthreads();
    if (r0 == 0 && c < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
        out[c] = accumulate ? out[c] + s : s;
    }
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_col_sum(const float* x, float* out, int rows, int N, int accumulate, void* stream) {
    PDB_REQUIRE(x && out, "col_sum: null pointer");
    PDB_REQUIRE(rows > 0 && N > 0, "col_sum: non-positive size");
    col_sum_kernel<<<(unsigned)((N + 31) / 32), 256, 0, as_stream(stream)>>>(x, out, rows, N, accumulate);
    return launched("col_sum");
}
