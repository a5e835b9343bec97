// Column sums of a short, wide matrix: the bias gradients of the decoder's Linear layers (db = sum over the B * Q = 200 rows
// of dy; mask2former_transformer_decoder.py:148-208 via autograd) — 133 launches of ATen's generic reduce_kernel at ~10 us each
// in the C2 step (profiles/r02_step_profile_c2_v2.txt).  One CTA per 32 columns, 8 row lanes x 32 columns, rows strided by 8,
// 8 partial sums combined through shared memory: every warp load is one 128-byte line.  Tall matrices (the cross-attention k / v
// projections over 2 x 16384 memory tokens, the convolutions' 131072 pixels) take col_sum_tall_kernel: float4 lanes (a warp reads
// 512 contiguous bytes of a row), four rows in flight per warp, row blocks over blockIdx.y whose partial sums meet in `out`
// through red.add.
#include <algorithm>
#include <cuda_bf16.h>
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

__global__ void __launch_bounds__(256)
col_sum_tall_kernel(const float4* __restrict__ x, float* __restrict__ out, int rows, int N4, int rows_per_block) {
    __shared__ float4 part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c4 = blockIdx.x * 32 + lane;
    const int r_begin = blockIdx.y * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < N4) {
        for (int r = r_begin + warp; r < r_end; r += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = r + 8 * u;
                if (rr < r_end) {
                    const float4 v = __ldg(x + (int64_t)rr * N4 + c4);
                    a[u].x += v.x, a[u].y += v.y, a[u].z += v.z, a[u].w += v.w;
                }
            }
        }
    }
    part[warp][lane] = make_float4(a[0].x + a[1].x + a[2].x + a[3].x, a[0].y + a[1].y + a[2].y + a[3].y,
                                   a[0].z + a[1].z + a[2].z + a[3].z, a[0].w + a[1].w + a[2].w + a[3].w);
    __syncthreads();
    if (warp == 0 && c4 < N4) {
        float4 s = part[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) s.x += part[i][lane].x, s.y += part[i][lane].y, s.z += part[i][lane].z, s.w += part[i][lane].w;
        float* o = out + 4 * c4;
        atomicAdd(o, s.x), atomicAdd(o + 1, s.y), atomicAdd(o + 2, s.z), atomicAdd(o + 3, s.w);
    }
}

// bf16 rows (the output gradients of the autocast path's Linear layers): a lane owns two adjacent columns, a warp 64 columns
// (128 contiguous bytes of a row), four rows in flight per warp, fp32 sums, row blocks over blockIdx.y meeting through red.add.
__global__ void __launch_bounds__(256)
col_sum_bf16_kernel(const __nv_bfloat162* __restrict__ x, float* __restrict__ out, int rows, int N2, int rows_per_block) {
    __shared__ float2 part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c2 = blockIdx.x * 32 + lane;
    const int r_begin = blockIdx.y * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
    float2 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = make_float2(0.f, 0.f);
    if (c2 < N2) {
        for (int r = r_begin + warp; r < r_end; r += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = r + 8 * u;
                if (rr < r_end) {
                    const float2 v = __bfloat1622float2(x[(int64_t)rr * N2 + c2]);
                    a[u].x += v.x, a[u].y += v.y;
                }
            }
        }
    }
    part[warp][lane] = make_float2(a[0].x + a[1].x + a[2].x + a[3].x, a[0].y + a[1].y + a[2].y + a[3].y);
    __syncthreads();
    if (warp == 0 && c2 < N2) {
        float2 s = part[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) s.x += part[i][lane].x, s.y += part[i][lane].y;
        atomicAdd(out + 2 * c2, s.x), atomicAdd(out + 2 * c2 + 1, s.y);
    }
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_col_sum(const float* x, float* out, int rows, int N, int accumulate, void* stream) {
    PDB_REQUIRE(x && out, "col_sum: null pointer");
    PDB_REQUIRE(rows > 0 && N > 0, "col_sum: non-positive size");
    if (rows > 512 && N % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int col_blocks = (N / 4 + 31) / 32;
        int row_blocks = std::max(1, std::min((rows + 127) / 128, (4 * 148 + col_blocks - 1) / col_blocks));
        const int per = (rows + row_blocks - 1) / row_blocks;
        row_blocks = (rows + per - 1) / per;
        if (!accumulate)
            PDB_REQUIRE(cudaMemsetAsync(out, 0, sizeof(float) * N, as_stream(stream)) == cudaSuccess, "col_sum: memset failed");
        col_sum_tall_kernel<<<dim3((unsigned)col_blocks, (unsigned)row_blocks), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(x), out, rows, N / 4, per);
        return launched("col_sum_tall");
    }
    col_sum_kernel<<<(unsigned)((N + 31) / 32), 256, 0, as_stream(stream)>>>(x, out, rows, N, accumulate);
    return launched("col_sum");
}

// out[n] (+)= sum_r x[r*N + n] for a bf16 matrix with an even number of columns; fp32 sums.
extern "C" int pdb_col_sum_bf16(const void* x, float* out, int rows, int N, int accumulate, void* stream) {
    PDB_REQUIRE(x && out, "col_sum_bf16: null pointer");
    PDB_REQUIRE(rows > 0 && N > 0 && N % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0, "col_sum_bf16: N must be even, x 4-byte aligned");
    const int col_blocks = (N / 2 + 31) / 32;
    int row_blocks = std::max(1, std::min((rows + 127) / 128, (4 * 148 + col_blocks - 1) / col_blocks));
    const int per = (rows + row_blocks - 1) / row_blocks;
    row_blocks = (rows + per - 1) / per;
    if (!accumulate)
        PDB_REQUIRE(cudaMemsetAsync(out, 0, sizeof(float) * N, as_stream(stream)) == cudaSuccess, "col_sum_bf16: memset failed");
    col_sum_bf16_kernel<<<dim3((unsigned)col_blocks, (unsigned)row_blocks), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat162*>(x), out, rows, N / 2, per);
    return launched("col_sum_bf16");
}
