// bf16 GEMM on the 5th-generation tensor cores (tcgen05 kind::f16, fp32 accumulation in tensor memory) — the dense
// contraction engine of the AUTOCAST path (BASELINE configs[2]: PartDistillation under bf16 autocast; the reference trains
// under AMP, configs/mask2former/coco/instance-segmentation/Base-COCO-InstanceSegmentation.yaml:34-35, where nn.Linear of the
// Swin backbone and of the transformer decoder run in the low-precision dtype through cuBLAS):
//     C[m][n] = sum_k A[m][k] * B[n][k]  (+ bias[n]) (ReLU | GELU),   A (M x K) and B (N x K) bf16, K contiguous,
//     C fp32 or bf16 row-major.
// nn.Linear forward is A = x, B = W; the input gradient dx = dy W is A = dy, B = W^T; the weight gradient dW = dy^T x is
// A = dy^T, B = x^T (functional.py materialises the transposes: decoder-sized operands, and the frozen backbone has no
// backward).  No hi / lo split, 2-byte tiles: one 128 x 128 x 64 k-block is 32 KB of shared memory and four UMMA_K = 16
// instructions, against 48-64 KB and twelve for the 3xTF32 kernel of gemm_tc.cu.
//
// Persistent kernel, one CTA per SM, 6 warps:
//   warp 0      TMA producer: A / B k-blocks (128 rows x 128 bytes, SWIZZLE_128B) into a 6-deep ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  epilogue, one per TMEM lane quarter: tcgen05.ld -> (+bias, activation) -> global stores, overlapped with the
//               next tile's main loop through a double-buffered accumulator (2 x 128 columns)
// TMA zero-fills out-of-range rows / k, so ragged M, N, K need no masking on the operand side (K % 8 == 0: 16-byte row pitch).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace pdb {

constexpr int H_BM = 128, H_BN = 128, H_BK = 64;       // bf16 elements: one k-block row = 128 bytes
constexpr int H_STAGES = 6;
constexpr int H_THREADS = 6 * 32;
constexpr int H_STAGE_BYTES = (H_BM + H_BN) * H_BK * 2;      // 32 KB
constexpr int H_SMEM = H_STAGES * H_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct HParams {
    void* C;
    const float* bias;
    int M, N, K;
    int64_t ldc;
    int out_bf16, act;
    int nt, total_tiles, num_kb;
};

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float h_act(float x, int mode) {
    if (mode == 1) return fmaxf(x, 0.f);
    if (mode == 2) return x * 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    return x;
}

__global__ void __launch_bounds__(H_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const HParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + H_STAGES * H_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + H_STAGES;
    uint64_t* acc_full = bars + 2 * H_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tc::prefetch_tensormap(&tm_a);
        tc::prefetch_tensormap(&tm_b);
        for (int s = 0; s < H_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], 4 * 32);
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<256>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.nt) * H_BM, n0 = (tile % p.nt) * H_BN;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    tc::mbar_expect_tx(&full[s], H_STAGE_BYTES);
                    uint8_t* a = smem + s * H_STAGE_BYTES;
                    tc::tma_load_2d(a, &tm_a, &full[s], kb * H_BK, m0);
                    tc::tma_load_2d(a + H_BM * H_BK * 2, &tm_b, &full[s], kb * H_BK, n0);
                    if (++s == H_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(H_BM, H_BN);
            int s = 0;
            uint32_t ph = 0, t = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
                const uint32_t ab = t & 1;
                tc::mbar_wait(&acc_empty[ab], ((t >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t tmem_d = tmem_base + ab * H_BN;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    tc::mbar_wait(&full[s], ph);
                    tc::tc_fence_after();
                    const uint8_t* a = smem + s * H_STAGE_BYTES;
                    const uint8_t* b = a + H_BM * H_BK * 2;
#pragma unroll
                    for (int k = 0; k < H_BK / 16; ++k)          // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
                        mma_bf16(tmem_d, tc::umma_desc_k_sw128(a, k * 32), tc::umma_desc_k_sw128(b, k * 32), idesc, (kb | k) != 0);
                    tc::tc_commit(&empty[s]);
                    if (++s == H_STAGES) { s = 0; ph ^= 1u; }
                }
                tc::tc_commit(&acc_full[ab]);
            }
        }
    } else {
        // epilogue: this warp may touch TMEM lanes 32 * (warp % 4) .. + 31 = rows of the tile
        const int quarter = warp & 3;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
            const int m0 = (tile / p.nt) * H_BM, n0 = (tile % p.nt) * H_BN;
            const uint32_t ab = t & 1;
            tc::mbar_wait(&acc_full[ab], (t >> 1) & 1);
            tc::tc_fence_after();
            const int row = m0 + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ab * H_BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < H_BN / 16; ++c) {
                const int col0 = n0 + c * 16;
                if (col0 >= p.N) break;                          // warp-uniform
                float v[16];
                tc::tmem_ld16(taddr + c * 16, v);
                if (row < p.M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float bj = (p.bias && col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f;
                        v[j] = h_act(v[j] + bj, p.act);
                    }
                    const bool full16 = col0 + 16 <= p.N;
                    if (p.out_bf16) {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + col0;
                        if (full16 && (p.ldc & 7) == 0) {
                            uint32_t w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                                w[j] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        } else {
                            for (int j = 0; j < 16 && col0 + j < p.N; ++j) o[j] = __float2bfloat16_rn(v[j]);
                        }
                    } else {
                        float* o = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
                        if (full16 && (p.ldc & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
                            for (int j = 0; j < 16 && col0 + j < p.N; ++j) o[j] = v[j];
                        }
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_empty[ab]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<256>(tmem_base);
}

typedef CUresult (*EncodeTiledFnH)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (rows x K) bf16, K contiguous, row pitch ld elements: box (64, 128), 128-byte swizzle
static int make_map_bf16(CUtensorMap* map, const void* base, int rows, int K, int64_t ld) {
    static EncodeTiledFnH encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(PDB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
        encode = reinterpret_cast<EncodeTiledFnH>(fn);
    }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)H_BK, (cuuint32_t)H_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(PDB_ERR_INVALID, "gemm_bf16: cuTensorMapEncodeTiled failed (%d): %d x %d, ld %lld", (int)r, rows, K, (long long)ld);
    return PDB_OK;
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_gemm_bf16(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int64_t lda,
                             int64_t ldb, int64_t ldc, int act, int out_bf16, void* stream) {
    PDB_REQUIRE(A && B && C, "gemm_bf16: null pointer");
    PDB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: non-positive dimension");
    PDB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm_bf16: K, lda, ldb must be multiples of 8 (16-byte rows)");
    PDB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm_bf16: A and B must be 16-byte aligned");
    PDB_REQUIRE(act >= 0 && act <= 2, "gemm_bf16: activation %d (0 none, 1 ReLU, 2 GELU)", act);
    CUtensorMap ta, tb;
    PDB_TRY(make_map_bf16(&ta, A, M, K, lda));
    PDB_TRY(make_map_bf16(&tb, B, N, K, ldb));
    HParams p;
    p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.out_bf16 = out_bf16; p.act = act;
    p.nt = (N + H_BN - 1) / H_BN;
    p.total_tiles = ((M + H_BM - 1) / H_BM) * p.nt;
    p.num_kb = (K + H_BK - 1) / H_BK;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "gemm_bf16: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    gemm_bf16_kernel<<<grid, H_THREADS, H_SMEM, as_stream(stream)>>>(ta, tb, p);
    return launched("gemm_bf16");
}
