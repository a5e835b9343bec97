// bf16 GEMM on the 5th-generation tensor cores (tcgen05 kind::f16, fp32 accumulation in tensor memory) — the dense
// contraction engine of the AUTOCAST path (BASELINE configs[2]: PartDistillation under bf16 autocast; the reference trains
// under AMP, configs/mask2former/coco/instance-segmentation/Base-COCO-InstanceSegmentation.yaml:34-35, where nn.Linear of the
// Swin backbone and of the transformer decoder run in the low-precision dtype through cuBLAS):
//     C[m][n] = sum_k A[m][k] * B[n][k]  (+ bias[n]) (ReLU | GELU),   A (M x K) and B (N x K) bf16, K contiguous,
//     C fp32 or bf16 row-major.
// nn.Linear forward is A = x, B = W; the input gradient dx = dy W is A = dy, B = W^T (cached); the weight gradient
// dW = dy^T x reads dy and x IN PLACE as MN-major operands (layout bit 1 / 2: A(m, k) = A[k][m], B(n, k) = B[k][n], i.e. the
// contraction index is the row index of the stored matrices): TMA boxes of 64 mn-elements x 64 k-rows, UMMA descriptors with
// the MN-major 128-byte-swizzle canonical layout (64-element mn blocks LBO = 8 KB apart, 8-row k groups SBO = 1 KB apart).  No hi / lo split, 2-byte tiles: one 128 x 128 x 64 k-block is 32 KB of shared memory and four UMMA_K = 16
// instructions, against 48-64 KB and twelve for the 3xTF32 kernel of gemm_tc.cu.
//
// Persistent kernel, one CTA per SM, 6 warps:
//   warp 0      TMA producer: A / B k-blocks (128 rows x 128 bytes, SWIZZLE_128B) into a 5-deep ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..9  epilogue, two per TMEM lane quarter (64 columns each): four tcgen05.ld in flight -> (+bias, activation) ->
//               shared-memory transpose -> coalesced 128-byte row segments (a thread holds 64 outputs of ONE row; storing them
//               directly costs one L1 wavefront per lane), overlapped with the next tile's main loop through a double-buffered
//               accumulator (2 x 128 columns)
// TMA zero-fills out-of-range rows / k, so ragged M, N, K need no masking on the operand side (K % 8 == 0: 16-byte row pitch).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace pdb {

constexpr int H_BM = 128, H_BN = 128, H_BK = 64;       // bf16 elements: one k-block row = 128 bytes
constexpr int H_STAGES = 5;
constexpr int H_EPI_WARPS = 8;                         // two per TMEM lane quarter, each takes 64 of the 128 columns
constexpr int H_THREADS = (2 + H_EPI_WARPS) * 32;
constexpr int H_STAGE_BYTES = (H_BM + H_BN) * H_BK * 2;      // 32 KB
constexpr int H_STG_WORDS = 32 * 36;                    // per epilogue warp: 32 rows x (32 + 4 pad) words
constexpr int H_SMEM = H_STAGES * H_STAGE_BYTES + 256 /*barriers*/ + H_EPI_WARPS * H_STG_WORDS * 4 + 1024 /*align*/;

struct HParams {
    void* C;
    const float* bias;
    int M, N, K;
    int64_t ldc;
    int out_bf16, act;
    int nt, total_tiles, num_kb;
    int accumulate;                         // C += instead of C = (red.add); implied by ksplit > 1
    int mn_tiles, ksplit, kb_per_split;     // split-K (weight gradients: few output tiles, very long K): tile = split * mn_tiles + mn,
                                            // fp32 partial sums meet in a zero-filled C through red.add
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

// GELU (erf form) for a bf16 result: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the 2^-9 of the output
// rounding) on the fast exp / reciprocal units — erff() alone would make the epilogue longer than the tile's MMAs.
__device__ __forceinline__ float fast_erf(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float r = 1.f - poly * t * __expf(-ax * ax);
    return copysignf(r, x);
}
__device__ __forceinline__ float h_act(float x, int mode) {
    if (mode == 1) return fmaxf(x, 0.f);
    if (mode == 2) return x * 0.5f * (1.f + fast_erf(x * 0.70710678118654752440f));
    return x;
}

// MN-major operand tile: [mn block j of 64][k row r of 64][64 mn elements = 128 bytes, swizzled]; one UMMA_K = 16 step = 2 KB
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_h(const void* smem_tile, uint32_t byte_offset) {
    uint32_t addr = tc::smem_u32(smem_tile) + byte_offset;
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((8192 >> 4) & 0x3FFF) << 16;     // LBO: next 64-element block along MN
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;     // SBO: next group of 8 k rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}

template <bool A_MN, bool B_MN>
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
            tc::mbar_init(&acc_empty[s], H_EPI_WARPS * 32);
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
                const int mn = tile % p.mn_tiles, kb0 = (tile / p.mn_tiles) * p.kb_per_split;
                const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                const int m0 = (mn / p.nt) * H_BM, n0 = (mn % p.nt) * H_BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    tc::mbar_expect_tx(&full[s], H_STAGE_BYTES);
                    uint8_t* a = smem + s * H_STAGE_BYTES;
                    uint8_t* b = a + H_BM * H_BK * 2;
                    if (A_MN) {
                        tc::tma_load_2d(a, &tm_a, &full[s], m0, kb * H_BK);
                        tc::tma_load_2d(a + 8192, &tm_a, &full[s], m0 + 64, kb * H_BK);
                    } else {
                        tc::tma_load_2d(a, &tm_a, &full[s], kb * H_BK, m0);
                    }
                    if (B_MN) {
                        tc::tma_load_2d(b, &tm_b, &full[s], n0, kb * H_BK);
                        tc::tma_load_2d(b + 8192, &tm_b, &full[s], n0 + 64, kb * H_BK);
                    } else {
                        tc::tma_load_2d(b, &tm_b, &full[s], kb * H_BK, n0);
                    }
                    if (++s == H_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(H_BM, H_BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
            int s = 0;
            uint32_t ph = 0, t = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
                const uint32_t ab = t & 1;
                tc::mbar_wait(&acc_empty[ab], ((t >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t tmem_d = tmem_base + ab * H_BN;
                const int kb0 = (tile / p.mn_tiles) * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait(&full[s], ph);
                    tc::tc_fence_after();
                    const uint8_t* a = smem + s * H_STAGE_BYTES;
                    const uint8_t* b = a + H_BM * H_BK * 2;
#pragma unroll
                    for (int k = 0; k < H_BK / 16; ++k)          // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
                        mma_bf16(tmem_d, A_MN ? umma_desc_mn_sw128_h(a, k * 2048) : tc::umma_desc_k_sw128(a, k * 32),
                                 B_MN ? umma_desc_mn_sw128_h(b, k * 2048) : tc::umma_desc_k_sw128(b, k * 32), idesc, ((kb - kb0) | k) != 0);
                    tc::tc_commit(&empty[s]);
                    if (++s == H_STAGES) { s = 0; ph ^= 1u; }
                }
                tc::tc_commit(&acc_full[ab]);
            }
        }
    } else {
        // epilogue: this warp may touch TMEM lanes 32 * (warp % 4) .. + 31 = rows of the tile; it takes columns 64 * half .. + 63
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
            const int mn = tile % p.mn_tiles;
            const int m0 = (mn / p.nt) * H_BM, n0 = (mn % p.nt) * H_BN;
            const bool split = p.accumulate != 0, first_split = tile < p.mn_tiles;
            const uint32_t ab = t & 1;
            tc::mbar_wait(&acc_full[ab], (t >> 1) & 1);
            tc::tc_fence_after();
            const int row = m0 + quarter * 32 + lane;
            const int col0 = n0 + half * 64;
            const uint32_t taddr = tmem_base + ab * H_BN + half * 64 + ((uint32_t)(quarter * 32) << 16);
            uint32_t r[4][16];
#pragma unroll
            for (int c = 0; c < 4; ++c) tc::tmem_ld16_nowait(taddr + c * 16, r[c]);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_empty[ab]);                     // the accumulator is in registers: the MMA warp may reuse it
            const bool full64 = col0 + 64 <= p.N;
            const bool fast = full64 && (p.out_bf16 ? (p.ldc & 7) == 0 : (p.ldc & 3) == 0) && ((uintptr_t)p.C & 15) == 0;
            if (col0 < p.N) {
                // bias + activation in registers (thread = one row of the tile, 64 columns)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int cc = col0 + c * 16;
                    if (p.bias && first_split) {
                        if (full64) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cc) + j);
                                r[c][4 * j] = __float_as_uint(__uint_as_float(r[c][4 * j]) + b4.x);
                                r[c][4 * j + 1] = __float_as_uint(__uint_as_float(r[c][4 * j + 1]) + b4.y);
                                r[c][4 * j + 2] = __float_as_uint(__uint_as_float(r[c][4 * j + 2]) + b4.z);
                                r[c][4 * j + 3] = __float_as_uint(__uint_as_float(r[c][4 * j + 3]) + b4.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                r[c][j] = __float_as_uint(__uint_as_float(r[c][j]) + ((cc + j < p.N) ? __ldg(p.bias + cc + j) : 0.f));
                        }
                    }
                    if (p.act) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[c][j] = __float_as_uint(h_act(__uint_as_float(r[c][j]), p.act));
                    }
                }
                if (fast) {
                    uint32_t* stg = reinterpret_cast<uint32_t*>(smem + H_STAGES * H_STAGE_BYTES + 256) + (warp - 2) * H_STG_WORDS;
                    const int rr = lane >> 3, seg = lane & 7;
                    if (p.out_bf16) {
                        // 64 bf16 = 32 words per row: stage, then every warp instruction writes 4 rows x 128 contiguous bytes
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[c][2 * j]), __uint_as_float(r[c][2 * j + 1]));
                                w[j] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            *reinterpret_cast<uint4*>(stg + lane * 36 + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                            *reinterpret_cast<uint4*>(stg + lane * 36 + c * 8 + 4) = make_uint4(w[4], w[5], w[6], w[7]);
                        }
                        __syncwarp();
                        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.C) + col0 + seg * 8;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int lr = it * 4 + rr;
                            const int gr = m0 + quarter * 32 + lr;
                            const uint4 val = *reinterpret_cast<const uint4*>(stg + lr * 36 + seg * 4);
                            if (gr < p.M) *reinterpret_cast<uint4*>(base + (int64_t)gr * p.ldc) = val;
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {             // 32 fp32 columns = 32 words per row and pass
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    *reinterpret_cast<uint4*>(stg + lane * 36 + c * 16 + j * 4) =
                                        make_uint4(r[h2 * 2 + c][4 * j], r[h2 * 2 + c][4 * j + 1], r[h2 * 2 + c][4 * j + 2], r[h2 * 2 + c][4 * j + 3]);
                            __syncwarp();
                            float* base = reinterpret_cast<float*>(p.C) + col0 + h2 * 32 + seg * 4;
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int lr = it * 4 + rr;
                                const int gr = m0 + quarter * 32 + lr;
                                const uint4 val = *reinterpret_cast<const uint4*>(stg + lr * 36 + seg * 4);
                                if (gr < p.M) {
                                    float* dst = base + (int64_t)gr * p.ldc;
                                    if (split)
                                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(val.x)),
                                                     "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w)) : "memory");
                                    else
                                        *reinterpret_cast<uint4*>(dst) = val;
                                }
                            }
                            __syncwarp();
                        }
                    }
                } else if (row < p.M) {
                    // ragged column tile / unaligned rows: element-wise stores
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        const int cc = col0 + c * 16;
                        for (int j = 0; j < 16 && cc + j < p.N; ++j) {
                            float val = 0.f;
#pragma unroll
                            for (int c2 = 0; c2 < 4; ++c2)
#pragma unroll
                                for (int j2 = 0; j2 < 16; ++j2)
                                    if (c2 == c && j2 == j) val = __uint_as_float(r[c2][j2]);
                            if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.C)[(int64_t)row * p.ldc + cc + j] = __float2bfloat16_rn(val);
                            else if (split) atomicAdd(reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + cc + j, val);
                            else reinterpret_cast<float*>(p.C)[(int64_t)row * p.ldc + cc + j] = val;
                        }
                    }
                }
            }
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
static EncodeTiledFnH encode_tiled_entry() {
    static EncodeTiledFnH encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn) encode = reinterpret_cast<EncodeTiledFnH>(fn);
    }
    return encode;
}

static int make_map_bf16(CUtensorMap* map, const void* base, int rows, int K, int64_t ld) {
    EncodeTiledFnH encode = encode_tiled_entry();
    if (!encode) return fail(PDB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
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

// MN-major operand: the stored matrix is (K rows x mn), mn contiguous, row pitch ld elements: box (64 mn, 64 k rows)
static int make_map_bf16_mn(CUtensorMap* map, const void* base, int mn, int K, int64_t ld) {
    EncodeTiledFnH encode = encode_tiled_entry();
    if (!encode) return fail(PDB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)H_BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(PDB_ERR_INVALID, "gemm_bf16: cuTensorMapEncodeTiled (MN-major) failed (%d): %d x %d, ld %lld", (int)r, K, mn, (long long)ld);
    return PDB_OK;
}

template <bool A_MN, bool B_MN>
static int launch_bf16(const CUtensorMap& ta, const CUtensorMap& tb, const HParams& p, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "gemm_bf16: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    gemm_bf16_kernel<A_MN, B_MN><<<grid, H_THREADS, H_SMEM, st>>>(ta, tb, p);
    return launched("gemm_bf16");
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_gemm_bf16(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int64_t lda,
                             int64_t ldb, int64_t ldc, int act, int out_bf16, int ksplit, int accumulate, int layout, void* stream) {
    PDB_REQUIRE(A && B && C, "gemm_bf16: null pointer");
    PDB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: non-positive dimension");
    PDB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm_bf16: lda, ldb must be multiples of 8 (16-byte rows)");
    PDB_REQUIRE(layout == 0 || layout == 3, "gemm_bf16: layout %d (0: A, B K-major; 3: both MN-major)", layout);
    PDB_REQUIRE(layout != 0 || K % 8 == 0, "gemm_bf16: K must be a multiple of 8 for K-major operands");
    PDB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm_bf16: A and B must be 16-byte aligned");
    PDB_REQUIRE(act >= 0 && act <= 2, "gemm_bf16: activation %d (0 none, 1 ReLU, 2 GELU)", act);
    CUtensorMap ta, tb;
    if (layout == 3) {
        PDB_TRY(make_map_bf16_mn(&ta, A, M, K, lda));
        PDB_TRY(make_map_bf16_mn(&tb, B, N, K, ldb));
    } else {
        PDB_TRY(make_map_bf16(&ta, A, M, K, lda));
        PDB_TRY(make_map_bf16(&tb, B, N, K, ldb));
    }
    HParams p;
    p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.out_bf16 = out_bf16; p.act = act;
    p.nt = (N + H_BN - 1) / H_BN;
    p.mn_tiles = ((M + H_BM - 1) / H_BM) * p.nt;
    p.num_kb = (K + H_BK - 1) / H_BK;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > p.num_kb) ksplit = p.num_kb;
    if (ksplit > 1) accumulate = 1;
    PDB_REQUIRE(!accumulate || (!out_bf16 && act == 0), "gemm_bf16: split-K / accumulation needs an fp32 C and no activation");
    p.accumulate = accumulate;
    p.kb_per_split = (p.num_kb + ksplit - 1) / ksplit;
    p.ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.total_tiles = p.mn_tiles * p.ksplit;
    return layout == 3 ? launch_bf16<true, true>(ta, tb, p, as_stream(stream)) : launch_bf16<false, false>(ta, tb, p, as_stream(stream));
}
