// Mask-head einsum on the 5th-generation tensor cores:  out[b,q,p] = sum_c embed[b,q,c] * feat[b,p,c]
// (reference: torch.einsum("bqc,bchw->bqhw"), mask2former_transformer_decoder.py:449), with the feature
// map stored pixel-major (B, HW, C) so that both operands are K-major.
//
// fp32 accuracy on TF32 tensor cores ("3xTF32"): every fp32 operand x is split into hi = x with the low
// 13 mantissa bits cleared (exactly representable in tf32) and lo = x - hi (exact in fp32); the product
// is accumulated in fp32 TMEM as  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term is ~2^-22 relative).
//
// One CTA computes a 128-pixel x BN-query tile (BN = Q rounded up to 16, <= 128):
//   warp 0      TMA producer: 128x32 fp32 feature tile + BNx32 embed tile per k-block, 128B-swizzled
//   warps 2..5  split the landed tiles in shared memory into hi (in place) / lo (second buffer); the split is
//               element-wise, hence independent of the swizzle; then they drain the accumulator (epilogue)
//   warp 1      allocates TMEM and issues tcgen05.mma kind::tf32 (M=128, N=BN, K=8), 3 terms x 4 per k-block
// Pipeline: full[s] (TMA bytes landed) -> split[s] (hi/lo ready) -> tcgen05.commit -> empty[s].
// The accumulator (128 lanes x BN columns of TMEM) is read with tcgen05.ld 32x32b: lane = pixel, so
// for a fixed query the 32 lanes of a warp store 32 consecutive pixels (128-byte coalesced rows of out).
#include "common.cuh"
#include "tc_gemm.cuh"

namespace pdb {

constexpr int TC_BM = 128;          // pixels per tile
constexpr int TC_BK = 32;           // fp32 per 128-byte swizzled row
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 192;

template <int BN>
struct EinsumSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;      // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int TOTAL = TC_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void split_chunk(float4* hi_ptr, float4* lo_ptr) {
    float4 x = *hi_ptr;
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
    *hi_ptr = h;
    *lo_ptr = l;
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
mask_einsum_tc_kernel(const __grid_constant__ CUtensorMap tm_feat, const __grid_constant__ CUtensorMap tm_embed,
                      float* __restrict__ out, int Q, int HW, int C, int tiles_per_image) {
    using S = EinsumSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * S::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* split = bars + TC_STAGES;
    uint64_t* empty = bars + 2 * TC_STAGES;
    uint64_t* accum = bars + 3 * TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x / tiles_per_image;
    const int tile = blockIdx.x - b * tiles_per_image;
    const int pix0 = tile * TC_BM;
    const int q0 = blockIdx.y * BN;                 // query tile (Q > BN: several CTAs per pixel tile)
    const int num_kb = (C + TC_BK - 1) / TC_BK;

    auto a_hi = [&](int s) { return smem + s * S::STAGE_BYTES; };
    auto a_lo = [&](int s) { return smem + s * S::STAGE_BYTES + S::A_BYTES; };
    auto b_hi = [&](int s) { return smem + s * S::STAGE_BYTES + 2 * S::A_BYTES; };
    auto b_lo = [&](int s) { return smem + s * S::STAGE_BYTES + 2 * S::A_BYTES + S::B_BYTES; };

    if (threadIdx.x == 0) {
        tc::prefetch_tensormap(&tm_feat);
        tc::prefetch_tensormap(&tm_embed);
        for (int s = 0; s < TC_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&split[s], 128);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(accum, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<128>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                tc::mbar_expect_tx(&full[s], S::A_BYTES + S::B_BYTES);
                tc::tma_load_2d(a_hi(s), &tm_feat, &full[s], kb * TC_BK, b * HW + pix0);
                tc::tma_load_2d(b_hi(s), &tm_embed, &full[s], kb * TC_BK, b * Q + q0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_tf32(TC_BM, BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::mbar_wait(&split[s], ph);
                tc::tc_fence_after();
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint32_t koff = k * 8 * 4;
                    const uint64_t dah = tc::umma_desc_k_sw128(a_hi(s), koff), dal = tc::umma_desc_k_sw128(a_lo(s), koff);
                    const uint64_t dbh = tc::umma_desc_k_sw128(b_hi(s), koff), dbl = tc::umma_desc_k_sw128(b_lo(s), koff);
                    tc::mma_tf32(tmem_d, dal, dbh, idesc, (kb | k) != 0);      // small terms first
                    tc::mma_tf32(tmem_d, dah, dbl, idesc, 1);
                    tc::mma_tf32(tmem_d, dah, dbh, idesc, 1);
                }
                tc::tc_commit(&empty[s]);          // frees the stage once these MMAs have read it
            }
            tc::tc_commit(accum);
        }
    } else {
        // ---- hi/lo split of every landed stage (warps 2..5 = 128 threads)
        const int t = threadIdx.x - 64;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % TC_STAGES;
            const uint32_t ph = (kb / TC_STAGES) & 1;
            tc::mbar_wait(&full[s], ph);
            float4* ah = reinterpret_cast<float4*>(a_hi(s));
            float4* al = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
            for (int i = 0; i < S::A_BYTES / 16 / 128; ++i) split_chunk(ah + i * 128 + t, al + i * 128 + t);
            float4* bh = reinterpret_cast<float4*>(b_hi(s));
            float4* bl = reinterpret_cast<float4*>(b_lo(s));
            for (int i = t; i < S::B_BYTES / 16; i += 128) split_chunk(bh + i, bl + i);
            tc::fence_proxy_async();
            tc::mbar_arrive(&split[s]);
        }
        // ---- epilogue: TMEM -> registers -> global (transposed store: lane = pixel, column = query)
        tc::mbar_wait(accum, 0);
        tc::tc_fence_after();
        const int quarter = warp & 3;                          // TMEM lanes this warp may access
        const int pix = pix0 + quarter * 32 + lane;
        float* orow = out + (int64_t)b * Q * HW + pix;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + c0, v);
            if (pix < HW) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (q0 + c0 + i < Q) orow[(int64_t)(q0 + c0 + i) * HW] = v[i];
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<128>(tmem_d);
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                           uint32_t box_inner, uint32_t box_outer) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(PDB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {outer_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PDB_ERR_INVALID, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PDB_OK;
}

template <int BN>
static int launch_einsum_tc(const float* embed, const float* feat_pm, float* out, int B, int Q, int C, int64_t HW,
                            cudaStream_t st) {
    using S = EinsumSmem<BN>;
    CUtensorMap tm_feat, tm_embed;
    PDB_TRY(make_tensor_map_2d_f32(&tm_feat, feat_pm, (uint64_t)C, (uint64_t)B * HW, (uint64_t)C * 4, TC_BK, TC_BM));
    PDB_TRY(make_tensor_map_2d_f32(&tm_embed, embed, (uint64_t)C, (uint64_t)B * Q, (uint64_t)C * 4, TC_BK, BN));
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mask_einsum_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "mask_einsum_tc: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    int tiles = (int)((HW + TC_BM - 1) / TC_BM);
    dim3 grid((unsigned)(B * tiles), (unsigned)((Q + BN - 1) / BN));
    mask_einsum_tc_kernel<BN><<<grid, TC_THREADS, S::TOTAL, st>>>(tm_feat, tm_embed, out, Q, (int)HW, C, tiles);
    return launched("mask_einsum_tc");
}

int mask_einsum_forward_tc(const float* embed, const float* feat_pm, float* out, int B, int Q, int C, int64_t HW,
                           cudaStream_t st) {
    if (Q <= 16) return launch_einsum_tc<16>(embed, feat_pm, out, B, Q, C, HW, st);
    if (Q <= 32) return launch_einsum_tc<32>(embed, feat_pm, out, B, Q, C, HW, st);
    if (Q <= 64) return launch_einsum_tc<64>(embed, feat_pm, out, B, Q, C, HW, st);
    if (Q <= 112 || (Q > 128 && Q <= 224)) return launch_einsum_tc<112>(embed, feat_pm, out, B, Q, C, HW, st);
    return launch_einsum_tc<128>(embed, feat_pm, out, B, Q, C, HW, st);
}

}  // namespace pdb
