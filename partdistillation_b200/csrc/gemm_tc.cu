// General fp32 GEMM on the 5th-generation tensor cores ("3xTF32"), the dense-contraction engine of the
// hot path:  for each batch b
//     C_b[m][n] (+)= sum_k A_b(m,k) * B_b(n,k)  (+ bias[n]) (ReLU)
// It serves nn.Linear forward / input-gradient / weight-gradient of the encoder and decoder layers
// (reference: msdeformattn.py:120-135, ops/modules/ms_deform_attn.py:102-130,
// mask2former_transformer_decoder.py:148-208) and the two gradient products of the mask-head einsum
// (mask2former_transformer_decoder.py:449), all of which cuBLAS runs as SIMT sgemm when TF32 is off.
//
// fp32 accuracy on TF32 tensor cores: every operand x is split into hi = x with the low 13 mantissa
// bits cleared (exact in tf32) and lo = x - hi (exact in fp32).  hi*hi accumulates in one fp32 TMEM
// accumulator and the correction lo*hi + hi*lo in a second one; the epilogue adds them (the dropped lo*lo
// term is ~2^-22 relative).  Two accumulators because the tensor core truncates when it aligns addends to the
// accumulator: adding the 2^-11-sized corrections into the main sum would cost one truncation each, tripling
// the (biased) rounding error; measured error vs fp64 ~1e-6 of max at K=256.
//
// Operand layouts are described to TMA / UMMA instead of being materialised:
//   K-major  operand (k contiguous):  3-D tensor map (k, mn, batch), one box (32, rows, 1) per k-block,
//                                     canonical SWIZZLE_128B K-major tile (rows of 128 B, 8-row groups 1 KB apart)
//   MN-major operand (mn contiguous): 3-D tensor map (mn, k, batch), rows/32 boxes (32, 32, 1) per k-block,
//                                     swizzle 128B_ATOM_32B (the only MN-major layout tf32 operands may use):
//                                     32-element mn blocks LBO = 4 KB apart, 4-k groups SBO = 512 B apart
//                                     (UMMA layout SWIZZLE_128B_BASE32B, instruction-descriptor a_major/b_major = 1)
// TMA zero-fills out-of-range rows / k, so ragged M, N, K (and per-batch K ranges) need no masking.
//
// One CTA computes a 128 x BN tile of one batch item and one K slice:
//   warp 0      TMA producer            warp 1   TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  hi/lo split of every landed stage in shared memory (element-wise, hence layout-agnostic),
//               then the epilogue: tcgen05.ld -> (+bias, ReLU) -> coalesced store / red.add
// Pipeline per stage: full (TMA bytes landed) -> split (hi/lo ready) -> tcgen05.commit -> empty.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace pdb {

constexpr int G_BM = 128;
constexpr int G_BK = 32;              // fp32 per 128-byte swizzled row
constexpr int G_THREADS = 192;

struct GemmParams {
    float* C;
    const float* bias;
    int M, N, K;
    int64_t ldc, sc;
    int batch, ksplit, kchunk;        // kchunk: multiple of G_BK
    int c_trans, relu, atomic;
};

template <int BN>
struct GemmSmem {
    static constexpr int STAGES = BN <= 128 ? 3 : 2;
    static constexpr int A_BYTES = G_BM * G_BK * 4;       // 16 KB
    static constexpr int B_BYTES = BN * G_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN;              // main (hi*hi) + correction (lo*hi + hi*lo) accumulators
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// MN-major SWIZZLE_128B_BASE32B operand: 32-element mn blocks `lbo` bytes apart, 4-k groups `sbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(const void* smem_tile, uint32_t byte_offset, uint32_t lbo, uint32_t sbo) {
    uint32_t addr = tc::smem_u32(smem_tile) + byte_offset;
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                         // layout type SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ void split_f4(float4* hi_ptr, float4* lo_ptr) {
    float4 x = *hi_ptr;
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
    *hi_ptr = h;
    *lo_ptr = l;
}

__device__ int g_desc_mode = 0;     // debug hook (pdb_debug_set_desc_mode): 1 swaps LBO / SBO of MN-major descriptors

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const GemmParams p) {
    using S = GemmSmem<BN>;
    constexpr int STAGES = S::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* split = bars + STAGES;
    uint64_t* empty = bars + 2 * STAGES;
    uint64_t* accum = bars + 3 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * G_BM;
    const int n0 = blockIdx.y * BN;
    const int b = blockIdx.z / p.ksplit;
    const int ks = blockIdx.z - b * p.ksplit;
    const int k_begin = ks * p.kchunk;
    const int k_end = min(p.K, k_begin + p.kchunk);
    const int num_kb = (k_end - k_begin + G_BK - 1) / G_BK;
    // MMA N actually needed by this tile (multiple of 16): ragged N costs no tensor time
    const int bn_eff = min(BN, ((p.N - n0 + 15) >> 4) << 4);

    auto a_hi = [&](int s) { return smem + s * S::STAGE_BYTES; };
    auto a_lo = [&](int s) { return smem + s * S::STAGE_BYTES + S::A_BYTES; };
    auto b_hi = [&](int s) { return smem + s * S::STAGE_BYTES + 2 * S::A_BYTES; };
    auto b_lo = [&](int s) { return smem + s * S::STAGE_BYTES + 2 * S::A_BYTES + S::B_BYTES; };

    if (threadIdx.x == 0) {
        tc::prefetch_tensormap(&tm_a);
        tc::prefetch_tensormap(&tm_b);
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&split[s], 128);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(accum, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<S::TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                const int k0 = k_begin + kb * G_BK;
                tc::mbar_wait(&empty[s], ph ^ 1);
                tc::mbar_expect_tx(&full[s], S::A_BYTES + S::B_BYTES);
                if (A_MN) {
#pragma unroll
                    for (int i = 0; i < G_BM / 32; ++i) tma_load_3d(a_hi(s) + i * 4096, &tm_a, &full[s], m0 + 32 * i, k0, b);
                } else {
                    tma_load_3d(a_hi(s), &tm_a, &full[s], k0, m0, b);
                }
                if (B_MN) {
#pragma unroll
                    for (int i = 0; i < BN / 32; ++i) tma_load_3d(b_hi(s) + i * 4096, &tm_b, &full[s], n0 + 32 * i, k0, b);
                } else {
                    tma_load_3d(b_hi(s), &tm_b, &full[s], k0, n0, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_tf32(G_BM, bn_eff) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
            const uint32_t lbo = g_desc_mode ? 512u : 4096u, sbo = g_desc_mode ? 4096u : 512u;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::mbar_wait(&split[s], ph);
                tc::tc_fence_after();
#pragma unroll
                for (int k = 0; k < G_BK / 8; ++k) {
                    uint64_t dah, dal, dbh, dbl;
                    if (A_MN) {
                        dah = umma_desc_mn_sw128(a_hi(s), k * 1024, lbo, sbo);
                        dal = umma_desc_mn_sw128(a_lo(s), k * 1024, lbo, sbo);
                    } else {
                        dah = tc::umma_desc_k_sw128(a_hi(s), k * 32);
                        dal = tc::umma_desc_k_sw128(a_lo(s), k * 32);
                    }
                    if (B_MN) {
                        dbh = umma_desc_mn_sw128(b_hi(s), k * 1024, lbo, sbo);
                        dbl = umma_desc_mn_sw128(b_lo(s), k * 1024, lbo, sbo);
                    } else {
                        dbh = tc::umma_desc_k_sw128(b_hi(s), k * 32);
                        dbl = tc::umma_desc_k_sw128(b_lo(s), k * 32);
                    }
                    tc::mma_tf32(tmem_d + BN, dal, dbh, idesc, (kb | k) != 0);    // correction accumulator
                    tc::mma_tf32(tmem_d + BN, dah, dbl, idesc, 1);
                    tc::mma_tf32(tmem_d, dah, dbh, idesc, (kb | k) != 0);          // main accumulator
                }
                tc::tc_commit(&empty[s]);          // frees the stage once these MMAs have read it
            }
            tc::tc_commit(accum);
        }
    } else {
        // ---- hi/lo split of every landed stage (warps 2..5 = 128 threads)
        const int t = threadIdx.x - 64;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            tc::mbar_wait(&full[s], ph);
            float4* ah = reinterpret_cast<float4*>(a_hi(s));
            float4* al = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
            for (int i = 0; i < S::A_BYTES / 16 / 128; ++i) split_f4(ah + i * 128 + t, al + i * 128 + t);
            float4* bh = reinterpret_cast<float4*>(b_hi(s));
            float4* bl = reinterpret_cast<float4*>(b_lo(s));
#pragma unroll
            for (int i = 0; i < S::B_BYTES / 16 / 128; ++i) split_f4(bh + i * 128 + t, bl + i * 128 + t);
            tc::fence_proxy_async();
            tc::mbar_arrive(&split[s]);
        }
        // ---- epilogue
        tc::mbar_wait(accum, 0);
        tc::tc_fence_after();
        const int quarter = warp & 3;                          // TMEM lanes this warp may access
        const uint32_t tbase = tmem_d + ((uint32_t)(quarter * 32) << 16);
        float* Cb = p.C + (int64_t)b * p.sc;
        if (p.c_trans) {
            // C[n][m]: lane = m, so for a fixed n the warp stores 32 consecutive floats
            const int m = m0 + quarter * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < bn_eff; c0 += 16) {
                float v[16], w[16];
                tc::tmem_ld16(tbase + c0, v);
                tc::tmem_ld16(tbase + BN + c0, w);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += w[i];
                if (m < p.M) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int n = n0 + c0 + i;
                        if (n < p.N) {
                            float x = v[i];
                            if (p.bias) x += __ldg(p.bias + n);
                            if (p.relu) x = fmaxf(x, 0.f);
                            float* dst = Cb + (int64_t)n * p.ldc + m;
                            if (p.atomic) atomicAdd(dst, x); else *dst = x;
                        }
                    }
                }
            }
        } else {
            // C[m][n]: transpose 32x32 blocks through shared memory so that each store instruction
            // writes 32 consecutive floats of one row
            float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
            const int mrow0 = m0 + quarter * 32;
#pragma unroll 1
            for (int c0 = 0; c0 < bn_eff; c0 += 32) {
                float v[32], w[32];
                tc::tmem_ld16(tbase + c0, v);
                tc::tmem_ld16(tbase + c0 + 16, v + 16);      // columns beyond bn_eff hold stale TMEM: never stored
                tc::tmem_ld16(tbase + BN + c0, w);
                tc::tmem_ld16(tbase + BN + c0 + 16, w + 16);
#pragma unroll
                for (int i = 0; i < 32; ++i) tile[lane * 33 + i] = v[i] + w[i];
                __syncwarp();
                const int n = n0 + c0 + lane;
                const bool n_ok = n < p.N;
                const float bv = (p.bias && n_ok) ? __ldg(p.bias + n) : 0.f;
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int m = mrow0 + r;
                    if (m < p.M && n_ok) {
                        float x = tile[r * 33 + lane] + bv;
                        if (p.relu) x = fmaxf(x, 0.f);
                        float* dst = Cb + (int64_t)m * p.ldc + n;
                        if (p.atomic) atomicAdd(dst, x); else *dst = x;
                    }
                }
                __syncwarp();
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<S::TMEM_COLS>(tmem_d);
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tensor_map_3d_f32(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                                  uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1,
                                  CUtensorMapSwizzle swizzle) {
    static EncodeTiledFn3 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(PDB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
        encode = reinterpret_cast<EncodeTiledFn3>(fn);
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(PDB_ERR_INVALID, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu, strides %llu / %llu", (int)r,
                    (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                    (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes);
    return PDB_OK;
}

// operand with `rows` = M or N: K-major -> dims (K, rows, batch), MN-major -> dims (rows, K, batch)
static int make_operand_map(CUtensorMap* map, const float* base, bool mn_major, int rows, int K, int batch, int64_t ld,
                            int64_t bstride, int box_rows) {
    uint64_t outer = mn_major ? (uint64_t)K : (uint64_t)rows;
    uint64_t inner = mn_major ? (uint64_t)rows : (uint64_t)K;
    uint64_t s1 = (uint64_t)ld * 4;
    uint64_t s2 = batch > 1 ? (uint64_t)bstride * 4 : s1 * outer;
    return make_tensor_map_3d_f32(map, base, inner, outer, (uint64_t)batch, s1, s2, 32, mn_major ? (uint32_t)G_BK : (uint32_t)box_rows,
                                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "gemm_tf32x3: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    dim3 grid((unsigned)((p.M + G_BM - 1) / G_BM), (unsigned)((p.N + BN - 1) / BN), (unsigned)(p.batch * p.ksplit));
    gemm_tf32x3_kernel<BN, A_MN, B_MN><<<grid, G_THREADS, S::TOTAL, st>>>(ta, tb, p);
    return launched("gemm_tf32x3");
}

template <int BN>
static int dispatch_layout(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, bool a_mn, bool b_mn, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(ta, tb, p, st);
    if (!a_mn && b_mn) return launch_gemm<BN, false, true>(ta, tb, p, st);
    if (a_mn && !b_mn) return launch_gemm<BN, true, false>(ta, tb, p, st);
    return launch_gemm<BN, true, true>(ta, tb, p, st);
}

int gemm_tf32x3(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int batch,
                int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn, int b_mn,
                int c_trans, int relu, int accumulate, int ksplit, cudaStream_t st) {
    PDB_REQUIRE(A && B && C, "gemm_tf32x3: null pointer");
    PDB_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "gemm_tf32x3: non-positive dimension");
    PDB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm_tf32x3: operands must be 16-byte aligned");
    PDB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && sa % 4 == 0 && sb % 4 == 0,
                "gemm_tf32x3: leading dimensions / batch strides must be multiples of 4 floats (TMA 16-byte strides)");
    PDB_REQUIRE(ksplit >= 1 && (ksplit == 1 || accumulate), "gemm_tf32x3: split-K needs accumulate mode");
    const int BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    GemmParams p;
    p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.sc = sc; p.batch = batch;
    int kb_total = (K + G_BK - 1) / G_BK;
    if (ksplit > kb_total) ksplit = kb_total;
    p.ksplit = ksplit;
    p.kchunk = ((kb_total + ksplit - 1) / ksplit) * G_BK;
    p.ksplit = (K + p.kchunk - 1) / p.kchunk;           // no empty slices
    p.c_trans = c_trans; p.relu = relu; p.atomic = accumulate;
    CUtensorMap ta, tb;
    PDB_TRY(make_operand_map(&ta, A, a_mn != 0, M, K, batch, lda, sa, G_BM));
    PDB_TRY(make_operand_map(&tb, B, b_mn != 0, N, K, batch, ldb, sb, BN));
    if (BN == 32) return dispatch_layout<32>(ta, tb, p, a_mn != 0, b_mn != 0, st);
    if (BN == 64) return dispatch_layout<64>(ta, tb, p, a_mn != 0, b_mn != 0, st);
    return dispatch_layout<128>(ta, tb, p, a_mn != 0, b_mn != 0, st);
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_gemm_tf32x3(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int batch,
                               int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn,
                               int b_mn, int c_trans, int relu, int accumulate, int ksplit, void* stream) {
    return gemm_tf32x3(A, B, C, bias, M, N, K, batch, lda, ldb, ldc, sa, sb, sc, a_mn, b_mn, c_trans, relu, accumulate,
                       ksplit, as_stream(stream));
}

extern "C" PDB_API int pdb_debug_set_desc_mode(int mode) {
    cudaError_t e = cudaMemcpyToSymbol(g_desc_mode, &mode, sizeof(int));
    return e == cudaSuccess ? PDB_OK : fail(PDB_ERR_LAUNCH, "set_desc_mode: %s", cudaGetErrorString(e));
}
