// General fp32 GEMM on the 5th-generation tensor cores ("3xTF32"), the dense-contraction engine of the
// hot path:  for each batch b
//     C_b[m][n] (+)= sum_k A_b(m,k) * B_b(n,k)  (+ bias[n]) (ReLU | GELU)
// It serves nn.Linear forward / input-gradient / weight-gradient of the encoder and decoder layers
// (reference: msdeformattn.py:120-135, ops/modules/ms_deform_attn.py:102-130,
// mask2former_transformer_decoder.py:148-208) and the mask-head einsum with its two gradient products
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
// Persistent kernel, one CTA per SM, 18 warps; a CTA walks 128 x BN output tiles (n fastest, so the CTAs that
// run together share their A rows in L2):
//   warp 0      TMA producer: raw fp32 A / B k-blocks into a 3-deep ring
//   warps 2..9  split every landed k-block: the raw tile is the hi operand (tf32 ignores the low 13 bits), lo =
//               x - trunc(x) is written next to it: B_lo right behind B_hi (so one N = 2*BN MMA yields
//               hi*hi and hi*lo side by side in TMEM), A_lo into its own 2-deep ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane): per 8-wide k step
//               [main | corr] += A_hi x [B_hi ; B_lo]^T   and   corr += A_lo x B_hi^T
//   warps 10..17 epilogue (two per TMEM lane quarter): tcgen05.ld (main + corr) -> (+bias, ReLU / GELU) -> coalesced store / red.add,
//               overlapped with the next tile's main loop through a double-buffered TMEM accumulator
// ALO kernels (K-major A, N <= 112 or BN < 128): the A_lo k-blocks go to tensor memory (tcgen05.st) instead of shared memory and
// the correction MMA reads its A operand from there; with a transposed store (the mask einsum) the ring is 5 stages deep.
// The producer hands each tile's coordinates to the MMA thread through shared memory, and one tcgen05.commit per k-block
// releases both the raw stage (to the producer) and the A_lo stage (to the split warps).  gemm_impl picks narrower column tiles
// (BN = 64 / 32) for problems with few row tiles.  Measurements and the dead ends are in DESIGN.md, section 3.2.
#include <algorithm>
#include "common.cuh"
#include "tc_gemm.cuh"

namespace pdb {

constexpr int G_BM = 128;
constexpr int G_BK = 32;              // fp32 per 128-byte swizzled row
constexpr int G_SPLIT_THREADS = 256;  // warps 2..9
constexpr int G_EPI_WARP0 = 2 + G_SPLIT_THREADS / 32;
constexpr int G_EPI_WARPS = 8;          // two warps per TMEM lane quarter, each takes every other column chunk
constexpr int G_THREADS = (G_EPI_WARP0 + G_EPI_WARPS) * 32;     // 576
constexpr int G_RS = 3;               // raw ring depth (general case)
constexpr int G_RS_MAX = 6;           // barrier slots; the deep ring of the einsum-shaped problems uses 5 (see launch_gemm_alo)
constexpr int G_LS = 2;               // A_lo ring depth

struct GemmParams {
    float* C;
    const float* bias;
    int M, N, K;
    int64_t ldc, sc;
    int batch, ksplit, kchunk;        // kchunk: multiple of G_BK
    int c_trans, relu, atomic;
    int mt, nt, total_tiles;
    int taps, kb_per_tap;             // A rows shifted per K segment (3x3 convolution as one GEMM): see gemm_tf32x3_taps
    int tap_off[9];
    int presplit;                     // B_lo comes pre-computed from global memory (tm_blo) instead of being split here
    int b_box_rows;                   // rows of the K-major B box (BN, or the 16-multiple covering N when N < BN)
    long long* trace;                 // debug timeline of CTA 0 (pdb_debug_set_trace), normally NULL
    int stages, stage_bytes;          // raw ring: depth and bytes per stage (A raw/hi | B raw/hi | B lo)
    int crop_wp, crop_w;              // CROP kernels (3x3 convolution on the padded-width pixel grid): row m = y * crop_wp + x is stored
                                      // at row y * crop_w + x when x < crop_w and dropped otherwise (the two garbage columns)
    const float* gate;                // ReLU backward fused into the store: C[m][n] = gate[m][n] > 0 ? value : 0 (same layout as C;
                                      // row-major stores only) — the input gradient of the layer behind a ReLU
};

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = G_BM * G_BK * 4;       // 16 KB
    static constexpr int B_BYTES = BN * G_BK * 4;
    static constexpr int RAW_STAGE = A_BYTES + 2 * B_BYTES;            // A raw/hi | B raw/hi | B lo
    static constexpr int RAW_TOTAL = G_RS * RAW_STAGE;
    static constexpr int ALO_TOTAL = G_LS * A_BYTES;
    static constexpr int STAGING = G_EPI_WARPS * 32 * 36 * 4;                    // epilogue transpose tiles (row stride 36 floats)
    // deep ring (ALO kernels with a transposed store: no A_lo ring, no staging tiles): 5 stages of A + 2 x 112 B rows
    static constexpr int DEEP_STAGE = A_BYTES + 2 * 112 * G_BK * 4;
    static constexpr int DEEP_STAGES = 5;
    static constexpr int BAR_OFF = (BN == 128 && DEEP_STAGES * DEEP_STAGE > RAW_TOTAL + ALO_TOTAL + STAGING)
                                       ? DEEP_STAGES * DEEP_STAGE : RAW_TOTAL + ALO_TOTAL + STAGING;
    static constexpr int TOTAL = BAR_OFF + 1024 /*align slack*/ + 512 /*barriers, tile-coordinate ring*/;
    static constexpr int ACC_COLS = 2 * BN;               // main (hi*hi) | correction (lo*hi + hi*lo)
    // double-buffered accumulators + (ALO kernels) a 2-deep ring of 32-column A_lo k-blocks.  BN = 128 has all 512 columns
    // taken by the accumulators unless every tile needs at most 112 columns (N <= 112, e.g. the 100-query mask einsum):
    // then the two accumulators are packed 224 columns apart and A_lo takes columns 448..511.
    static constexpr int TMEM_COLS = BN == 32 ? 256 : 512;
    static constexpr int ACC_STRIDE_PACKED = BN == 128 ? 224 : 2 * BN;
    static constexpr int ALO_COL = BN == 128 ? 448 : 4 * BN;
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

template <bool MN>
__device__ __forceinline__ uint64_t operand_desc(const void* tile, int k8) {
    return MN ? umma_desc_mn_sw128(tile, k8 * 1024, 4096u, 512u) : tc::umma_desc_k_sw128(tile, k8 * 32);
}

// epilogue activation: 1 = ReLU, 2 = GELU (erf form, the association of ATen's GeluCUDAKernelImpl: x * 0.5 * (1 + erf(x / sqrt 2)))
__device__ __forceinline__ float epi_act(float x, int mode) {
    if (mode == 1) return fmaxf(x, 0.f);
    if (mode == 2) return x * 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    return x;
}

__device__ __forceinline__ void split4(const float4 x, float4& h, float4& l) {
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
}

// Debug timeline (tools/sweep_gemm.py --trace): compiled in only with -DPDB_GEMM_TRACE (PDB_GEMM_TRACE=1 python -m
// partdistillation_b200.build --force).  Even the `p.trace != NULL` test costs: the single thread that issues the MMAs runs
// latency-bound scalar code between two tcgen05.mma, and every instruction there shows up in the k-block cadence.
__device__ __forceinline__ void trace_evt(const GemmParams& p, int role, int idx, int slot) {
#ifdef PDB_GEMM_TRACE
    if (p.trace && blockIdx.x == 0 && idx < 256) p.trace[(role * 256 + idx) * 4 + slot] = clock64();
#endif
}

// Writes lo = x - trunc_tf32(x) for NV float4 per thread of a tile.  The raw tile itself serves as the hi operand:
// kind::tf32 reads the top 19 bits of each 32-bit container and ignores the low 13 mantissa bits, i.e. the
// tensor core sees exactly trunc_tf32(x) (verified on B200: bit-identical results with and without masking).
template <int NV>
__device__ __forceinline__ void split_tile(const float4* raw, float4* lo, int t) {
    float4 x[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) x[i] = raw[i * G_SPLIT_THREADS + t];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float4 h, l;
        split4(x[i], h, l);
        lo[i * G_SPLIT_THREADS + t] = l;
    }
}

struct TileCoord {
    int m0, n0, b, k_begin, num_kb, bn_eff;
};

template <int BN>
__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int tile) {
    TileCoord c;
    const int ni = tile % p.nt;
    int r = tile / p.nt;
    const int mi = r % p.mt;
    r /= p.mt;
    const int ks = r % p.ksplit;
    c.b = r / p.ksplit;
    c.m0 = mi * G_BM;
    c.n0 = ni * BN;
    c.k_begin = ks * p.kchunk;
    const int k_end = min(p.K, c.k_begin + p.kchunk);
    c.num_kb = (k_end - c.k_begin + G_BK - 1) / G_BK;
    c.bn_eff = min(BN, ((p.N - c.n0 + 15) >> 4) << 4);     // MMA N actually needed: ragged N costs no tensor time
    return c;
}

// Where B_lo of a tile lives relative to B_hi, how many bytes TMA delivers for B_hi, and whether B_lo is directly
// behind the bn_eff valid rows / blocks (so that one N = 2*bn_eff MMA covers [B_hi ; B_lo]).
template <int BN, bool B_MN>
__device__ __forceinline__ void b_layout(const GemmParams& p, int bn_eff, int& blo_off, int& hi_bytes, bool& stacked) {
    const int adjacent = B_MN ? ((bn_eff + 31) / 32) * 4096 : bn_eff * 128;
    if (!p.presplit) {                  // lo written by the split warps after the (full-box) TMA has landed
        blo_off = adjacent;
        hi_bytes = BN * G_BK * 4;
    } else if (B_MN) {                  // only the needed 32-column blocks are loaded
        blo_off = adjacent;
        hi_bytes = adjacent;
    } else {                            // K-major box rows are fixed by the tensor map
        hi_bytes = p.b_box_rows * 128;
        blo_off = (p.b_box_rows == bn_eff) ? adjacent : BN * G_BK * 4;
    }
    stacked = (blo_off == adjacent) && (!B_MN || bn_eff % 32 == 0);
}

// ALO: the A_lo k-blocks live in tensor memory instead of shared memory: the split warps write them with tcgen05.st (thread =
// one row = one TMEM lane) and the correction MMA takes its A operand from TMEM, which removes 16 KB of shared-memory writes and
// 16 KB of operand reads per k-block.  K-major A only, and the accumulators must leave 64 columns (see GemmSmem).
template <int BN, bool A_MN, bool B_MN, bool ALO, bool GATE = false, bool CROP = false>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const __grid_constant__ CUtensorMap tm_blo, const GemmParams p) {
    using S = GemmSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment by pointer arithmetic (no integer round trip), so the compiler keeps the shared state space
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* alo_ring = smem + S::RAW_TOTAL;
    float* staging = reinterpret_cast<float*>(smem + S::RAW_TOTAL + S::ALO_TOTAL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* raw_full = bars;                      // TMA bytes landed
    uint64_t* raw_empty = bars + G_RS_MAX;          // MMAs reading the stage completed
    uint64_t* split_done = bars + 2 * G_RS_MAX;     // hi / lo of the stage ready (128 arrivals)
    uint64_t* alo_empty = bars + 3 * G_RS_MAX;      // [G_LS]
    uint64_t* acc_full = alo_empty + G_LS;          // [2]
    uint64_t* acc_empty = acc_full + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    // tile coordinates handed from the producer to the MMA thread (which would otherwise spend ~700 clk of dependent integer
    // divisions between two tiles with the tensor pipe drained): entry t & 7 belongs to this CTA's t-th tile; the producer is at
    // most G_RS_MAX k-blocks, hence tiles, ahead of the MMA thread
    TileCoord* coord_ring = reinterpret_cast<TileCoord*>(smem + S::BAR_OFF + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int nstages = p.stages;
    const int stage_bytes = p.stage_bytes;
    // every role walks the raw ring with its own (stage, phase parity) pair
    auto ring_next = [&](int& s, uint32_t& ph) {
        if (++s == nstages) {
            s = 0;
            ph ^= 1u;
        }
    };
    auto a_raw = [&](int s) { return smem + s * stage_bytes; };
    auto b_raw = [&](int s) { return smem + s * stage_bytes + S::A_BYTES; };
    auto a_lo = [&](int s) { return alo_ring + s * S::A_BYTES; };

    if (threadIdx.x == 0) {
        tc::prefetch_tensormap(&tm_a);
        tc::prefetch_tensormap(&tm_b);
        for (int s = 0; s < G_RS_MAX; ++s) {
            tc::mbar_init(&raw_full[s], 1);
            tc::mbar_init(&raw_empty[s], 1);
            tc::mbar_init(&split_done[s], G_SPLIT_THREADS);
        }
        for (int s = 0; s < G_LS; ++s) tc::mbar_init(&alo_empty[s], 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], G_EPI_WARPS * 32);
        }
        tc::fence_barrier_init();
        // SM clock vs wall clock over the kernel (debug trace only).  Keep thread-divergent code like this in front of the
        // __syncthreads below: a divergent branch after it makes the compiler treat the role loops as non-uniform, and every
        // UTCHMMA / UTMALDG then issues through an ELECT + R2UR waterfall (measured: all GEMMs 20 % slower).
#ifdef PDB_GEMM_TRACE
        if (p.trace && blockIdx.x == 0) {
            long long gt;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
            p.trace[(3 * 256 + 200) * 4 + 0] = clock64();
            p.trace[(3 * 256 + 200) * 4 + 1] = gt;
        }
#endif
    }
    if (warp == 1) tc::tmem_alloc<S::TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    static_assert(!(ALO && A_MN), "A_lo in tensor memory needs a K-major A operand");
    constexpr uint32_t acc_stride = ALO ? S::ACC_STRIDE_PACKED : S::ACC_COLS;
    const uint32_t tmem_alo = tmem_base + S::ALO_COL;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0, t = 0;
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
                const TileCoord c = tile_coord<BN>(p, tile);
                coord_ring[t & 7] = c;          // published by the release of this tile's first mbarrier.arrive.expect_tx
                int blo_off, hi_bytes;
                bool stacked;
                b_layout<BN, B_MN>(p, c.bn_eff, blo_off, hi_bytes, stacked);
                const int nblk = p.presplit ? (c.bn_eff + 31) / 32 : BN / 32;      // MN-major B: 32-column blocks to load
                for (int kb = 0; kb < c.num_kb; ++kb, ++it, ring_next(s, ph)) {
                    const int k0 = c.k_begin + kb * G_BK;
                    tc::mbar_wait(&raw_empty[s], ph ^ 1);
                    trace_evt(p, 0, it, 0);
                    tc::mbar_expect_tx(&raw_full[s], S::A_BYTES + hi_bytes * (p.presplit ? 2 : 1));
                    if (A_MN) {
#pragma unroll
                        for (int i = 0; i < G_BM / 32; ++i) tma_load_3d(a_raw(s) + i * 4096, &tm_a, &raw_full[s], c.m0 + 32 * i, k0, c.b);
                    } else if (p.taps > 1) {
                        // K = taps * C: k-block kbg of tap t reads channels of the A rows shifted by tap_off[t]
                        const int kbg = k0 / G_BK;
                        const int t = kbg / p.kb_per_tap;
                        tma_load_3d(a_raw(s), &tm_a, &raw_full[s], (kbg - t * p.kb_per_tap) * G_BK, c.m0 + p.tap_off[t], c.b);
                    } else {
                        tma_load_3d(a_raw(s), &tm_a, &raw_full[s], k0, c.m0, c.b);
                    }
                    if (B_MN) {
                        for (int i = 0; i < nblk; ++i) tma_load_3d(b_raw(s) + i * 4096, &tm_b, &raw_full[s], c.n0 + 32 * i, k0, c.b);
                        if (p.presplit)
                            for (int i = 0; i < nblk; ++i)
                                tma_load_3d(b_raw(s) + blo_off + i * 4096, &tm_blo, &raw_full[s], c.n0 + 32 * i, k0, c.b);
                    } else {
                        tma_load_3d(b_raw(s), &tm_b, &raw_full[s], k0, c.n0, c.b);
                        if (p.presplit) tma_load_3d(b_raw(s) + blo_off, &tm_blo, &raw_full[s], k0, c.n0, c.b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t it = 0, t = 0;
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
                tc::mbar_wait(&raw_full[s], ph);          // first k-block of the tile landed => its coordinates are visible
                const TileCoord c = coord_ring[t & 7];
                const uint32_t ab = t & 1;
                const uint32_t tmem_main = tmem_base + ab * acc_stride;
                const uint32_t tmem_corr = tmem_main + c.bn_eff;
                // one N = 2*bn_eff MMA when B_lo sits directly behind the bn_eff valid rows / blocks of B_hi
                int blo_off, hi_bytes;
                bool stacked;
                b_layout<BN, B_MN>(p, c.bn_eff, blo_off, hi_bytes, stacked);
                const uint32_t major = (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
                const uint32_t idesc1 = tc::umma_idesc_tf32(G_BM, c.bn_eff) | major;
                const uint32_t idesc2 = tc::umma_idesc_tf32(G_BM, 2 * c.bn_eff) | major;
                trace_evt(p, 1, it, 3);
                tc::mbar_wait(&acc_empty[ab], ((t >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                for (int kb = 0; kb < c.num_kb; ++kb, ++it, ring_next(s, ph)) {
                    const int ls = it % G_LS;
                    tc::mbar_wait(&raw_full[s], ph);
                    trace_evt(p, 1, it, 0);
                    tc::mbar_wait(&split_done[s], ph);
                    tc::tc_fence_after();
                    trace_evt(p, 1, it, 1);
                    const uint8_t* blo = b_raw(s) + blo_off;
#pragma unroll
                    for (int k = 0; k < G_BK / 8; ++k) {
                        const uint64_t dah = operand_desc<A_MN>(a_raw(s), k);
                        const uint64_t dal = operand_desc<A_MN>(a_lo(ls), k);
                        const uint64_t dbh = operand_desc<B_MN>(b_raw(s), k);
                        const uint32_t acc = (kb | k) != 0;
                        if (ALO) {
                            const uint32_t tal = tmem_alo + ls * G_BK + k * 8;   // A_lo rows = TMEM lanes, k = columns
                            if (stacked) {
                                tc::mma_tf32(tmem_main, dah, dbh, idesc2, acc);
                                tc::mma_tf32_ta(tmem_corr, tal, dbh, idesc1, 1);
                            } else {
                                const uint64_t dbl = operand_desc<B_MN>(blo, k);
                                tc::mma_tf32_ta(tmem_corr, tal, dbh, idesc1, acc);
                                tc::mma_tf32(tmem_corr, dah, dbl, idesc1, 1);
                                tc::mma_tf32(tmem_main, dah, dbh, idesc1, acc);
                            }
                        } else if (stacked) {
                            tc::mma_tf32(tmem_main, dah, dbh, idesc2, acc);      // [main | corr] += A_hi x [B_hi ; B_lo]^T
                            tc::mma_tf32(tmem_corr, dal, dbh, idesc1, 1);        // corr += A_lo x B_hi^T
                        } else {
                            const uint64_t dbl = operand_desc<B_MN>(blo, k);
                            tc::mma_tf32(tmem_corr, dal, dbh, idesc1, acc);
                            tc::mma_tf32(tmem_corr, dah, dbl, idesc1, 1);
                            tc::mma_tf32(tmem_main, dah, dbh, idesc1, acc);
                        }
                    }
                    tc::tc_commit(&raw_empty[s]);          // arrives once the MMAs above have read their operands (raw stage and A_lo stage)
                    trace_evt(p, 1, it, 2);
                }
                tc::tc_commit(&acc_full[ab]);
            }
        }
    } else if (warp < G_EPI_WARP0) {
        // ------------------------------------------------------------------ hi / lo split (warps 2..9)
        const int tid = threadIdx.x - 64;
        uint32_t it = 0;
        int s = 0, s2 = 0;              // s2 / ph2: stage and parity of k-block it - G_LS
        uint32_t ph = 0, ph2 = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord c = tile_coord<BN>(p, tile);
            int blo_off, hi_bytes;
            bool stacked;
            b_layout<BN, B_MN>(p, c.bn_eff, blo_off, hi_bytes, stacked);
            for (int kb = 0; kb < c.num_kb; ++kb, ++it, ring_next(s, ph)) {
                const int ls = it % G_LS;
                if (tid == 0) trace_evt(p, 2, it, 0);
                tc::mbar_wait(&raw_full[s], ph);
                if (tid == 0) trace_evt(p, 2, it, 1);
                // A_lo stage ls was last read by the MMAs of k-block it - G_LS; their completion is that k-block's commit on its
                // raw_empty barrier (whose next phase cannot complete before this k-block has been split): no second commit per
                // k-block in the issuing thread
                if (it >= G_LS) {
                    tc::mbar_wait(&raw_empty[s2], ph2);
                    ring_next(s2, ph2);
                }
                if (tid == 0) trace_evt(p, 2, it, 2);
                if (ALO) {
                    // thread = one row of the tile (TMEM lane 32 * (warp % 4) + lane), 16 of the 32 k columns: four 16-byte chunks
                    // of the 128-byte-swizzled row (chunk j of row r sits at position j ^ (r & 7))
                    tc::tc_fence_after();
                    const int row = (warp & 3) * 32 + lane;
                    const int kh = (warp - 2) >> 2;
                    const float4* rowp = reinterpret_cast<const float4*>(a_raw(s) + row * 128);
                    uint32_t lo[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 h, l;
                        split4(rowp[(kh * 4 + j) ^ (row & 7)], h, l);
                        lo[4 * j] = __float_as_uint(l.x); lo[4 * j + 1] = __float_as_uint(l.y);
                        lo[4 * j + 2] = __float_as_uint(l.z); lo[4 * j + 3] = __float_as_uint(l.w);
                    }
                    tc::tmem_st16(tmem_alo + ((uint32_t)((warp & 3) * 32) << 16) + ls * G_BK + kh * 16, lo);
                    tc::tmem_st_wait();
                    tc::tc_fence_before();
                } else {
                    split_tile<S::A_BYTES / 16 / G_SPLIT_THREADS>(reinterpret_cast<float4*>(a_raw(s)), reinterpret_cast<float4*>(a_lo(ls)), tid);
                }
                float4* braw = reinterpret_cast<float4*>(b_raw(s));
                float4* blo = reinterpret_cast<float4*>(b_raw(s) + blo_off);
                if (p.presplit) {
                    // B_lo arrived by TMA
                } else if (c.bn_eff == BN) {
                    split_tile<S::B_BYTES / 16 / G_SPLIT_THREADS>(braw, blo, tid);
                } else {
                    // ragged N tile: only the first bn_eff rows (K-major) / 32-column blocks (MN-major) matter; B_lo
                    // starts right behind them, over the unused (zero-filled) tail of the raw tile, so the ranges
                    // read (index < nvec) and written as lo (index >= nvec) are disjoint
                    const int nvec = B_MN ? ((c.bn_eff + 31) / 32) * 256 : c.bn_eff * 8;
                    constexpr int NVB = S::B_BYTES / 16 / G_SPLIT_THREADS;
                    float4 x[NVB];
#pragma unroll
                    for (int i = 0; i < NVB; ++i) x[i] = (i * G_SPLIT_THREADS + tid < nvec) ? braw[i * G_SPLIT_THREADS + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < NVB; ++i) {
                        if (i * G_SPLIT_THREADS + tid < nvec) {
                            float4 h, l;
                            split4(x[i], h, l);
                            blo[i * G_SPLIT_THREADS + tid] = l;
                        }
                    }
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&split_done[s]);
                if (tid == 0) trace_evt(p, 2, it, 3);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (last 4 warps)
        const int quarter = warp & 3;                          // TMEM lanes this warp may access
        const int half = (warp - G_EPI_WARP0) >> 2;            // which of the two warps of this quarter
        float* tile_s = staging + (warp - G_EPI_WARP0) * (32 * 36);
        const bool trace_thread = threadIdx.x == G_EPI_WARP0 * 32;
        // 16-byte row stores need aligned rows
        const bool vec_ok = !p.c_trans && (p.N % 4 == 0) && (p.ldc % 4 == 0) && (p.sc % 4 == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
            const TileCoord c = tile_coord<BN>(p, tile);
            const uint32_t ab = t & 1;
            const uint32_t tbase = tmem_base + ab * acc_stride + ((uint32_t)(quarter * 32) << 16);
            if (trace_thread) trace_evt(p, 3, t, 0);
            tc::mbar_wait(&acc_full[ab], (t >> 1) & 1);
            tc::tc_fence_after();
            if (trace_thread) trace_evt(p, 3, t, 1);
            float* Cb = p.C + (int64_t)c.b * p.sc;
            if (p.c_trans) {
                // C[n][m]: lane = m, so for a fixed n the warp stores 32 consecutive floats
                const int m = c.m0 + quarter * 32 + lane;
                const bool m_ok = m < p.M;
                float* col = Cb + (int64_t)c.n0 * p.ldc + m;
                const int nvalid = min(c.bn_eff, p.N - c.n0);
                col += (int64_t)(half * 16) * p.ldc;
                if (half * 16 >= c.bn_eff) {                   // nothing to read for this warp: release at once
                    tc::tc_fence_before();
                    tc::mbar_arrive(&acc_empty[ab]);
                }
#pragma unroll 1
                for (int c0 = half * 16; c0 < c.bn_eff; c0 += 32) {
                    uint32_t vr[16], wr[16];
                    tc::tmem_ld16_nowait(tbase + c0, vr);
                    tc::tmem_ld16_nowait(tbase + c.bn_eff + c0, wr);
                    tc::tmem_ld_wait();
                    if (c0 + 32 >= c.bn_eff) {                 // last read of this accumulator: hand it back to the MMA warp
                        tc::tc_fence_before();
                        tc::mbar_arrive(&acc_empty[ab]);
                    }
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x = __uint_as_float(vr[i]) + __uint_as_float(wr[i]);
                        if (p.bias) x += __ldg(p.bias + min(c.n0 + c0 + i, p.N - 1));
                        v[i] = epi_act(x, p.relu);
                    }
                    if (m_ok) {
                        const int lim = nvalid - c0;           // columns of this group that exist
                        if (p.atomic) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (i < lim) atomicAdd(col + (int64_t)i * p.ldc, v[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (i < lim) col[(int64_t)i * p.ldc] = v[i];
                        }
                    }
                    col += 32 * p.ldc;
                }
            } else {
                // C[m][n]: transpose 32 x 32 blocks through shared memory so that a warp store instruction writes
                // four 128-byte row segments (float4 per lane)
                const int mrow0 = c.m0 + quarter * 32;
                const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
                if (half * 32 >= c.bn_eff) {
                    tc::tc_fence_before();
                    tc::mbar_arrive(&acc_empty[ab]);
                }
#pragma unroll 1
                for (int c0 = half * 32; c0 < c.bn_eff; c0 += 64) {
                    const bool two = c0 + 16 < c.bn_eff;       // bn_eff is a multiple of 16, not of 32
                    uint32_t v0[16], w0[16], v1[16], w1[16];
                    tc::tmem_ld16_nowait(tbase + c0, v0);
                    tc::tmem_ld16_nowait(tbase + c.bn_eff + c0, w0);
                    if (two) {
                        tc::tmem_ld16_nowait(tbase + c0 + 16, v1);
                        tc::tmem_ld16_nowait(tbase + c.bn_eff + c0 + 16, w1);
                    }
                    tc::tmem_ld_wait();
                    if (c0 + 64 >= c.bn_eff) {
                        tc::tc_fence_before();
                        tc::mbar_arrive(&acc_empty[ab]);
                    }
                    float4* srow = reinterpret_cast<float4*>(tile_s + lane * 36);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        srow[i] = make_float4(__uint_as_float(v0[4 * i]) + __uint_as_float(w0[4 * i]),
                                              __uint_as_float(v0[4 * i + 1]) + __uint_as_float(w0[4 * i + 1]),
                                              __uint_as_float(v0[4 * i + 2]) + __uint_as_float(w0[4 * i + 2]),
                                              __uint_as_float(v0[4 * i + 3]) + __uint_as_float(w0[4 * i + 3]));
                    if (two) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            srow[4 + i] = make_float4(__uint_as_float(v1[4 * i]) + __uint_as_float(w1[4 * i]),
                                                      __uint_as_float(v1[4 * i + 1]) + __uint_as_float(w1[4 * i + 1]),
                                                      __uint_as_float(v1[4 * i + 2]) + __uint_as_float(w1[4 * i + 2]),
                                                      __uint_as_float(v1[4 * i + 3]) + __uint_as_float(w1[4 * i + 3]));
                    }
                    __syncwarp();
                    const int ncols = min(two ? 32 : 16, p.N - c.n0 - c0);       // valid columns of this block
                    if (vec_ok) {
                        const int n = c.n0 + c0 + sub_c;
                        const bool n_ok = sub_c < ncols;                            // N % 4 == 0: whole float4 valid
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias && n_ok) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                        float* dst = Cb + (int64_t)(mrow0 + sub_r) * p.ldc + n;
                        // ReLU gate: all eight loads of this thread in flight at once (one at a time inside the store loop cost
                        // eight L2 round trips per 32-column block: +140 us on the encoder's 43 008 x 1024 input gradient)
                        float4 hg[GATE ? 8 : 1];
                        if (GATE) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                hg[i] = (n_ok && mrow0 + i * 4 + sub_r < p.M)
                                            ? __ldg(reinterpret_cast<const float4*>(p.gate + (dst - p.C) + (int64_t)i * 4 * p.ldc))
                                            : make_float4(1.f, 1.f, 1.f, 1.f);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + sub_r;
                            float4 x = *reinterpret_cast<const float4*>(tile_s + r * 36 + sub_c);
                            x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
                            if (p.relu) { x.x = epi_act(x.x, p.relu); x.y = epi_act(x.y, p.relu); x.z = epi_act(x.z, p.relu); x.w = epi_act(x.w, p.relu); }
                            if (CROP) {
                                const int m = mrow0 + r, y = m / p.crop_wp, xc = m - y * p.crop_wp;
                                if (n_ok && m < p.M && xc < p.crop_w)
                                    *reinterpret_cast<float4*>(Cb + ((int64_t)y * p.crop_w + xc) * p.ldc + n) = x;
                            } else if (n_ok && mrow0 + r < p.M) {
                                if (GATE) {
                                    const float4 h = hg[GATE ? i : 0];
                                    x.x = h.x > 0.f ? x.x : 0.f; x.y = h.y > 0.f ? x.y : 0.f;
                                    x.z = h.z > 0.f ? x.z : 0.f; x.w = h.w > 0.f ? x.w : 0.f;
                                }
                                if (p.atomic) red_add_v4(dst, x.x, x.y, x.z, x.w);
                                else *reinterpret_cast<float4*>(dst) = x;
                            }
                            dst += 4 * p.ldc;
                        }
                    } else {
                        const int n = c.n0 + c0 + lane;
                        const bool n_ok = lane < ncols;
                        const float bv = (p.bias && n_ok) ? __ldg(p.bias + n) : 0.f;
                        float* dst = Cb + (int64_t)mrow0 * p.ldc + n;
#pragma unroll 4
                        for (int r = 0; r < 32; ++r) {
                            if (CROP) {
                                const int m = mrow0 + r, y = m / p.crop_wp, xc = m - y * p.crop_wp;
                                if (m < p.M && n_ok && xc < p.crop_w)
                                    Cb[((int64_t)y * p.crop_w + xc) * p.ldc + n] = epi_act(tile_s[r * 36 + lane] + bv, p.relu);
                            } else if (mrow0 + r < p.M && n_ok) {
                                float x = tile_s[r * 36 + lane] + bv;
                                x = epi_act(x, p.relu);
                                if (GATE && !(__ldg(p.gate + (dst - p.C)) > 0.f)) x = 0.f;
                                if (p.atomic) atomicAdd(dst, x); else *dst = x;
                            }
                            dst += p.ldc;
                        }
                    }
                    __syncwarp();
                }
            }
            if (trace_thread) trace_evt(p, 3, t, 2);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
#ifdef PDB_GEMM_TRACE
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
        long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        p.trace[(3 * 256 + 200) * 4 + 2] = clock64();
        p.trace[(3 * 256 + 200) * 4 + 3] = gt;
    }
#endif
    if (warp == 1) tc::tmem_dealloc<S::TMEM_COLS>(tmem_base);
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

static long long* g_trace_ptr = nullptr;
static int g_narrow_tiles = 1;                // pdb_debug_set_gemm_alo_tmem(3) switches the narrow-tile heuristic off
static int g_deep_ring = 1;                   // pdb_debug_set_gemm_alo_tmem(2) keeps A_lo in TMEM but the 3-stage ring
static int g_alo_tmem = 1;                    // pdb_debug_set_gemm_alo_tmem(0) forces A_lo through shared memory

template <int BN, bool A_MN, bool B_MN, bool ALO>
static int launch_gemm_alo(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbl, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, A_MN, B_MN, ALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "gemm_tf32x3: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    GemmParams q = p;
    q.trace = g_trace_ptr;
    // Ring depth.  The general kernel fits 3 stages next to the A_lo ring and the epilogue staging tiles.  With A_lo in tensor
    // memory and a transposed store (the mask einsum) neither is used, and a stage shrinks to A + 2 x 112 B rows: 5 stages, which
    // is what it takes to keep the HBM stream of A in flight (TMA latency under load is 2 000 - 3 000 clk, a k-block ~700 clk).
    const bool deep = ALO && !B_MN && BN == 128 && p.c_trans && g_deep_ring;      // (MN-major B blocks are 4 KB apart: 32 KB per stage)
    q.stages = deep ? S::DEEP_STAGES : G_RS;
    q.stage_bytes = deep ? S::DEEP_STAGE : S::RAW_STAGE;
    q.mt = (p.M + G_BM - 1) / G_BM;
    q.nt = (p.N + BN - 1) / BN;
    q.total_tiles = q.mt * q.nt * p.batch * p.ksplit;
    const unsigned grid = (unsigned)std::min(q.total_tiles, kNumSMs);        // persistent: one CTA per SM
    gemm_tf32x3_kernel<BN, A_MN, B_MN, ALO><<<grid, G_THREADS, S::TOTAL, st>>>(ta, tb, tbl, q);
    return launched("gemm_tf32x3");
}

// The gated store (ReLU backward fused into the input-gradient GEMM) exists for the one shape that uses it: 128-wide tiles, K-major
// A (dy), MN-major B (W read in place).  Its own instantiation keeps the eight gate loads per thread out of every other GEMM.
static int launch_gemm_gated(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbl, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<128>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<128, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "gemm_tf32x3: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    GemmParams q = p;
    q.trace = g_trace_ptr;
    q.stages = G_RS;
    q.stage_bytes = S::RAW_STAGE;
    q.mt = (p.M + G_BM - 1) / G_BM;
    q.nt = (p.N + 127) / 128;
    q.total_tiles = q.mt * q.nt * p.batch * p.ksplit;
    const unsigned grid = (unsigned)std::min(q.total_tiles, kNumSMs);
    gemm_tf32x3_kernel<128, false, true, false, true><<<grid, G_THREADS, S::TOTAL, st>>>(ta, tb, tbl, q);
    return launched("gemm_tf32x3(gated)");
}

// The cropping store of the 3x3 convolutions (K-major A and B, 128-wide tiles), its own instantiation like the gated one.
static int launch_gemm_crop(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbl, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<128>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<128, false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "gemm_tf32x3: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    GemmParams q = p;
    q.trace = g_trace_ptr;
    q.stages = G_RS;
    q.stage_bytes = S::RAW_STAGE;
    q.mt = (p.M + G_BM - 1) / G_BM;
    q.nt = (p.N + 127) / 128;
    q.total_tiles = q.mt * q.nt * p.batch * p.ksplit;
    const unsigned grid = (unsigned)std::min(q.total_tiles, kNumSMs);
    gemm_tf32x3_kernel<128, false, false, false, false, true><<<grid, G_THREADS, S::TOTAL, st>>>(ta, tb, tbl, q);
    return launched("gemm_tf32x3(crop)");
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbl, const GemmParams& p, cudaStream_t st) {
    // A_lo in tensor memory: K-major A only (the split warps own whole rows), and the accumulators must leave 64 columns
    if constexpr (!A_MN) {
        if (g_alo_tmem && (BN < 128 || p.N <= 112)) return launch_gemm_alo<BN, A_MN, B_MN, true>(ta, tb, tbl, p, st);
    }
    return launch_gemm_alo<BN, A_MN, B_MN, false>(ta, tb, tbl, p, st);
}

template <int BN>
static int dispatch_layout(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbl, const GemmParams& p, bool a_mn, bool b_mn, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(ta, tb, tbl, p, st);
    if (!a_mn && b_mn) return launch_gemm<BN, false, true>(ta, tb, tbl, p, st);
    if (a_mn && !b_mn) return launch_gemm<BN, true, false>(ta, tb, tbl, p, st);
    return launch_gemm<BN, true, true>(ta, tb, tbl, p, st);
}

static int gemm_impl(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M, int N, int K, int batch,
                     int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn, int b_mn,
                     int c_trans, int relu, int accumulate, int ksplit, int taps, const int32_t* tap_off, int a_rows,
                     cudaStream_t st, const float* gate = nullptr, int crop_wp = 0, int crop_w = 0) {
    PDB_REQUIRE(A && B && C, "gemm_tf32x3: null pointer");
    PDB_REQUIRE(!gate || (!c_trans && ksplit == 1 && !accumulate && (reinterpret_cast<uintptr_t>(gate) & 15) == 0),
                "gemm_tf32x3: the ReLU gate needs a plain row-major store (no transpose, split-K or accumulation), 16-byte aligned");
    PDB_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "gemm_tf32x3: non-positive dimension");
    PDB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm_tf32x3: operands must be 16-byte aligned");
    PDB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && sa % 4 == 0 && sb % 4 == 0,
                "gemm_tf32x3: leading dimensions / batch strides must be multiples of 4 floats (TMA 16-byte strides)");
    PDB_REQUIRE(ksplit >= 1 && (ksplit == 1 || accumulate), "gemm_tf32x3: split-K needs accumulate mode");
    GemmParams p;
    p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.sc = sc; p.batch = batch;
    p.gate = gate;
    p.crop_wp = crop_wp; p.crop_w = crop_w;
    int kb_total = (K + G_BK - 1) / G_BK;
    if (ksplit > kb_total) ksplit = kb_total;
    p.ksplit = ksplit;
    p.kchunk = ((kb_total + ksplit - 1) / ksplit) * G_BK;
    p.ksplit = (K + p.kchunk - 1) / p.kchunk;           // no empty slices
    // Column-tile width.  Problems with few row tiles (the decoder's 200-row linears: 2 row tiles, and K = 2048 in its FFN)
    // would occupy a handful of SMs with 128-wide tiles; narrower tiles put up to 4x more SMs to work.  Deterministic, unlike
    // split-K (which these forward / input-gradient products could not use anyway: bias, fixed summation order).
    int BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    if (BN == 128 && taps <= 1 && g_narrow_tiles) {
        const int64_t row_tiles = (int64_t)((M + G_BM - 1) / G_BM) * batch * p.ksplit;
        if (row_tiles * ((N + 127) / 128) * 2 <= kNumSMs) BN = (row_tiles * ((N + 63) / 64) * 2 <= kNumSMs) ? 32 : 64;
    }
    p.c_trans = c_trans; p.relu = relu; p.atomic = accumulate;
    p.taps = taps;
    p.kb_per_tap = taps > 1 ? (K / taps) / G_BK : 0;
    for (int t = 0; t < 9; ++t) p.tap_off[t] = (taps > 1 && t < taps) ? tap_off[t] : 0;
    CUtensorMap ta, tb, tbl;
    // with taps the A operand is (a_rows x K / taps) per batch item, read at shifted rows
    if (taps > 1) PDB_TRY(make_operand_map(&ta, A, false, a_rows, K / taps, batch, lda, sa, G_BM));
    else PDB_TRY(make_operand_map(&ta, A, a_mn != 0, M, K, batch, lda, sa, G_BM));
    p.presplit = B_lo != nullptr;
    // pre-split K-major B with a single n tile: the box covers exactly the rows the MMA needs, B_lo lands right behind
    p.b_box_rows = (p.presplit && N < BN) ? ((N + 15) / 16) * 16 : BN;
    PDB_REQUIRE(!B_lo || (reinterpret_cast<uintptr_t>(B_lo) & 15) == 0, "gemm_tf32x3: B_lo must be 16-byte aligned");
    PDB_TRY(make_operand_map(&tb, B, b_mn != 0, N, K, batch, ldb, sb, p.b_box_rows));
    tbl = tb;
    if (B_lo) PDB_TRY(make_operand_map(&tbl, B_lo, b_mn != 0, N, K, batch, ldb, sb, p.b_box_rows));
    if (crop_wp > 0) {
        PDB_REQUIRE(BN == 128 && !a_mn && !b_mn && N > 112 && !c_trans && !accumulate && ksplit == 1 && crop_w > 0 &&
                        crop_w <= crop_wp && M % crop_wp == 0,
                    "gemm_tf32x3: the cropping store needs 128-wide tiles (N > 112), K-major operands, a plain store and M a "
                    "multiple of the padded width");
        return launch_gemm_crop(ta, tb, tbl, p, st);
    }
    if (gate) {
        PDB_REQUIRE(BN == 128 && !a_mn && b_mn && N > 112 && N % 4 == 0,
                    "gemm_tf32x3: the gated store needs 128-wide tiles (many rows, N > 112, N %% 4 == 0), K-major A and MN-major B");
        return launch_gemm_gated(ta, tb, tbl, p, st);
    }
    if (BN == 32) return dispatch_layout<32>(ta, tb, tbl, p, a_mn != 0, b_mn != 0, st);
    if (BN == 64) return dispatch_layout<64>(ta, tb, tbl, p, a_mn != 0, b_mn != 0, st);
    return dispatch_layout<128>(ta, tb, tbl, p, a_mn != 0, b_mn != 0, st);
}

int gemm_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M, int N, int K, int batch,
                int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn, int b_mn,
                int c_trans, int relu, int accumulate, int ksplit, cudaStream_t st) {
    return gemm_impl(A, B, B_lo, C, bias, M, N, K, batch, lda, ldb, ldc, sa, sb, sc, a_mn, b_mn, c_trans, relu, accumulate,
                     ksplit, 1, nullptr, 0, st);
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_gemm_taps_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M,
                                    int N, int Ck, int batch, int a_rows, int64_t lda, int64_t ldc, int64_t sa, int64_t sc,
                                    int taps, const int32_t* tap_off, int relu, void* stream) {
    PDB_REQUIRE(taps >= 1 && taps <= 9 && tap_off, "gemm_taps: 1 <= taps <= 9");
    PDB_REQUIRE(Ck > 0 && Ck % G_BK == 0, "gemm_taps: channels per tap (%d) must be a multiple of %d", Ck, G_BK);
    for (int t = 0; t < taps; ++t) PDB_REQUIRE(tap_off[t] >= 0, "gemm_taps: negative row offset");
    return gemm_impl(A, B, B_lo, C, bias, M, N, taps * Ck, batch, lda, (int64_t)taps * Ck, ldc, sa, 0, sc, 0, 0, 0, relu, 0, 1,
                     taps, tap_off, a_rows, as_stream(stream));
}

// pdb_gemm_taps_tf32x3 whose store drops the garbage columns of the padded-width pixel grid: row m = y * wp + x of the product is
// written to row y * w + x of C (x < w) — C is the (H, W, N) map itself (batch stride sc = H * w * N), no crop pass afterwards.
extern "C" int pdb_gemm_taps_cropped_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M,
                                            int N, int Ck, int batch, int a_rows, int64_t lda, int64_t ldc, int64_t sa, int64_t sc,
                                            int taps, const int32_t* tap_off, int relu, int wp, int w, void* stream) {
    PDB_REQUIRE(taps >= 1 && taps <= 9 && tap_off, "gemm_taps: 1 <= taps <= 9");
    PDB_REQUIRE(Ck > 0 && Ck % G_BK == 0, "gemm_taps: channels per tap (%d) must be a multiple of %d", Ck, G_BK);
    PDB_REQUIRE(wp > 0 && w > 0, "gemm_taps_cropped: widths");
    for (int t = 0; t < taps; ++t) PDB_REQUIRE(tap_off[t] >= 0, "gemm_taps: negative row offset");
    return gemm_impl(A, B, B_lo, C, bias, M, N, taps * Ck, batch, lda, (int64_t)taps * Ck, ldc, sa, 0, sc, 0, 0, 0, relu, 0, 1,
                     taps, tap_off, a_rows, as_stream(stream), nullptr, wp, w);
}

extern "C" int pdb_gemm_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M, int N, int K, int batch,
                               int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn,
                               int b_mn, int c_trans, int relu, int accumulate, int ksplit, void* stream) {
    return gemm_tf32x3(A, B, B_lo, C, bias, M, N, K, batch, lda, ldb, ldc, sa, sb, sc, a_mn, b_mn, c_trans, relu, accumulate,
                       ksplit, as_stream(stream));
}

// pdb_gemm_tf32x3 with the backward of a ReLU fused into the store: C[m][n] = gate[m][n] > 0 ? (A B^T)[m][n] : 0, gate laid out
// like C.  The input gradient of a Linear whose INPUT was relu(.) : dh = (dy W) * (h > 0) in one pass (gate = h).
extern "C" int pdb_gemm_tf32x3_gated(const float* A, const float* B, const float* B_lo, float* C, const float* gate, int M, int N,
                                     int K, int batch, int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc,
                                     int a_mn, int b_mn, void* stream) {
    PDB_REQUIRE(gate, "gemm_tf32x3_gated: null gate");
    return gemm_impl(A, B, B_lo, C, nullptr, M, N, K, batch, lda, ldb, ldc, sa, sb, sc, a_mn, b_mn, 0, 0, 0, 1, 1, nullptr, 0,
                     as_stream(stream), gate);
}

namespace pdb {
__global__ void __launch_bounds__(256) split_lo_kernel(const float4* __restrict__ x, float4* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 h, l;
        split4(__ldg(x + i), h, l);
        lo[i] = l;
    }
}
}  // namespace pdb

extern "C" int pdb_split_lo(const float* x, float* lo, int64_t n, void* stream) {
    PDB_REQUIRE(x && lo && n >= 0 && n % 4 == 0, "split_lo: null pointer or n not a multiple of 4");
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0, "split_lo: 16-byte alignment");
    if (n == 0) return PDB_OK;
    int64_t n4 = n / 4;
    int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 8 * kNumSMs);
    split_lo_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(lo), n4);
    return launched("split_lo");
}

extern "C" PDB_API int pdb_debug_set_gemm_alo_tmem(int on) {
    g_alo_tmem = on != 0;
    g_deep_ring = on != 2;
    g_narrow_tiles = on != 3;
    return PDB_OK;
}

extern "C" PDB_API int pdb_debug_set_trace(long long* buf) {
    g_trace_ptr = buf;
    return PDB_OK;
}
