// MSDeformAttn of the pixel decoder's ENCODER (queries = the pixels of the value pyramid, Lq == S) with the value
// footprint of a query patch staged in shared memory by TMA.
//
// Reference arithmetic: ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:55-75); the reference's own
// (unused) CUDA kernel is ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304.
//
// Mapping.  A CTA owns a 16 x 8 patch of queries of one level, ONE head, one image (128 (query, head) slots; 8 lanes
// per slot, 4 slots per warp, 4 rounds of 32 slots).  In the encoder a query samples level l around its own position
// scaled to that level, so the patch's footprint in level l is a (16 s_x + 2 HALO + 2) x (8 s_y + 2 HALO + 2) pixel
// rectangle (s = W_l / W_lq).  For every level that is not finer than the query level (s <= 1) that rectangle of the
// head's 128-byte rows is fetched with ONE cp.async.bulk.tensor (5-D map over (D, M, W_l, H_l, N); out-of-map pixels are
// zero-filled by the TMA unit, which is exactly grid_sample's padding_mode="zeros") and every tap whose 2 x 2 footprint
// lies inside it reads its four corners with conflict-free LDS.128 (8 lanes x 16 B = all 32 banks).  Taps outside the
// rectangle (offsets beyond HALO pixels) and levels finer than the query level (their footprints do not overlap, a tile
// would cost more shared-memory traffic than it saves) take the global path of msda.cu (clamped 2 x 2 footprint, LDG.128).
// Results do not depend on which path a tap takes: same weights, same products, same summation order.  (Against msda.cu's
// kernels only the order in which the left and the right pixel column of a tap are summed differs.)
//
// Reduced-byte variant (opt-in, msda_fwd_tma<true>): the value pyramid repacked once per call as fp16, head-major
// (N, M, S, 32): a horizontal corner pair is then 128 contiguous bytes = ONE shared-memory wavefront, i.e. 2 instead of
// 4 wavefronts per tap, which is what bounds this kernel.  Lanes 0-3 of a slot take the left pixel
// (8 channels each), lanes 4-7 the right one; the two halves are combined with one shuffle per channel at the end.
// Weights, products and accumulation stay fp32; only the stored value is rounded (11-bit mantissa).
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace pdb {

constexpr int kSPW = 16, kSPH = 8;               // query patch of a CTA
constexpr int kSPSlots = kSPW * kSPH;            // 128 slots
constexpr int kRounds = kSPSlots / 32;           // 4 rounds of 32 slots (256 threads, 8 lanes per slot)
constexpr int kHalo = 4;                         // offsets up to +-4 pixels stay inside the tile
constexpr int kTmaLevels = 4;                    // tensor maps are kernel parameters: L <= 4 on this path
constexpr int kP = 4;
constexpr int kMaxTilePx = 480;                    // largest tile of a stage (26 x 18 = 468 pixels for equal levels)

struct TileGeom {
    int first[kTmaLevels + 1];                   // first patch index of each level, [L] = patches per image
    int px[kTmaLevels];                          // patches per row of each level
    int tw[kTmaLevels][kTmaLevels];              // [query level][sampled level] tile width in pixels, 0 = no tile
    int th[kTmaLevels][kTmaLevels];
    float sx[kTmaLevels][kTmaLevels], sy[kTmaLevels][kTmaLevels];     // W_l / W_q, H_l / H_q
    int per_round[kTmaLevels][kTmaLevels];       // 1: the tile covers one round's 8 x 4 sub-patch (levels finer than the
                                                 // query level: the whole patch's footprint would not fit a stage)
};

struct TileMaps {
    CUtensorMap m[kTmaLevels][kTmaLevels];       // [query level][sampled level]; box = (32, 1, tw, th, 1) / (32, tw, th, 1)
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// One tap.  In-tile: w = the 4 bilinear weights x attention weight, off = element offset of the anchor pixel inside the
// tile (>= 0).  Otherwise the clamped-footprint form of msda.cu: off = -1 - (element offset of the clamped anchor
// inside the head plane).  `px_stride` = elements between horizontally adjacent pixels in global memory.
__device__ __forceinline__ void tile_tap(float lx, float ly, float a, int H, int W, int level_start, int px_stride,
                                         int tx0, int ty0, int tw, int th, float4& w, int& off) {
    float gx = __fsub_rn(__fmul_rn(2.f, lx), 1.f);
    float gy = __fsub_rn(__fmul_rn(2.f, ly), 1.f);
    float x = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
    float y = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
    x = fminf(fmaxf(x, -2.f), (float)W + 1.f);      // keeps the float->int conversion defined for wild offsets
    y = fminf(fmaxf(y, -2.f), (float)H + 1.f);
    const float x0f = floorf(x), y0f = floorf(y);
    const float wx1 = x - x0f, wy1 = y - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int rx = x0 - tx0, ry = y0 - ty0;
    if (rx >= 0 && rx + 1 < tw && ry >= 0 && ry + 1 < th) {
        // the tile holds zeros outside the map, so the four corners keep their own weights
        w = make_float4(a * (wy0 * wx0), a * (wy0 * wx1), a * (wy1 * wx0), a * (wy1 * wx1));
        off = (ry * tw + rx) * 32;
        return;
    }
    const int xb = min(max(x0, 0), W - 2), yb = min(max(y0, 0), H - 2);
    const float wxa = (x0 == xb ? wx0 : 0.f) + (x0 + 1 == xb ? wx1 : 0.f);
    const float wxb = (x0 == xb + 1 ? wx0 : 0.f) + (x0 == xb ? wx1 : 0.f);
    const float wya = (y0 == yb ? wy0 : 0.f) + (y0 + 1 == yb ? wy1 : 0.f);
    const float wyb = (y0 == yb + 1 ? wy0 : 0.f) + (y0 == yb ? wy1 : 0.f);
    w = make_float4(a * (wya * wxa), a * (wya * wxb), a * (wyb * wxa), a * (wyb * wxb));
    off = -1 - (level_start + yb * W + xb) * px_stride;
}

// slot `sl` (0..31) of round r -> query pixel inside the CTA's patch
__device__ __forceinline__ void round_xy(int r, int sl, int& dx, int& dy) {
    dx = (r & 1) * 8 + (sl & 7);
    dy = (r >> 1) * 4 + (sl >> 3);
}

// Work item = (image, patch, head); unit = (item, level).  CTAs are persistent: CTA c takes items c, c + grid, ...
struct Item {
    int n, m, lq, qx0, qy0;
};
__device__ __forceinline__ Item decode_item(int item, const TileGeom& tg, int L, int M) {
    Item it;
    const int per_image = tg.first[L] * M;
    it.n = item / per_image;
    const int rem = item - it.n * per_image;
    const int patch = rem / M;
    it.m = rem - patch * M;                     // heads innermost: the 8 heads of a pixel share DRAM pages
    int lq = 0;
    while (lq + 1 < L && patch >= tg.first[lq + 1]) ++lq;
    it.lq = lq;
    const int pl = patch - tg.first[lq];
    const int py = pl / tg.px[lq];
    it.qx0 = (pl - py * tg.px[lq]) * kSPW;
    it.qy0 = py * kSPH;
    return it;
}
// tile origin in level l for the patch of `it`: anchor of the left-most / top-most query minus the halo
// (qx, qy): first query of the patch, or of one round's sub-patch
__device__ __forceinline__ void tile_origin(int qx, int qy, float sx, float sy, int& tx0, int& ty0) {
    tx0 = (int)floorf(((float)qx + 0.5f) * sx - 0.5f) - kHalo;
    ty0 = (int)floorf(((float)qy + 0.5f) * sy - 0.5f) - kHalo;
}

// One tap's gather for one lane.  fp32: the lane owns channels [chA, chA + 4) and [chB, chB + 4) of ITS pixel of the corner
// pair (rows top, bottom): 4 x 16 bytes; fp16: 8 channels of its pixel: 2 x 16 bytes.  acc: 4 float2 (fp32: A01 A23 B01 B23;
// fp16: channels 01 23 45 67).  FFMA2 = two IEEE fmas per instruction.
template <bool HALF>
__device__ __forceinline__ void tap_fma(float2 (&acc)[4], const float2 w, const uint4 t0, const uint4 t1, const uint4 b0,
                                        const uint4 b1) {
    const float2 wt = make_float2(w.x, w.x), wb = make_float2(w.y, w.y);
    if (!HALF) {
        acc[0] = __ffma2_rn(wt, make_float2(__uint_as_float(t0.x), __uint_as_float(t0.y)), acc[0]);
        acc[1] = __ffma2_rn(wt, make_float2(__uint_as_float(t0.z), __uint_as_float(t0.w)), acc[1]);
        acc[2] = __ffma2_rn(wt, make_float2(__uint_as_float(t1.x), __uint_as_float(t1.y)), acc[2]);
        acc[3] = __ffma2_rn(wt, make_float2(__uint_as_float(t1.z), __uint_as_float(t1.w)), acc[3]);
        acc[0] = __ffma2_rn(wb, make_float2(__uint_as_float(b0.x), __uint_as_float(b0.y)), acc[0]);
        acc[1] = __ffma2_rn(wb, make_float2(__uint_as_float(b0.z), __uint_as_float(b0.w)), acc[1]);
        acc[2] = __ffma2_rn(wb, make_float2(__uint_as_float(b1.x), __uint_as_float(b1.y)), acc[2]);
        acc[3] = __ffma2_rn(wb, make_float2(__uint_as_float(b1.z), __uint_as_float(b1.w)), acc[3]);
    } else {
        const __half2* t2 = reinterpret_cast<const __half2*>(&t0);
        const __half2* b2 = reinterpret_cast<const __half2*>(&b0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[c] = __ffma2_rn(wt, __half22float2(t2[c]), acc[c]);
            acc[c] = __ffma2_rn(wb, __half22float2(b2[c]), acc[c]);
        }
    }
}

// Persistent, warp-specialised: the last warp streams the value tiles of the CTA's (item, level[, round]) steps through an
// NSTAGE ring (cp.async.bulk.tensor + full / empty mbarriers); the CW consumer warps set up the tap records of a level
// (sampling locations and weights prefetched one level ahead into registers), wait for its tile and gather.  Lane roles
// inside a slot (8 lanes): lanes 0-3 take the LEFT pixel of every corner pair, lanes 4-7 the RIGHT one, so a lane needs only
// two of a tap's four weights (one LDS.64); fp32: each lane reads two 16-byte quarters of its pixel per row, chosen so that
// the 8 lanes of a slot always cover all 32 banks; fp16: one 16-byte quarter per row (8 channels).
// A "round" is an 8 x 4 sub-patch (32 slots); a pass of the CTA covers CW / 8 rounds (warps 8 g .. 8 g + 7 take round g of
// the pass); a thread therefore owns PASSES = 32 / CW slots.
template <bool HALF, int NSTAGE, int CW>
__global__ void __launch_bounds__(CW * 32 + 32, (HALF && CW == 8) ? 2 : 1)
msda_fwd_tma(const void* __restrict__ value_, const __grid_constant__ LevelTable lt, const __grid_constant__ TileGeom tg,
             const __grid_constant__ TileMaps maps, const float* __restrict__ loc, const float* __restrict__ attn,
             float* __restrict__ out, int S, int M, int L, int items, int stage_bytes) {
    constexpr int kConsumerThreads = CW * 32;
    constexpr int PASSES = 32 / CW;                 // 4 (8 warps) or 2 (16 warps)
    constexpr int RPP = CW / 8;                     // rounds per pass
    constexpr int RPL = PASSES * kP / 8;            // tap records a lane sets up per level: 2 or 1
    constexpr int ESZ = HALF ? 2 : 4;
    extern __shared__ __align__(128) uint8_t smem_tile[];            // NSTAGE tiles of stage_bytes
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE];
    __shared__ __align__(8) float2 s_w[kSPSlots * kP * 2];          // [pass][slot][tap][left | right] = (w_top, w_bot)
    __shared__ int s_off[kSPSlots * kP];                            // >= 0: byte offset inside the tile; < 0: global path

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], CW);
        }
        tc::fence_barrier_init();
    }
    __syncthreads();

    if (threadIdx.x >= kConsumerThreads) {
        // ------------------------------------------------------------------ producer
        if (threadIdx.x == kConsumerThreads) {
            int k = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const Item it = decode_item(item, tg, L, M);
                for (int l = 0; l < L; ++l) {
                    const int tw = tg.tw[it.lq][l], th = tg.th[it.lq][l];
                    if (tw == 0) continue;                      // no tile for this level: every tap takes the global path
                    const int steps = tg.per_round[it.lq][l] ? kRounds : 1;
                    for (int r = 0; r < steps; ++r, ++k) {
                        const int stage = k % NSTAGE, use = k / NSTAGE;
                        if (use > 0) tc::mbar_wait(&empty[stage], (use - 1) & 1);
                        int tx0, ty0;
                        tile_origin(it.qx0 + (steps > 1 ? (r & 1) * 8 : 0), it.qy0 + (steps > 1 ? (r >> 1) * 4 : 0), tg.sx[it.lq][l],
                                    tg.sy[it.lq][l], tx0, ty0);
                        tc::mbar_expect_tx(&full[stage], (uint32_t)(tw * th * 32 * ESZ));
                        void* dst = smem_tile + (size_t)stage * stage_bytes;
                        if (HALF) tma_load_4d(dst, &maps.m[it.lq][l], &full[stage], 0, tx0, ty0, it.n * M + it.m);
                        else tma_load_5d(dst, &maps.m[it.lq][l], &full[stage], 0, it.m, tx0, ty0, it.n);
                    }
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    const int sl = threadIdx.x >> 3, j = threadIdx.x & 7;   // sl: slot of the pass (0 .. 4 CW - 1)
    const int h = j >> 2, kq = j & 3;                       // h: left / right pixel of the corner pair
    const int grp = sl >> 5, s5 = sl & 31;                  // round of the pass this warp works on, slot inside the round
    const int LP = L * kP;
    const int px_stride = HALF ? 32 : M * 32;              // elements between horizontally adjacent pixels (global memory)
    const int chA = (h ? 16 : 0) + 4 * kq;                 // fp32: first channel quarter of this lane (second: chA ^ 16)
    // byte offset of this lane inside a tile pixel pair, and from its first to its second quarter (fp32)
    const int lane_byte = HALF ? j * 16 : (h * 32 + chA) * 4;
    const int d2 = h ? -64 : 64;
    // record set-up role: records e = j RPL + i (i < RPL) of slot sl: pass e / 4, tap e % 4
    int s_dx[RPL > 1 ? 1 : 1], s_dy[1];
    const int s_pass = (j * RPL) / kP, s_tap0 = (j * RPL) % kP;
    const int s_round = s_pass * RPP + grp;
    s_dx[0] = (s_round & 1) * 8 + (s5 & 7);
    s_dy[0] = (s_round >> 1) * 4 + (s5 >> 3);

    float pre_l[2 * RPL], pre_a[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) pre_l[2 * i] = pre_l[2 * i + 1] = pre_a[i] = 0.f;
    auto prefetch = [&](const Item& it, int l) {
        if (it.qx0 + s_dx[0] < lt.w[it.lq] && it.qy0 + s_dy[0] < lt.h[it.lq]) {
            const int64_t g = ((int64_t)it.n * S + lt.start[it.lq] + (it.qy0 + s_dy[0]) * lt.w[it.lq] + it.qx0 + s_dx[0]) * M + it.m;
            const float2* lp = reinterpret_cast<const float2*>(loc) + g * LP + l * kP + s_tap0;
            const float* ap = attn + g * LP + l * kP + s_tap0;
            if (RPL == 2) {
                float4 t;
                float2 a;
                asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(lp));
                asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(a.x), "=f"(a.y) : "l"(ap));
                pre_l[0] = t.x; pre_l[1] = t.y; pre_l[2 * RPL - 2] = t.z; pre_l[2 * RPL - 1] = t.w;
                pre_a[0] = a.x; pre_a[RPL - 1] = a.y;
            } else {
                float2 t;
                float a;
                asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "l"(lp));
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(a) : "l"(ap));
                pre_l[0] = t.x; pre_l[1] = t.y;
                pre_a[0] = a;
            }
        } else {
#pragma unroll
            for (int i = 0; i < RPL; ++i) pre_l[2 * i] = pre_l[2 * i + 1] = pre_a[i] = 0.f;   // weight 0, in-range location
        }
    };

    int k = 0;
    if ((int)blockIdx.x < items) prefetch(decode_item(blockIdx.x, tg, L, M), 0);
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const Item it = decode_item(item, tg, L, M);
        float2 acc[PASSES][4];
#pragma unroll
        for (int r = 0; r < PASSES; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
            const int H = lt.h[l], W = lt.w[l];
            const int tw = tg.tw[it.lq][l], th = tg.th[it.lq][l];
            const bool per_round = tg.per_round[it.lq][l] != 0;
            {   // records of this level from the values prefetched while the previous one was gathered; the tile they are
                // relative to is the patch's, or the one of the record's round
                int tx0, ty0;
                tile_origin(it.qx0 + (per_round ? (s_round & 1) * 8 : 0), it.qy0 + (per_round ? (s_round >> 1) * 4 : 0),
                            tg.sx[it.lq][l], tg.sy[it.lq][l], tx0, ty0);
#pragma unroll
                for (int i = 0; i < RPL; ++i) {
                    float4 w;
                    int off;
                    tile_tap(pre_l[2 * i], pre_l[2 * i + 1], pre_a[i], H, W, lt.start[l], px_stride, tx0, ty0, tw, th, w, off);
                    const int rs = (s_pass * (CW * 4) + sl) * kP + s_tap0 + i;
                    s_w[rs * 2] = make_float2(w.x, w.z);
                    s_w[rs * 2 + 1] = make_float2(w.y, w.w);
                    s_off[rs] = off >= 0 ? off * ESZ : off;
                }
            }
            // next level's locations / weights: in flight while this one is gathered
            if (l + 1 < L) prefetch(it, l + 1);
            else if (item + (int)gridDim.x < items) prefetch(decode_item(item + gridDim.x, tg, L, M), 0);
            __syncwarp();
            const int down_s = tw * 32 * ESZ;                                    // bytes
            const int down_g = HALF ? W * 32 : W * M * 32;                       // elements
            const uint8_t* vbg = HALF ? reinterpret_cast<const uint8_t*>(reinterpret_cast<const __half*>(value_) +
                                                                         ((int64_t)(it.n * M + it.m) * S) * 32 + j * 8)
                                      : reinterpret_cast<const uint8_t*>(reinterpret_cast<const float*>(value_) +
                                                                         ((int64_t)it.n * S * M + it.m) * 32 + h * (M * 32) + chA);
            int stage = 0;
            if (tw > 0 && !per_round) {
                stage = k % NSTAGE;
                tc::mbar_wait(&full[stage], (k / NSTAGE) & 1);
            }
#pragma unroll
            for (int r = 0; r < PASSES; ++r) {
                if (per_round) {
                    stage = (k + r * RPP + grp) % NSTAGE;
                    tc::mbar_wait(&full[stage], ((k + r * RPP + grp) / NSTAGE) & 1);
                }
                const uint8_t* tile = smem_tile + (size_t)stage * stage_bytes + lane_byte;
                const int t0 = (r * (CW * 4) + sl) * kP;
                int off[kP];
#pragma unroll
                for (int p = 0; p < kP; ++p) off[p] = s_off[t0 + p];
                if (!__any_sync(0xffffffffu, (off[0] | off[1] | off[2] | off[3]) < 0)) {
                    // every tap of the warp's 4 slots lies inside the tile: straight-line LDS.128 + FFMA2
#pragma unroll
                    for (int ph = 0; ph < kP; ph += (HALF ? 4 : 2)) {
                        constexpr int G = HALF ? 4 : 2;
                        float2 w[G];
                        uint4 a0[G], a1[G], b0[G], b1[G];
#pragma unroll
                        for (int p = 0; p < G; ++p) {
                            w[p] = s_w[(t0 + ph + p) * 2 + h];
                            const uint8_t* p0 = tile + off[ph + p];
                            a0[p] = *reinterpret_cast<const uint4*>(p0);
                            b0[p] = *reinterpret_cast<const uint4*>(p0 + down_s);
                            if (!HALF) {
                                a1[p] = *reinterpret_cast<const uint4*>(p0 + d2);
                                b1[p] = *reinterpret_cast<const uint4*>(p0 + down_s + d2);
                            } else {
                                a1[p] = b1[p] = make_uint4(0, 0, 0, 0);
                            }
                        }
#pragma unroll
                        for (int p = 0; p < G; ++p) tap_fma<HALF>(acc[r], w[p], a0[p], a1[p], b0[p], b1[p]);
                    }
                } else {
#pragma unroll 1
                    for (int p = 0; p < kP; ++p) {
                        const float2 w = s_w[(t0 + p) * 2 + h];
                        uint4 a0, a1 = make_uint4(0, 0, 0, 0), b0, b1 = make_uint4(0, 0, 0, 0);
                        int o = off[0];
                        o = p == 1 ? off[1] : o;
                        o = p == 2 ? off[2] : o;
                        o = p == 3 ? off[3] : o;
                        if (o >= 0) {
                            const uint8_t* p0 = tile + o;
                            a0 = *reinterpret_cast<const uint4*>(p0);
                            b0 = *reinterpret_cast<const uint4*>(p0 + down_s);
                            if (!HALF) {
                                a1 = *reinterpret_cast<const uint4*>(p0 + d2);
                                b1 = *reinterpret_cast<const uint4*>(p0 + down_s + d2);
                            }
                        } else {
                            const uint8_t* p0 = vbg + (int64_t)(-1 - o) * ESZ;
                            a0 = __ldg(reinterpret_cast<const uint4*>(p0));
                            b0 = __ldg(reinterpret_cast<const uint4*>(p0 + (int64_t)down_g * ESZ));
                            if (!HALF) {
                                a1 = __ldg(reinterpret_cast<const uint4*>(p0 + d2));
                                b1 = __ldg(reinterpret_cast<const uint4*>(p0 + (int64_t)down_g * ESZ + d2));
                            }
                        }
                        tap_fma<HALF>(acc[r], w, a0, a1, b0, b1);
                    }
                }
                if (per_round) {
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0)
                        for (int c = 0; c < RPP; ++c) tc::mbar_arrive(&empty[stage]);      // 8 warps use this tile: CW arrivals
                }
            }
            __syncwarp();
            if (tw > 0) {
                if (!per_round && (threadIdx.x & 31) == 0) tc::mbar_arrive(&empty[stage]);
                k += per_round ? kRounds : 1;
            }
        }
        // store the item: lanes h = 0 / 1 exchange the halves they accumulated for each other
#pragma unroll
        for (int r = 0; r < PASSES; ++r) {
            const int round = r * RPP + grp;
            const int dx = (round & 1) * 8 + (s5 & 7), dy = (round >> 1) * 4 + (s5 >> 3);
            const bool ok = it.qx0 + dx < lt.w[it.lq] && it.qy0 + dy < lt.h[it.lq];
            const int64_t g = ((int64_t)it.n * S + lt.start[it.lq] + (it.qy0 + dy) * lt.w[it.lq] + it.qx0 + dx) * M + it.m;
            if (!HALF) {
                // this lane: channels chA.. of its pixel in acc[0..1], chA ^ 16 .. in acc[2..3]; the partner lane (j ^ 4) holds the
                // other pixel's share of the same channels in the opposite halves
                float4 o;
                o.x = acc[r][0].x + __shfl_xor_sync(0xffffffffu, acc[r][2].x, 4);
                o.y = acc[r][0].y + __shfl_xor_sync(0xffffffffu, acc[r][2].y, 4);
                o.z = acc[r][1].x + __shfl_xor_sync(0xffffffffu, acc[r][3].x, 4);
                o.w = acc[r][1].y + __shfl_xor_sync(0xffffffffu, acc[r][3].y, 4);
                if (ok) *reinterpret_cast<float4*>(out + g * 32 + chA) = o;
            } else {
                // both lanes of a pair end with the full sums of channels 8 kq .. + 7; lane h stores channels 8 kq + 4 h .. + 3
                float t[8];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    t[2 * c] = acc[r][c].x + __shfl_xor_sync(0xffffffffu, acc[r][c].x, 4);
                    t[2 * c + 1] = acc[r][c].y + __shfl_xor_sync(0xffffffffu, acc[r][c].y, 4);
                }
                const float4 o = h ? make_float4(t[4], t[5], t[6], t[7]) : make_float4(t[0], t[1], t[2], t[3]);
                if (ok) *reinterpret_cast<float4*>(out + g * 32 + kq * 8 + h * 4) = o;
            }
        }
    }
}

// (N, S, M, 32) f32 -> (N, M, S, 32) f16 (round to nearest even); one thread per 8 channels
__global__ void __launch_bounds__(256)
msda_pack_value_h(const float* __restrict__ value, __half* __restrict__ out, int64_t total8, int S, int M) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total8) return;
    const int c8 = (int)(i & 3);
    const int64_t px = i >> 2;                   // (n * S + s) * M + m
    const int m = (int)(px % M);
    const int64_t ns = px / M;
    const int64_t n = ns / S, s = ns - n * S;
    const float4 a = __ldg(reinterpret_cast<const float4*>(value + px * 32 + c8 * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(value + px * 32 + c8 * 8 + 4));
    __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y),
                    __floats2half2_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(out + (((n * M + m) * S + s) * 32 + c8 * 8)) = *reinterpret_cast<const uint4*>(h);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFnN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                        const cuuint64_t* strides, const cuuint32_t* box) {
    static EncodeTiledFnN encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(PDB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
        encode = reinterpret_cast<EncodeTiledFnN>(fn);
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PDB_ERR_INVALID, "msda: cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
    return PDB_OK;
}

// geometry of the patch grid and of the value tiles; returns the number of CTAs and the largest tile (pixels)
static int64_t make_tile_geom(const LevelTable& lt, int L, int N, int M, TileGeom& tg, int& max_px) {
    int total = 0;
    for (int l = 0; l < L; ++l) {
        tg.first[l] = total;
        tg.px[l] = (lt.w[l] + kSPW - 1) / kSPW;
        total += tg.px[l] * ((lt.h[l] + kSPH - 1) / kSPH);
    }
    tg.first[L] = total;
    max_px = 0;
    for (int q = 0; q < L; ++q)
        for (int l = 0; l < L; ++l) {
            tg.tw[q][l] = tg.th[q][l] = tg.per_round[q][l] = 0;
            const double sx = (double)lt.w[l] / lt.w[q], sy = (double)lt.h[l] / lt.h[q];
            tg.sx[q][l] = (float)sx;
            tg.sy[q][l] = (float)sy;
            // footprint of the whole 16 x 8 patch; if that does not fit a stage (levels finer than the query level), of one
            // round's 8 x 4 sub-patch; if that does not fit either the level is gathered from global memory
            int tw = (int)ceil((kSPW - 1) * sx) + 1 + 2 * kHalo + 2, th = (int)ceil((kSPH - 1) * sy) + 1 + 2 * kHalo + 2;
            int pr = 0;
            if (tw * th > kMaxTilePx) {
                tw = (int)ceil(7 * sx) + 1 + 2 * kHalo + 2;
                th = (int)ceil(3 * sy) + 1 + 2 * kHalo + 2;
                pr = 1;
            }
            if (tw * th > kMaxTilePx || tw > 256 || th > 256) continue;
            tg.tw[q][l] = tw;
            tg.th[q][l] = th;
            tg.per_round[q][l] = pr;
            if (tw * th > max_px) max_px = tw * th;
        }
    return (int64_t)total * M * N;
}

bool msda_tma_eligible(const LevelTable& lt, int L, int S, int M, int D, int Lq, int P) {
    if (!(D == 32 && P == kP && Lq == S && L >= 1 && L <= kTmaLevels && (int64_t)S * M * 32 < (1ll << 31))) return false;
    for (int l = 0; l < L; ++l)
        if (lt.h[l] < 2 || lt.w[l] < 2) return false;
    // queries must be exactly the pyramid's pixels, levels stored back to back
    int s = 0;
    for (int l = 0; l < L; ++l) {
        if (lt.start[l] != s) return false;
        s += lt.h[l] * lt.w[l];
    }
    return s == S;
}

static int g_tma_variant = 0;     // experiments: 1 = fp16 kernel as 2 CTAs / SM x 8 consumer warps, 2-stage rings

template <bool HALF>
static int launch_fwd_tma(const void* value, const LevelTable& lt, const float* loc, const float* attn, float* out, int N, int S,
                          int M, int L, cudaStream_t st) {
    TileGeom tg;
    TileMaps maps;
    int max_px = 0;
    const int64_t blocks = make_tile_geom(lt, L, N, M, tg, max_px);
    PDB_REQUIRE(blocks < (1ll << 31), "msda_forward: too many CTAs");
    memset(&maps, 0, sizeof(maps));
    for (int q = 0; q < L; ++q)
        for (int l = 0; l < L; ++l) {
            if (!tg.tw[q][l]) continue;
            if (HALF) {
                const __half* base = (const __half*)value + (int64_t)lt.start[l] * 32;
                cuuint64_t dims[4] = {32, (cuuint64_t)lt.w[l], (cuuint64_t)lt.h[l], (cuuint64_t)N * M};
                cuuint64_t strides[3] = {64, (cuuint64_t)lt.w[l] * 64, (cuuint64_t)S * 64};
                cuuint32_t box[4] = {32, (cuuint32_t)tg.tw[q][l], (cuuint32_t)tg.th[q][l], 1};
                PDB_TRY(encode_tiled(&maps.m[q][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box));
            } else {
                const float* base = (const float*)value + (int64_t)lt.start[l] * M * 32;
                cuuint64_t dims[5] = {32, (cuuint64_t)M, (cuuint64_t)lt.w[l], (cuuint64_t)lt.h[l], (cuuint64_t)N};
                cuuint64_t strides[4] = {128, (cuuint64_t)M * 128, (cuuint64_t)lt.w[l] * M * 128, (cuuint64_t)S * M * 128};
                cuuint32_t box[5] = {32, 1, (cuuint32_t)tg.tw[q][l], (cuuint32_t)tg.th[q][l], 1};
                PDB_TRY(encode_tiled(&maps.m[q][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box));
            }
        }
    const int stage_bytes = ((max_px * 32 * (HALF ? 2 : 4)) + 127) & ~127;
    auto go = [&](auto kern, int nstage, int cw, int ctas_per_sm, size_t& attr) -> int {
        const size_t smem = (size_t)stage_bytes * nstage;
        PDB_REQUIRE(smem <= 200 * 1024, "msda_forward: tile ring of %zu bytes", smem);
        if (smem > attr) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "msda_fwd_tma: smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
            attr = smem;
        }
        const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
        const int64_t grid = blocks < cap ? blocks : cap;
        kern<<<(unsigned)grid, cw * 32 + 32, smem, st>>>(value, lt, tg, maps, loc, attn, out, S, M, L, (int)blocks, stage_bytes);
        return launched(HALF ? "msda_fwd_tma<half>" : "msda_fwd_tma");
    };
    static size_t attr[3] = {0, 0, 0};
    if (!HALF) return go(msda_fwd_tma<false, 3, 16>, 3, 16, 1, attr[0]);
    if (g_tma_variant == 1) return go(msda_fwd_tma<true, 2, 8>, 2, 8, 2, attr[2]);
    return go(msda_fwd_tma<true, 4, 16>, 4, 16, 1, attr[1]);
}

int g_tma_variant_set(int v) {
    g_tma_variant = v;
    return 0;
}

int msda_forward_tma_f32(const void* value, const LevelTable& lt, const float* loc, const float* attn, float* out, int N, int S,
                         int M, int L, cudaStream_t st) {
    return launch_fwd_tma<false>(value, lt, loc, attn, out, N, S, M, L, st);
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_msda_pack_value_h(const float* value, void* value_h, int N, int S, int M, int D, void* stream) {
    PDB_REQUIRE(value && value_h, "msda_pack_value_h: null pointer");
    PDB_REQUIRE(D == 32 && N > 0 && S > 0 && M > 0, "msda_pack_value_h: D must be 32");
    const int64_t total8 = (int64_t)N * S * M * 4;
    msda_pack_value_h<<<(unsigned)((total8 + 255) / 256), 256, 0, as_stream(stream)>>>(value, (__half*)value_h, total8, S, M);
    return launched("msda_pack_value_h");
}

extern "C" int pdb_msda_forward_h(const void* value_h, const int64_t* shapes_hw, const int64_t* level_start, const float* loc,
                                  const float* attn, float* out, int N, int S, int M, int D, int Lq, int L, int P, void* stream) {
    PDB_REQUIRE(value_h && loc && attn && out && shapes_hw && level_start, "msda_forward_h: null pointer");
    PDB_REQUIRE(L >= 1 && L <= kTmaLevels, "msda_forward_h: L=%d outside [1,%d]", L, kTmaLevels);
    LevelTable lt;
    for (int l = 0; l < L; ++l) {
        lt.h[l] = (int)shapes_hw[2 * l];
        lt.w[l] = (int)shapes_hw[2 * l + 1];
        lt.start[l] = (int)level_start[l];
    }
    PDB_REQUIRE(msda_tma_eligible(lt, L, S, M, D, Lq, P),
                "msda_forward_h: needs D=32, P=4, Lq == S (encoder self-attention over the value pyramid), levels >= 2x2");
    return launch_fwd_tma<true>(value_h, lt, loc, attn, out, N, S, M, L, as_stream(stream));
}
