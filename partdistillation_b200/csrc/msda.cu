// Multi-scale deformable attention: bilinear gather (forward) and scatter (backward).
//
// Semantics follow the reference's production arithmetic, ms_deform_attn_core_pytorch
// (ops/functions/ms_deform_attn_func.py:55-75): per level F.grid_sample(value_l, 2*loc-1, bilinear,
// zeros, align_corners=False), times the attention weight, summed over levels and points.  The
// reference's own (unused) CUDA kernel is ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304 (forward)
// and :306-408 (backward, D=32 variant).
//
// Layout in HBM (unchanged from the reference operator boundary):
//   value (N, S, M, D): one (pixel, head) = D contiguous floats = one 128-byte line for D=32.
//   loc (N, Lq, M, L, P, 2), attn (N, Lq, M, L, P), out (N, Lq, M*D).
//
// Fast path (f32, D == 32): a "slot" is one (n, q, m).  8 lanes share a slot, each lane owns 4
// channels and issues LDG.128, so a warp instruction moves 4 slots x 128 B: the L1 data path
// (128 B/clk/SM), not the LSU issue rate, is the limiter.  Sampling locations and weights of the
// 32 slots of a CTA are staged through shared memory with coalesced loads and read back as
// broadcasts.
#include "common.cuh"

namespace pdb {

// ------------------------------------------------------------------------------------------------
// generic path: one thread per (slot, channel); float or double; any D
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void msda_fwd_generic(const T* __restrict__ value, const __grid_constant__ LevelTable lt, const T* __restrict__ loc,
                                 const T* __restrict__ attn, T* __restrict__ out, int64_t total, int S, int M,
                                 int D, int Lq, int L, int P) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int d = (int)(idx % D);
    int64_t g = idx / D;           // slot = (n*Lq + q)*M + m
    int m = (int)(g % M);
    int64_t n = g / M / Lq;
    const T* lp = loc + g * L * P * 2;
    const T* ap = attn + g * L * P;
    T acc = 0;
    for (int l = 0; l < L; ++l) {
        const int H = lt.h[l], W = lt.w[l];
        const T* vb = value + ((n * S + lt.start[l]) * M + m) * (int64_t)D + d;
        for (int p = 0; p < P; ++p) {
            T gx = T(2) * lp[(l * P + p) * 2] - T(1);
            T gy = T(2) * lp[(l * P + p) * 2 + 1] - T(1);
            T x = ((gx + T(1)) * T(W) - T(1)) * T(0.5);
            T y = ((gy + T(1)) * T(H) - T(1)) * T(0.5);
            T x0f = floor(x), y0f = floor(y);
            T wx1 = x - x0f, wy1 = y - y0f, wx0 = T(1) - wx1, wy0 = T(1) - wy1;
            int x0 = (int)x0f, y0 = (int)y0f;
            T s = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int xi = x0 + (c & 1), yi = y0 + (c >> 1);
                if (xi >= 0 && xi < W && yi >= 0 && yi < H) {
                    T w = ((c & 1) ? wx1 : wx0) * ((c >> 1) ? wy1 : wy0);
                    s += w * vb[((int64_t)yi * W + xi) * M * D];
                }
            }
            acc += ap[l * P + p] * s;
        }
    }
    out[idx] = acc;
}

template <typename T>
__global__ void msda_bwd_generic(const T* __restrict__ value, const __grid_constant__ LevelTable lt, const T* __restrict__ loc,
                                 const T* __restrict__ attn, const T* __restrict__ grad_out,
                                 T* __restrict__ grad_value, T* __restrict__ grad_loc, T* __restrict__ grad_attn,
                                 int64_t total, int S, int M, int D, int Lq, int L, int P) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int d = (int)(idx % D);
    int64_t g = idx / D;
    int m = (int)(g % M);
    int64_t n = g / M / Lq;
    const T* lp = loc + g * L * P * 2;
    const T* ap = attn + g * L * P;
    const T go = grad_out[idx];
    for (int l = 0; l < L; ++l) {
        const int H = lt.h[l], W = lt.w[l];
        const int64_t base = ((n * S + lt.start[l]) * M + m) * (int64_t)D + d;
        for (int p = 0; p < P; ++p) {
            T gx = T(2) * lp[(l * P + p) * 2] - T(1);
            T gy = T(2) * lp[(l * P + p) * 2 + 1] - T(1);
            T x = ((gx + T(1)) * T(W) - T(1)) * T(0.5);
            T y = ((gy + T(1)) * T(H) - T(1)) * T(0.5);
            T x0f = floor(x), y0f = floor(y);
            T wx1 = x - x0f, wy1 = y - y0f, wx0 = T(1) - wx1, wy0 = T(1) - wy1;
            int x0 = (int)x0f, y0 = (int)y0f;
            T a = ap[l * P + p];
            T v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int xi = x0 + (c & 1), yi = y0 + (c >> 1);
                v[c] = 0;
                if (xi >= 0 && xi < W && yi >= 0 && yi < H) {
                    int64_t off = base + ((int64_t)yi * W + xi) * M * D;
                    v[c] = value[off];
                    T w = ((c & 1) ? wx1 : wx0) * ((c >> 1) ? wy1 : wy0);
                    atomicAdd(grad_value + off, w * a * go);
                }
            }
            T ga = (wy0 * (wx0 * v[0] + wx1 * v[1]) + wy1 * (wx0 * v[2] + wx1 * v[3])) * go;
            T gxl = (wy0 * (v[1] - v[0]) + wy1 * (v[3] - v[2])) * go * a * T(W);
            T gyl = (wx0 * (v[2] - v[0]) + wx1 * (v[3] - v[1])) * go * a * T(H);
            atomicAdd(grad_attn + g * L * P + l * P + p, ga);
            atomicAdd(grad_loc + (g * L * P + l * P + p) * 2, gxl);
            atomicAdd(grad_loc + (g * L * P + l * P + p) * 2 + 1, gyl);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fast path: f32, D == 32, 8 lanes per slot, 32 slots per 256-thread CTA
// ------------------------------------------------------------------------------------------------
constexpr int kSlotsPerCta = 32;
constexpr int kFastThreads = 256;

struct Tap {
    float w[4];      // bilinear weight x attention weight per corner (0 when the corner is outside)
    int off[4];      // float offset of the corner inside the level (clamped when outside)
};

__device__ __forceinline__ void tap_setup(float lx, float ly, int H, int W, int rowstride,
                                          float& wx0, float& wx1, float& wy0, float& wy1, int off[4], bool ok[4]) {
    // grid_sample(align_corners=False) un-normalisation of g = 2*loc - 1, evaluated without FMA
    // contraction so that it follows the fp32 reference step by step.
    float gx = __fsub_rn(__fmul_rn(2.f, lx), 1.f);
    float gy = __fsub_rn(__fmul_rn(2.f, ly), 1.f);
    float x = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
    float y = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
    // keep the float->int conversion defined for wild (learned) offsets
    x = fminf(fmaxf(x, -2.f), (float)W + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)H + 1.f);
    float x0f = floorf(x), y0f = floorf(y);
    wx1 = x - x0f; wy1 = y - y0f; wx0 = 1.f - wx1; wy0 = 1.f - wy1;
    int x0 = (int)x0f, y0 = (int)y0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int xi = x0 + (c & 1), yi = y0 + (c >> 1);
        ok[c] = (xi >= 0) && (xi < W) && (yi >= 0) && (yi < H);
        int xc = min(max(xi, 0), W - 1), yc = min(max(yi, 0), H - 1);
        off[c] = (yc * W + xc) * rowstride;
    }
}

template <int P>
__global__ void __launch_bounds__(kFastThreads)
msda_fwd_d32(const float* __restrict__ value, const __grid_constant__ LevelTable lt, const float* __restrict__ loc,
             const float* __restrict__ attn, float* __restrict__ out, int64_t slots, int S, int M, int Lq, int L) {
    extern __shared__ float smem[];
    const int LP = L * P;
    const int loc_stride = LP * 2 + 2;
    const int attn_stride = LP + 1;
    float* s_loc = smem;
    float* s_attn = smem + kSlotsPerCta * loc_stride;

    const int64_t slot0 = (int64_t)blockIdx.x * kSlotsPerCta;
    const int nslots = (int)min((int64_t)kSlotsPerCta, slots - slot0);
    // coalesced staging of the CTA's sampling locations and weights
    {
        const float* gl = loc + slot0 * LP * 2;
        for (int i = threadIdx.x; i < nslots * LP * 2; i += kFastThreads) {
            int s = i / (LP * 2), r = i - s * (LP * 2);
            s_loc[s * loc_stride + r] = __ldg(gl + i);
        }
        const float* ga = attn + slot0 * LP;
        for (int i = threadIdx.x; i < nslots * LP; i += kFastThreads) {
            int s = i / LP, r = i - s * LP;
            s_attn[s * attn_stride + r] = __ldg(ga + i);
        }
    }
    __syncthreads();
    const int sl = threadIdx.x >> 3;     // slot within CTA
    const int j = threadIdx.x & 7;       // 4-channel group
    if (sl >= nslots) return;
    const int64_t g = slot0 + sl;
    const int m = (int)(g % M);
    const int64_t n = g / M / Lq;
    const int rowstride = M * 32;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* myloc = s_loc + sl * loc_stride;
    const float* myattn = s_attn + sl * attn_stride;
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        const int H = lt.h[l], W = lt.w[l];
        const float* vb = value + ((n * S + lt.start[l]) * M + m) * 32 + j * 4;
        float w[P][4];
        int off[P][4];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            float2 lxy = *reinterpret_cast<const float2*>(myloc + (l * P + p) * 2);
            float a = myattn[l * P + p];
            float wx0, wx1, wy0, wy1;
            bool ok[4];
            tap_setup(lxy.x, lxy.y, H, W, rowstride, wx0, wx1, wy0, wy1, off[p], ok);
            w[p][0] = ok[0] ? a * (wy0 * wx0) : 0.f;
            w[p][1] = ok[1] ? a * (wy0 * wx1) : 0.f;
            w[p][2] = ok[2] ? a * (wy1 * wx0) : 0.f;
            w[p][3] = ok[3] ? a * (wy1 * wx1) : 0.f;
        }
        float4 v[P][4];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) v[p][c] = __ldg(reinterpret_cast<const float4*>(vb + off[p][c]));
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc.x = fmaf(w[p][c], v[p][c].x, acc.x);
                acc.y = fmaf(w[p][c], v[p][c].y, acc.y);
                acc.z = fmaf(w[p][c], v[p][c].z, acc.z);
                acc.w = fmaf(w[p][c], v[p][c].w, acc.w);
            }
    }
    *reinterpret_cast<float4*>(out + g * 32 + j * 4) = acc;
}

template <int P>
__global__ void __launch_bounds__(kFastThreads)
msda_bwd_d32(const float* __restrict__ value, const __grid_constant__ LevelTable lt, const float* __restrict__ loc,
             const float* __restrict__ attn, const float* __restrict__ grad_out, float* __restrict__ grad_value,
             float* __restrict__ grad_loc, float* __restrict__ grad_attn, int64_t slots, int S, int M, int Lq, int L) {
    extern __shared__ float smem[];
    const int LP = L * P;
    const int loc_stride = LP * 2 + 2;
    const int attn_stride = LP + 1;
    float* s_loc = smem;
    float* s_attn = s_loc + kSlotsPerCta * loc_stride;
    float* s_gloc = s_attn + kSlotsPerCta * attn_stride;     // dense (slot, LP*2)
    float* s_gattn = s_gloc + kSlotsPerCta * LP * 2;         // dense (slot, LP)

    const int64_t slot0 = (int64_t)blockIdx.x * kSlotsPerCta;
    const int nslots = (int)min((int64_t)kSlotsPerCta, slots - slot0);
    {
        const float* gl = loc + slot0 * LP * 2;
        for (int i = threadIdx.x; i < nslots * LP * 2; i += kFastThreads) {
            int s = i / (LP * 2), r = i - s * (LP * 2);
            s_loc[s * loc_stride + r] = __ldg(gl + i);
        }
        const float* ga = attn + slot0 * LP;
        for (int i = threadIdx.x; i < nslots * LP; i += kFastThreads) {
            int s = i / LP, r = i - s * LP;
            s_attn[s * attn_stride + r] = __ldg(ga + i);
        }
    }
    __syncthreads();
    const int sl = threadIdx.x >> 3;
    const int j = threadIdx.x & 7;
    // Lanes whose slot lies beyond the problem skip the body below but do NOT exit (they wait at the barrier after it),
    // so the butterfly shuffles inside must name only the participating lanes: with a full mask a warp that mixes valid
    // and invalid slots (slots % 4 != 0, i.e. a head count that is not a multiple of 4) would deadlock.
    const unsigned lanes = __ballot_sync(0xffffffffu, sl < nslots);
    if (sl < nslots) {
        const int64_t g = slot0 + sl;
        const int m = (int)(g % M);
        const int64_t n = g / M / Lq;
        const int rowstride = M * 32;
        const float4 go = __ldg(reinterpret_cast<const float4*>(grad_out + g * 32 + j * 4));
        const float* myloc = s_loc + sl * loc_stride;
        const float* myattn = s_attn + sl * attn_stride;
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
            const int H = lt.h[l], W = lt.w[l];
            const int64_t lbase = ((n * S + lt.start[l]) * M + m) * 32 + j * 4;
            const float* vb = value + lbase;
            float* gvb = grad_value + lbase;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float2 lxy = *reinterpret_cast<const float2*>(myloc + (l * P + p) * 2);
                float a = myattn[l * P + p];
                float wx0, wx1, wy0, wy1;
                bool ok[4];
                int off[4];
                tap_setup(lxy.x, lxy.y, H, W, rowstride, wx0, wx1, wy0, wy1, off, ok);
                float dot[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    dot[c] = 0.f;
                    if (ok[c]) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(vb + off[c]));
                        dot[c] = v.x * go.x + v.y * go.y + v.z * go.z + v.w * go.w;
                        float wc = a * (((c >> 1) ? wy1 : wy0) * ((c & 1) ? wx1 : wx0));
                        red_add_v4(gvb + off[c], wc * go.x, wc * go.y, wc * go.z, wc * go.w);
                    }
                }
                float ga = wy0 * (wx0 * dot[0] + wx1 * dot[1]) + wy1 * (wx0 * dot[2] + wx1 * dot[3]);
                float gx = (wy0 * (dot[1] - dot[0]) + wy1 * (dot[3] - dot[2])) * a * (float)W;
                float gy = (wx0 * (dot[2] - dot[0]) + wx1 * (dot[3] - dot[1])) * a * (float)H;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    ga += __shfl_xor_sync(lanes, ga, o);
                    gx += __shfl_xor_sync(lanes, gx, o);
                    gy += __shfl_xor_sync(lanes, gy, o);
                }
                if (j == 0) {
                    s_gattn[sl * LP + l * P + p] = ga;
                    s_gloc[(sl * LP + l * P + p) * 2] = gx;
                    s_gloc[(sl * LP + l * P + p) * 2 + 1] = gy;
                }
            }
        }
    }
    __syncthreads();
    float* ol = grad_loc + slot0 * LP * 2;
    for (int i = threadIdx.x; i < nslots * LP * 2; i += kFastThreads) ol[i] = s_gloc[i];
    float* oa = grad_attn + slot0 * LP;
    for (int i = threadIdx.x; i < nslots * LP; i += kFastThreads) oa[i] = s_gattn[i];
}


// ------------------------------------------------------------------------------------------------
// tiled fast path (f32, D == 32, P == 4, every level at least 2 x 2)
//
// What bounds the gather: one tap corner of one head is one 128-byte line = one L1 wavefront, so a (query, head)
// slot moves LP*4*128 B through the L1 data pipe (128 B/clk/SM) against 448 B of compulsory HBM traffic; the kernel
// can at best be L1-wavefront-bound (~30-36 % of the HBM roofline) and must (1) keep the instruction stream below
// that bound and (2) make the wavefronts hit in L1 instead of L2.
//  (1) The tap set-up (un-normalise, floor, border handling, 4 corner weights) is done ONCE per tap by one of the 8
//      lanes of a slot (lane j sets up taps j and j+8) and handed to the others through shared memory as a 5-word
//      record {w00, w01, w10, w11, offset}; every lane then only issues 2 LDS + 4 LDG.128 + 16 FFMA per tap.
//      Border handling without branches: the 2x2 footprint is clamped into the map (xb in [0, W-2]) and each
//      in-map position receives the weight of the original corner that falls on it (0 if none), so all four loads
//      are always in bounds and relative to one base offset: {0, M*32, W*M*32, (W+1)*M*32} floats.
//  (2) When the queries are the pixels of the value pyramid itself (Lq == S: the encoder), a CTA takes an 8 x 4
//      patch of queries of one level and ONE head, so the 32 slots of a CTA read overlapping footprints of a single
//      head's 128-byte lines and L1 serves the reuse; otherwise slots are taken in memory order.
// ------------------------------------------------------------------------------------------------
constexpr int kPatchW = 8, kPatchH = 4;
constexpr int kTileSlots = kPatchW * kPatchH;       // 32 slots x 8 lanes = 256 threads

struct PatchTable {
    int first[kMaxLevels + 1];      // first patch index of each level (prefix sums), [L] = patches per image
    int px[kMaxLevels];             // patches per row of each level
};

struct TapRec {
    float4 w;                       // weights of the 4 clamped positions, already multiplied by the attention weight
};

// one tap: clamped base offset (floats, relative to the image's head plane) and the 4 position weights
__device__ __forceinline__ void tap_record(float lx, float ly, float a, int H, int W, int level_start, int rowstride,
                                           float4& w, int& off) {
    float gx = __fsub_rn(__fmul_rn(2.f, lx), 1.f);
    float gy = __fsub_rn(__fmul_rn(2.f, ly), 1.f);
    float x = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
    float y = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
    x = fminf(fmaxf(x, -2.f), (float)W + 1.f);      // keeps the float->int conversion defined for wild offsets
    y = fminf(fmaxf(y, -2.f), (float)H + 1.f);
    const float x0f = floorf(x), y0f = floorf(y);
    const float wx1 = x - x0f, wy1 = y - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int xb = min(max(x0, 0), W - 2), yb = min(max(y0, 0), H - 2);
    // weight of clamped position xb / xb+1 = weight of the original corner (x0 or x0+1) lying on it, if any
    const float wxa = (x0 == xb ? wx0 : 0.f) + (x0 + 1 == xb ? wx1 : 0.f);
    const float wxb = (x0 == xb + 1 ? wx0 : 0.f) + (x0 == xb ? wx1 : 0.f);
    const float wya = (y0 == yb ? wy0 : 0.f) + (y0 + 1 == yb ? wy1 : 0.f);
    const float wyb = (y0 == yb + 1 ? wy0 : 0.f) + (y0 == yb ? wy1 : 0.f);
    w = make_float4(a * (wya * wxa), a * (wya * wxb), a * (wyb * wxa), a * (wyb * wxb));
    off = (level_start + yb * W + xb) * rowstride;
}

// slot -> (n, q, m); returns false for slots outside the problem
__device__ __forceinline__ bool slot_coords(bool tiled, const LevelTable& lt, const PatchTable& pt, int L, int M, int Lq,
                                            int64_t slots, int sl, int& n, int& q, int& m) {
    if (tiled) {
        const int per_image = pt.first[L] * M;
        n = blockIdx.x / per_image;
        int r = blockIdx.x - n * per_image;
        const int patch = r / M;
        m = r - patch * M;                                   // heads innermost: the 8 heads of a pixel share DRAM pages
        int l = 0;
        while (l + 1 < L && patch >= pt.first[l + 1]) ++l;
        const int pl = patch - pt.first[l];
        const int py = pl / pt.px[l], pxi = pl - py * pt.px[l];
        const int x = pxi * kPatchW + (sl & (kPatchW - 1)), y = py * kPatchH + (sl / kPatchW);
        q = lt.start[l] + y * lt.w[l] + x;
        return x < lt.w[l] && y < lt.h[l];
    }
    const int64_t g = (int64_t)blockIdx.x * kTileSlots + sl;
    m = (int)(g % M);
    q = (int)((g / M) % Lq);
    n = (int)(g / M / Lq);
    return g < slots;
}

template <int P>
__global__ void __launch_bounds__(kTileSlots * 8)
msda_fwd_tiled(const float* __restrict__ value, const __grid_constant__ LevelTable lt, const __grid_constant__ PatchTable pt,
               const float* __restrict__ loc, const float* __restrict__ attn, float* __restrict__ out, int64_t slots,
               int S, int M, int Lq, int L, int tiled) {
    extern __shared__ __align__(16) uint8_t smem_msda[];
    const int LP = L * P;
    const int LPS = LP + 1;              // padded per-slot stride: the 4 slots of a warp read distinct banks
    float4* s_w = reinterpret_cast<float4*>(smem_msda);                        // [slot][LPS]
    int* s_off = reinterpret_cast<int*>(smem_msda + sizeof(float4) * kTileSlots * LPS);   // [slot][LPS]

    const int sl = threadIdx.x >> 3;     // slot within CTA (4 slots per warp: set-up and use stay inside the warp)
    const int j = threadIdx.x & 7;       // 4-channel group, and the lane that sets up taps j, j + 8, ...
    int n, q, m;
    const bool ok = slot_coords(tiled != 0, lt, pt, L, M, Lq, slots, sl, n, q, m);
    const int rowstride = M * 32;
    const int64_t g = ((int64_t)n * Lq + q) * M + m;
    if (ok) {
        const float2* lp = reinterpret_cast<const float2*>(loc) + g * LP;
        const float* ap = attn + g * LP;
        for (int t = j; t < LP; t += 8) {
            const int l = t / P;
            const float2 lxy = __ldg(lp + t);
            float4 w;
            int off;
            tap_record(lxy.x, lxy.y, __ldg(ap + t), lt.h[l], lt.w[l], lt.start[l], rowstride, w, off);
            s_w[sl * LPS + t] = w;
            s_off[sl * LPS + t] = off;
        }
    }
    __syncwarp();
    if (!ok) return;
    const float* vb = value + ((int64_t)n * S * M + m) * 32 + j * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        const int down = lt.w[l] * rowstride;
        float4 w[P];
        const float* p0[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            w[p] = s_w[sl * LPS + l * P + p];
            p0[p] = vb + s_off[sl * LPS + l * P + p];
        }
        float4 v[P][4];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            v[p][0] = __ldg(reinterpret_cast<const float4*>(p0[p]));
            v[p][1] = __ldg(reinterpret_cast<const float4*>(p0[p] + rowstride));
            v[p][2] = __ldg(reinterpret_cast<const float4*>(p0[p] + down));
            v[p][3] = __ldg(reinterpret_cast<const float4*>(p0[p] + down + rowstride));
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float wc[4] = {w[p].x, w[p].y, w[p].z, w[p].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc.x = fmaf(wc[c], v[p][c].x, acc.x);
                acc.y = fmaf(wc[c], v[p][c].y, acc.y);
                acc.z = fmaf(wc[c], v[p][c].z, acc.z);
                acc.w = fmaf(wc[c], v[p][c].w, acc.w);
            }
        }
    }
    *reinterpret_cast<float4*>(out + g * 32 + j * 4) = acc;
}


// ------------------------------------------------------------------------------------------------
// tiled backward (same slot mapping and per-tap set-up distribution as msda_fwd_tiled)
//
// Per tap the set-up lane publishes, for each of the 4 clamped positions p, the value weight w_p (without the
// attention weight) and the coefficients of d(out)/d(loc.x), d(out)/d(loc.y) in dot(value_p, grad_out):
//     grad_attn = sum_p w_p   * dot_p        grad_loc.x = sum_p gxw_p * dot_p       grad_loc.y = sum_p gyw_p * dot_p
//     grad_value[p] += (attn * w_p) * grad_out          (red.global.add.v4.f32, skipped when the weight is 0)
// Each of the 8 lanes of a slot holds 4 channels: it forms its partial of the three sums (12 FFMA) and the three
// values are butterfly-reduced over the 8 lanes; lane (tap % 8) keeps the totals and stores them at the end.
// ------------------------------------------------------------------------------------------------
struct BwdRec {
    float4 w, gxw, gyw;
    float a;
    int off;
    int pad0, pad1;
};     // 64 bytes

__device__ __forceinline__ void tap_record_bwd(float lx, float ly, float a, int H, int W, int level_start, int rowstride,
                                               BwdRec& r) {
    float gx = __fsub_rn(__fmul_rn(2.f, lx), 1.f);
    float gy = __fsub_rn(__fmul_rn(2.f, ly), 1.f);
    float x = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
    float y = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
    x = fminf(fmaxf(x, -2.f), (float)W + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)H + 1.f);
    const float x0f = floorf(x), y0f = floorf(y);
    const float wx1 = x - x0f, wy1 = y - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int xb = min(max(x0, 0), W - 2), yb = min(max(y0, 0), H - 2);
    // position xb (+0 / +1) carries original corner x0 (weight wx0, d/dx = -1) or x0 + 1 (weight wx1, d/dx = +1)
    const float wxa = (x0 == xb ? wx0 : 0.f) + (x0 + 1 == xb ? wx1 : 0.f);
    const float wxb = (x0 == xb + 1 ? wx0 : 0.f) + (x0 == xb ? wx1 : 0.f);
    const float dxa = (x0 == xb ? -1.f : 0.f) + (x0 + 1 == xb ? 1.f : 0.f);
    const float dxb = (x0 == xb + 1 ? -1.f : 0.f) + (x0 == xb ? 1.f : 0.f);
    const float wya = (y0 == yb ? wy0 : 0.f) + (y0 + 1 == yb ? wy1 : 0.f);
    const float wyb = (y0 == yb + 1 ? wy0 : 0.f) + (y0 == yb ? wy1 : 0.f);
    const float dya = (y0 == yb ? -1.f : 0.f) + (y0 + 1 == yb ? 1.f : 0.f);
    const float dyb = (y0 == yb + 1 ? -1.f : 0.f) + (y0 == yb ? 1.f : 0.f);
    r.w = make_float4(wya * wxa, wya * wxb, wyb * wxa, wyb * wxb);
    const float aw = a * (float)W, ah = a * (float)H;
    r.gxw = make_float4(aw * (wya * dxa), aw * (wya * dxb), aw * (wyb * dxa), aw * (wyb * dxb));
    r.gyw = make_float4(ah * (dya * wxa), ah * (dya * wxb), ah * (dyb * wxa), ah * (dyb * wxb));
    r.a = a;
    r.off = (level_start + yb * W + xb) * rowstride;
    r.pad0 = r.pad1 = 0;
}

template <int P>
__global__ void __launch_bounds__(kTileSlots * 8)
msda_bwd_tiled(const float* __restrict__ value, const __grid_constant__ LevelTable lt, const __grid_constant__ PatchTable pt,
               const float* __restrict__ loc, const float* __restrict__ attn, const float* __restrict__ grad_out,
               float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn, int64_t slots,
               int S, int M, int Lq, int L, int tiled) {
    extern __shared__ __align__(16) uint8_t smem_msda[];
    const int LP = L * P;
    const int LPS = LP * 4 + 1;          // float4 units per slot, padded (bank spread across the 4 slots of a warp)
    float4* s_rec = reinterpret_cast<float4*>(smem_msda);

    const int sl = threadIdx.x >> 3;
    const int j = threadIdx.x & 7;
    int n, q, m;
    const bool ok = slot_coords(tiled != 0, lt, pt, L, M, Lq, slots, sl, n, q, m);
    const int rowstride = M * 32;
    const int64_t g = ((int64_t)n * Lq + q) * M + m;
    if (ok) {
        const float2* lp = reinterpret_cast<const float2*>(loc) + g * LP;
        const float* ap = attn + g * LP;
        for (int t = j; t < LP; t += 8) {
            const int l = t / P;
            const float2 lxy = __ldg(lp + t);
            BwdRec r;
            tap_record_bwd(lxy.x, lxy.y, __ldg(ap + t), lt.h[l], lt.w[l], lt.start[l], rowstride, r);
            float4* dst = s_rec + sl * LPS + t * 4;
            dst[0] = r.w;
            dst[1] = r.gxw;
            dst[2] = r.gyw;
            dst[3] = make_float4(r.a, __int_as_float(r.off), 0.f, 0.f);
        }
    }
    __syncwarp();
    if (!ok) return;
    const int64_t plane = ((int64_t)n * S * M + m) * 32 + j * 4;
    const float* vb = value + plane;
    float* gvb = grad_value + plane;
    const float4 go = __ldg(reinterpret_cast<const float4*>(grad_out + g * 32 + j * 4));
    float ka0 = 0.f, kx0 = 0.f, ky0 = 0.f, ka1 = 0.f, kx1 = 0.f, ky1 = 0.f;          // totals of taps j and j + 8 (LP <= 16)
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        const int down = lt.w[l] * rowstride;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int t = l * P + p;
            const float4* rec = s_rec + sl * LPS + t * 4;
            const float4 w = rec[0], gxw = rec[1], gyw = rec[2], ao = rec[3];
            const int off = __float_as_int(ao.y);
            const float* p0 = vb + off;
            float4 v[4];
            v[0] = __ldg(reinterpret_cast<const float4*>(p0));
            v[1] = __ldg(reinterpret_cast<const float4*>(p0 + rowstride));
            v[2] = __ldg(reinterpret_cast<const float4*>(p0 + down));
            v[3] = __ldg(reinterpret_cast<const float4*>(p0 + down + rowstride));
            const float wc[4] = {w.x, w.y, w.z, w.w};
            const float xc[4] = {gxw.x, gxw.y, gxw.z, gxw.w};
            const float yc[4] = {gyw.x, gyw.y, gyw.z, gyw.w};
            const int poff[4] = {0, rowstride, down, down + rowstride};
            float pa = 0.f, px = 0.f, py = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float d = v[c].x * go.x + v[c].y * go.y + v[c].z * go.z + v[c].w * go.w;
                pa = fmaf(wc[c], d, pa);
                px = fmaf(xc[c], d, px);
                py = fmaf(yc[c], d, py);
                const float cw = ao.x * wc[c];
                if (cw != 0.f) red_add_v4(gvb + off + poff[c], cw * go.x, cw * go.y, cw * go.z, cw * go.w);
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                pa += __shfl_xor_sync(0xffffffffu, pa, o);
                px += __shfl_xor_sync(0xffffffffu, px, o);
                py += __shfl_xor_sync(0xffffffffu, py, o);
            }
            if (t == j) { ka0 = pa; kx0 = px; ky0 = py; }
            if (t == j + 8) { ka1 = pa; kx1 = px; ky1 = py; }
        }
    }
    if (j < LP) {
        grad_attn[g * LP + j] = ka0;
        *reinterpret_cast<float2*>(grad_loc + (g * LP + j) * 2) = make_float2(kx0, ky0);
    }
    if (j + 8 < LP) {
        grad_attn[g * LP + j + 8] = ka1;
        *reinterpret_cast<float2*>(grad_loc + (g * LP + j + 8) * 2) = make_float2(kx1, ky1);
    }
}

// host: patch grid of the tiled mapping; returns the number of CTAs
static int64_t make_patches(const LevelTable& lt, int L, int N, int M, PatchTable& pt) {
    int total = 0;
    for (int l = 0; l < L; ++l) {
        pt.first[l] = total;
        pt.px[l] = (lt.w[l] + kPatchW - 1) / kPatchW;
        total += pt.px[l] * ((lt.h[l] + kPatchH - 1) / kPatchH);
    }
    pt.first[L] = total;
    return (int64_t)total * M * N;
}

static bool levels_at_least_2x2(const LevelTable& lt, int L) {
    for (int l = 0; l < L; ++l)
        if (lt.h[l] < 2 || lt.w[l] < 2) return false;
    return true;
}

static int make_levels(const int64_t* shapes_hw, const int64_t* level_start, int L, int S, LevelTable& lt) {
    PDB_REQUIRE(L >= 1 && L <= kMaxLevels, "msda: L=%d outside [1,%d]", L, kMaxLevels);
    for (int l = 0; l < L; ++l) {
        int64_t h = shapes_hw[2 * l], w = shapes_hw[2 * l + 1], st = level_start[l];
        PDB_REQUIRE(h > 0 && w > 0 && st >= 0 && st + h * w <= S, "msda: level %d (%lld x %lld @ %lld) exceeds S=%d",
                    l, (long long)h, (long long)w, (long long)st, S);
        lt.h[l] = (int)h; lt.w[l] = (int)w; lt.start[l] = (int)st;
    }
    return PDB_OK;
}

template <typename T>
static int fwd_generic(const void* value, const LevelTable& lt, const void* loc, const void* attn, void* out,
                       int N, int S, int M, int D, int Lq, int L, int P, cudaStream_t st) {
    int64_t total = (int64_t)N * Lq * M * D;
    int threads = 256;
    int64_t blocks = (total + threads - 1) / threads;
    msda_fwd_generic<T><<<(unsigned)blocks, threads, 0, st>>>((const T*)value, lt, (const T*)loc, (const T*)attn,
                                                               (T*)out, total, S, M, D, Lq, L, P);
    return launched("msda_fwd_generic");
}

template <typename T>
static int bwd_generic(const void* value, const LevelTable& lt, const void* loc, const void* attn, const void* go,
                       void* gv, void* gl, void* ga, int N, int S, int M, int D, int Lq, int L, int P,
                       cudaStream_t st) {
    int64_t total = (int64_t)N * Lq * M * D;
    cudaMemsetAsync(gl, 0, sizeof(T) * (size_t)N * Lq * M * L * P * 2, st);
    cudaMemsetAsync(ga, 0, sizeof(T) * (size_t)N * Lq * M * L * P, st);
    int threads = 256;
    int64_t blocks = (total + threads - 1) / threads;
    msda_bwd_generic<T><<<(unsigned)blocks, threads, 0, st>>>((const T*)value, lt, (const T*)loc, (const T*)attn,
                                                               (const T*)go, (T*)gv, (T*)gl, (T*)ga, total, S, M, D,
                                                               Lq, L, P);
    return launched("msda_bwd_generic");
}

}  // namespace pdb

// msda_tile.cu: the encoder path with TMA-staged value tiles
namespace pdb {
bool msda_tma_eligible(const LevelTable& lt, int L, int S, int M, int D, int Lq, int P);
int msda_forward_tma_f32(const void* value, const LevelTable& lt, const float* loc, const float* attn, float* out, int N, int S,
                         int M, int L, cudaStream_t st);
int g_tma_variant_set(int v);
static int g_msda_path = 0;     // pdb_debug_set_msda_path: 0 = measured dispatch (below), 1 = L1-resident tiled kernels only, 4 = TMA tiles wherever eligible
}

using namespace pdb;

extern "C" int pdb_debug_set_msda_path(int path) {
    g_msda_path = path & 5;
    g_tma_variant_set((path >> 1) & 1);    // bit 1: experimental launch shape of the fp16-staged kernel
    return PDB_OK;
}

extern "C" int pdb_msda_forward(const void* value, const int64_t* shapes_hw, const int64_t* level_start,
                                const void* loc, const void* attn, void* out, int N, int S, int M, int D, int Lq,
                                int L, int P, int dtype, void* stream) {
    PDB_REQUIRE(value && loc && attn && out && shapes_hw && level_start, "msda_forward: null pointer");
    PDB_REQUIRE(N > 0 && S > 0 && M > 0 && D > 0 && Lq > 0 && P > 0, "msda_forward: non-positive dimension");
    LevelTable lt;
    PDB_TRY(make_levels(shapes_hw, level_start, L, S, lt));
    cudaStream_t st = as_stream(stream);
    if (dtype == PDB_F64) return fwd_generic<double>(value, lt, loc, attn, out, N, S, M, D, Lq, L, P, st);
    PDB_REQUIRE(dtype == PDB_F32, "msda_forward: dtype %d (only f32/f64, as ms_deform_attn_cuda.cu:70)", dtype);
    // Measured dispatch (profiles/r02_bench_msda_v*.json): the persistent TMA-staged kernel wins on the 4-level pyramid
    // (C5(i): 298 vs 322 us) and ties on 3 levels (C2: 126 vs 124 us), where the L1-resident kernel stays.
    if ((g_msda_path == 4 || (g_msda_path == 0 && L >= 4)) && msda_tma_eligible(lt, L, S, M, D, Lq, P))
        return msda_forward_tma_f32(value, lt, (const float*)loc, (const float*)attn, (float*)out, N, S, M, L, st);
    if (D == 32 && P == 4 && (int64_t)S * M * 32 < (1ll << 31) && levels_at_least_2x2(lt, L)) {
        int64_t slots = (int64_t)N * Lq * M;
        PatchTable pt;
        const bool tiled = Lq == S;      // queries are the pyramid's own pixels (encoder self-attention): 2-D patches
        int64_t blocks = make_patches(lt, L, N, M, pt);
        if (!tiled) blocks = (slots + kTileSlots - 1) / kTileSlots;
        PDB_REQUIRE(blocks < (1ll << 31), "msda_forward: too many CTAs");
        size_t smem = (sizeof(float4) + sizeof(int)) * kTileSlots * (L * P + 1);
        msda_fwd_tiled<4><<<(unsigned)blocks, kTileSlots * 8, smem, st>>>((const float*)value, lt, pt, (const float*)loc,
                                                                         (const float*)attn, (float*)out, slots, S, M, Lq,
                                                                         L, tiled ? 1 : 0);
        return launched("msda_fwd_tiled");
    }
    if (D == 32 && P == 4 && (int64_t)S * M * 32 < (1ll << 31)) {
        int64_t slots = (int64_t)N * Lq * M;
        int64_t blocks = (slots + kSlotsPerCta - 1) / kSlotsPerCta;
        size_t smem = sizeof(float) * kSlotsPerCta * ((L * P * 2 + 2) + (L * P + 1));
        msda_fwd_d32<4><<<(unsigned)blocks, kFastThreads, smem, st>>>((const float*)value, lt, (const float*)loc,
                                                                       (const float*)attn, (float*)out, slots, S, M,
                                                                       Lq, L);
        return launched("msda_fwd_d32");
    }
    return fwd_generic<float>(value, lt, loc, attn, out, N, S, M, D, Lq, L, P, st);
}

extern "C" int pdb_msda_backward(const void* value, const int64_t* shapes_hw, const int64_t* level_start,
                                 const void* loc, const void* attn, const void* grad_out, void* grad_value,
                                 void* grad_loc, void* grad_attn, int N, int S, int M, int D, int Lq, int L, int P,
                                 int dtype, void* stream) {
    PDB_REQUIRE(value && loc && attn && grad_out && grad_value && grad_loc && grad_attn && shapes_hw && level_start,
                "msda_backward: null pointer");
    PDB_REQUIRE(N > 0 && S > 0 && M > 0 && D > 0 && Lq > 0 && P > 0, "msda_backward: non-positive dimension");
    LevelTable lt;
    PDB_TRY(make_levels(shapes_hw, level_start, L, S, lt));
    cudaStream_t st = as_stream(stream);
    size_t esz = dtype == PDB_F64 ? 8 : 4;
    PDB_REQUIRE(dtype == PDB_F32 || dtype == PDB_F64, "msda_backward: dtype %d", dtype);
    cudaMemsetAsync(grad_value, 0, esz * (size_t)N * S * M * D, st);
    if (dtype == PDB_F64)
        return bwd_generic<double>(value, lt, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, S, M, D, Lq, L,
                                   P, st);
    if (D == 32 && P == 4 && L * P <= 16 && (int64_t)S * M * 32 < (1ll << 31) && levels_at_least_2x2(lt, L)) {
        int64_t slots = (int64_t)N * Lq * M;
        PatchTable pt;
        const bool tiled = Lq == S;
        int64_t blocks = make_patches(lt, L, N, M, pt);
        if (!tiled) blocks = (slots + kTileSlots - 1) / kTileSlots;
        PDB_REQUIRE(blocks < (1ll << 31), "msda_backward: too many CTAs");
        size_t smem = sizeof(float4) * kTileSlots * (L * P * 4 + 1);
        msda_bwd_tiled<4><<<(unsigned)blocks, kTileSlots * 8, smem, st>>>(
            (const float*)value, lt, pt, (const float*)loc, (const float*)attn, (const float*)grad_out,
            (float*)grad_value, (float*)grad_loc, (float*)grad_attn, slots, S, M, Lq, L, tiled ? 1 : 0);
        return launched("msda_bwd_tiled");
    }
    if (D == 32 && P == 4 && (int64_t)S * M * 32 < (1ll << 31)) {
        int64_t slots = (int64_t)N * Lq * M;
        int64_t blocks = (slots + kSlotsPerCta - 1) / kSlotsPerCta;
        int LP = L * P;
        size_t smem = sizeof(float) * kSlotsPerCta * ((LP * 2 + 2) + (LP + 1) + LP * 2 + LP);
        msda_bwd_d32<4><<<(unsigned)blocks, kFastThreads, smem, st>>>(
            (const float*)value, lt, (const float*)loc, (const float*)attn, (const float*)grad_out,
            (float*)grad_value, (float*)grad_loc, (float*)grad_attn, slots, S, M, Lq, L);
        return launched("msda_bwd_d32");
    }
    return bwd_generic<float>(value, lt, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, S, M, D, Lq, L, P,
                              st);
}
