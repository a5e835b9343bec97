// Masked cross-attention core: softmax(q k^T + mask) v per head, fused (scores never reach HBM).
// Reference: nn.MultiheadAttention inside CrossAttentionLayer.forward_post
// (mask2former_transformer_decoder.py:84,102-114) with the bool attn_mask built at :453-457; the
// reference materialises (B*heads, Q, Lk) fp32 scores and head-averaged weights.
//
// Layout in HBM: q (B, Q, heads*32) pre-scaled by 1/sqrt(32); k, v (B, Lk, heads*32): one
// (key, head) is one 128-byte line; mask (B, Q, Lk) uint8 shared by all heads (1 = not attended);
// row_any (B*Q): 0 => the row has no attended key => attend everywhere (the :405 reset).
//
// Forward: flash-style split over keys.  One thread owns one query (its 32-float q, accumulator and
// online-softmax state live in registers); K/V tiles of 64 keys are staged in shared memory and
// read as warp broadcasts; masked keys are skipped.  Partial (max, sum, acc) per key split are
// merged by a small combine kernel that also emits the log-sum-exp for backward.
// Backward: one pass with a thread per query (dq) and one with a thread per key (dk, dv).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace pdb {

constexpr int XD = 32;          // head dim
constexpr int XTK = 64;         // keys per shared-memory tile
constexpr int XTHREADS = 128;   // queries per CTA

static int xattn_nsplit(int B, int heads, int Q, int Lk) {
    int qtiles = (Q + XTHREADS - 1) / XTHREADS;
    int tiles = (Lk + XTK - 1) / XTK;
    int want = (2 * kNumSMs + B * heads * qtiles - 1) / (B * heads * qtiles);
    if (want < 1) want = 1;
    if (want > tiles) want = tiles;
    int tiles_per = (tiles + want - 1) / want;
    return (tiles + tiles_per - 1) / tiles_per;
}

__device__ __forceinline__ void load_kv_tile(float (*sK)[XD], float (*sV)[XD], const float* __restrict__ k,
                                             const float* __restrict__ v, int64_t rowbase, int j0, int Lk, int ld,
                                             int hoff) {
    // 64 keys x 8 float4 per matrix
    for (int e = threadIdx.x; e < XTK * (XD / 4); e += XTHREADS) {
        int jj = e >> 3, c = e & 7;
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (j0 + jj < Lk) {
            int64_t off = (rowbase + j0 + jj) * ld + hoff + c * 4;
            kv = __ldg(reinterpret_cast<const float4*>(k + off));
            vv = __ldg(reinterpret_cast<const float4*>(v + off));
        }
        *reinterpret_cast<float4*>(&sK[jj][c * 4]) = kv;
        *reinterpret_cast<float4*>(&sV[jj][c * 4]) = vv;
    }
}

__device__ __forceinline__ float dot32(const float* q, const float* krow) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < XD / 4; ++c) {
        float4 kk = *reinterpret_cast<const float4*>(krow + c * 4);
        s = fmaf(q[c * 4 + 0], kk.x, s);
        s = fmaf(q[c * 4 + 1], kk.y, s);
        s = fmaf(q[c * 4 + 2], kk.z, s);
        s = fmaf(q[c * 4 + 3], kk.w, s);
    }
    return s;
}

// 8 mask bytes for keys [j, j+8) of one query row -> bit i set = key j+i is masked
__device__ __forceinline__ unsigned mask_bits8(const uint8_t* __restrict__ mrow, int j, int Lk, bool use_mask,
                                               bool vec_ok) {
    unsigned bits = 0;
    if (use_mask) {
        if (vec_ok && j + 8 <= Lk) {
            uint2 mm = __ldg(reinterpret_cast<const uint2*>(mrow + j));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if ((mm.x >> (8 * i)) & 0xffu) bits |= 1u << i;
                if ((mm.y >> (8 * i)) & 0xffu) bits |= 1u << (4 + i);
            }
        } else {
            for (int i = 0; i < 8; ++i)
                if (j + i < Lk && __ldg(mrow + j + i)) bits |= 1u << i;
        }
    }
    for (int i = 0; i < 8; ++i)
        if (j + i >= Lk) bits |= 1u << i;
    return bits;
}

__global__ void __launch_bounds__(XTHREADS)
xattn_fwd_partial(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                  const uint8_t* __restrict__ mask, const int32_t* __restrict__ row_any, float* __restrict__ ws_acc,
                  float* __restrict__ ws_ml, int heads, int Q, int Lk, int nsplit, int tiles_per, int qtiles) {
    __shared__ __align__(16) float sK[XTK][XD];
    __shared__ __align__(16) float sV[XTK][XD];
    const int split = blockIdx.x % nsplit, qt = blockIdx.x / nsplit;
    const int h = blockIdx.y, b = blockIdx.z;
    const int ld = heads * XD, hoff = h * XD;
    const int qi = qt * XTHREADS + threadIdx.x;
    const bool active = qi < Q;
    float qr[XD], acc[XD];
    float m = -INFINITY, l = 0.f;
#pragma unroll
    for (int d = 0; d < XD; ++d) { qr[d] = 0.f; acc[d] = 0.f; }
    bool use_mask = false;
    const uint8_t* mrow = nullptr;
    if (active) {
        const float* qp = q + ((int64_t)b * Q + qi) * ld + hoff;
#pragma unroll
        for (int c = 0; c < XD / 4; ++c) {
            float4 t = __ldg(reinterpret_cast<const float4*>(qp + c * 4));
            qr[c * 4] = t.x; qr[c * 4 + 1] = t.y; qr[c * 4 + 2] = t.z; qr[c * 4 + 3] = t.w;
        }
        use_mask = mask != nullptr && (row_any == nullptr || row_any[(int64_t)b * Q + qi] != 0);
        if (mask) mrow = mask + ((int64_t)b * Q + qi) * Lk;
    }
    const bool vec_ok = (Lk % 8) == 0;
    const int tile_begin = split * tiles_per;
    const int tile_end = min(tile_begin + tiles_per, (Lk + XTK - 1) / XTK);
    for (int t = tile_begin; t < tile_end; ++t) {
        const int j0 = t * XTK;
        __syncthreads();
        load_kv_tile(sK, sV, k, v, (int64_t)b * Lk, j0, Lk, ld, hoff);
        __syncthreads();
        if (!active) continue;
#pragma unroll 1
        for (int g8 = 0; g8 < XTK; g8 += 8) {
            if (j0 + g8 >= Lk) break;
            unsigned bits = mask_bits8(mrow, j0 + g8, Lk, use_mask, vec_ok);
            if (bits == 0xffu) continue;
            float s[8];
            float gmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] = -INFINITY;
                if (!((bits >> i) & 1u)) {
                    s[i] = dot32(qr, sK[g8 + i]);
                    gmax = fmaxf(gmax, s[i]);
                }
            }
            float mn = fmaxf(m, gmax);
            float sc = (m == -INFINITY) ? 0.f : expf(m - mn);
            l *= sc;
#pragma unroll
            for (int d = 0; d < XD; ++d) acc[d] *= sc;
            m = mn;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((bits >> i) & 1u) continue;
                float p = expf(s[i] - mn);
                l += p;
                const float* vr = sV[g8 + i];
#pragma unroll
                for (int c = 0; c < XD / 4; ++c) {
                    float4 vv = *reinterpret_cast<const float4*>(vr + c * 4);
                    acc[c * 4 + 0] = fmaf(p, vv.x, acc[c * 4 + 0]);
                    acc[c * 4 + 1] = fmaf(p, vv.y, acc[c * 4 + 1]);
                    acc[c * 4 + 2] = fmaf(p, vv.z, acc[c * 4 + 2]);
                    acc[c * 4 + 3] = fmaf(p, vv.w, acc[c * 4 + 3]);
                }
            }
        }
    }
    if (active) {
        int64_t slot = (((int64_t)b * heads + h) * nsplit + split) * Q + qi;
        float* oa = ws_acc + slot * XD;
#pragma unroll
        for (int c = 0; c < XD / 4; ++c)
            *reinterpret_cast<float4*>(oa + c * 4) = make_float4(acc[c * 4], acc[c * 4 + 1], acc[c * 4 + 2], acc[c * 4 + 3]);
        ws_ml[slot * 2] = m;
        ws_ml[slot * 2 + 1] = l;
    }
}

// one warp per (b, h, q): lane = channel
__global__ void xattn_fwd_combine(const float* __restrict__ ws_acc, const float* __restrict__ ws_ml,
                                  float* __restrict__ out, float* __restrict__ lse, int B, int heads, int Q,
                                  int nsplit) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= (int64_t)B * heads * Q) return;
    int qi = (int)(w % Q);
    int64_t bh = w / Q;
    int h = (int)(bh % heads);
    int64_t b = bh / heads;
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, ws_ml[((bh * nsplit + s) * Q + qi) * 2]);
    float L = 0.f, o = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        int64_t slot = (bh * nsplit + s) * Q + qi;
        float ms = ws_ml[slot * 2];
        if (ms == -INFINITY) continue;
        float sc = expf(ms - M);
        L += ws_ml[slot * 2 + 1] * sc;
        o = fmaf(ws_acc[slot * XD + lane], sc, o);
    }
    out[((int64_t)b * Q + qi) * heads * XD + h * XD + lane] = o / L;
    if (lane == 0) lse[bh * Q + qi] = M + logf(L);
}

// grad_q: thread per query, split over keys, atomics into the zero-filled grad_q
__global__ void __launch_bounds__(XTHREADS)
xattn_bwd_dq(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
             const uint8_t* __restrict__ mask, const int32_t* __restrict__ row_any, const float* __restrict__ out,
             const float* __restrict__ lse, const float* __restrict__ gout, float* __restrict__ gq, int heads, int Q,
             int Lk, int nsplit, int tiles_per) {
    __shared__ __align__(16) float sK[XTK][XD];
    __shared__ __align__(16) float sV[XTK][XD];
    const int split = blockIdx.x % nsplit, qt = blockIdx.x / nsplit;
    const int h = blockIdx.y, b = blockIdx.z;
    const int ld = heads * XD, hoff = h * XD;
    const int qi = qt * XTHREADS + threadIdx.x;
    const bool active = qi < Q;
    float qr[XD], dor[XD], dq[XD];
    float delta = 0.f, lsei = 0.f;
#pragma unroll
    for (int d = 0; d < XD; ++d) { qr[d] = 0.f; dor[d] = 0.f; dq[d] = 0.f; }
    bool use_mask = false;
    const uint8_t* mrow = nullptr;
    if (active) {
        int64_t ro = ((int64_t)b * Q + qi) * ld + hoff;
#pragma unroll
        for (int d = 0; d < XD; ++d) {
            qr[d] = __ldg(q + ro + d);
            dor[d] = __ldg(gout + ro + d);
            delta = fmaf(dor[d], __ldg(out + ro + d), delta);
        }
        lsei = lse[((int64_t)b * heads + h) * Q + qi];
        use_mask = mask != nullptr && (row_any == nullptr || row_any[(int64_t)b * Q + qi] != 0);
        if (mask) mrow = mask + ((int64_t)b * Q + qi) * Lk;
    }
    const bool vec_ok = (Lk % 8) == 0;
    const int tile_begin = split * tiles_per;
    const int tile_end = min(tile_begin + tiles_per, (Lk + XTK - 1) / XTK);
    for (int t = tile_begin; t < tile_end; ++t) {
        const int j0 = t * XTK;
        __syncthreads();
        load_kv_tile(sK, sV, k, v, (int64_t)b * Lk, j0, Lk, ld, hoff);
        __syncthreads();
        if (!active) continue;
#pragma unroll 1
        for (int g8 = 0; g8 < XTK; g8 += 8) {
            if (j0 + g8 >= Lk) break;
            unsigned bits = mask_bits8(mrow, j0 + g8, Lk, use_mask, vec_ok);
            if (bits == 0xffu) continue;
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                if ((bits >> i) & 1u) continue;
                float s = dot32(qr, sK[g8 + i]);
                float p = expf(s - lsei);
                float dp = dot32(dor, sV[g8 + i]);
                float ds = p * (dp - delta);
                const float* kr = sK[g8 + i];
#pragma unroll
                for (int d = 0; d < XD; ++d) dq[d] = fmaf(ds, kr[d], dq[d]);
            }
        }
    }
    if (active) {
        float* o = gq + ((int64_t)b * Q + qi) * ld + hoff;
#pragma unroll
        for (int d = 0; d < XD; ++d) atomicAdd(o + d, dq[d]);
    }
}

// grad_k / grad_v: thread per key; all queries of the (b, h) staged in shared memory
__global__ void __launch_bounds__(XTHREADS)
xattn_bwd_dkv(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
              const uint8_t* __restrict__ mask, const int32_t* __restrict__ row_any, const float* __restrict__ out,
              const float* __restrict__ lse, const float* __restrict__ gout, float* __restrict__ gk,
              float* __restrict__ gv, int heads, int Q, int Lk) {
    constexpr int QT = 32;   // queries staged per round
    __shared__ __align__(16) float sQ[QT][XD];
    __shared__ __align__(16) float sDO[QT][XD];
    __shared__ float sLse[QT], sDelta[QT];
    __shared__ int sUse[QT];
    const int h = blockIdx.y, b = blockIdx.z;
    const int ld = heads * XD, hoff = h * XD;
    const int j = blockIdx.x * XTHREADS + threadIdx.x;
    const bool active = j < Lk;
    float kr[XD], vr[XD], dk[XD], dv[XD];
#pragma unroll
    for (int d = 0; d < XD; ++d) { kr[d] = 0.f; vr[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
    if (active) {
        int64_t off = ((int64_t)b * Lk + j) * ld + hoff;
#pragma unroll
        for (int c = 0; c < XD / 4; ++c) {
            float4 a = __ldg(reinterpret_cast<const float4*>(k + off + c * 4));
            float4 bb = __ldg(reinterpret_cast<const float4*>(v + off + c * 4));
            kr[c * 4] = a.x; kr[c * 4 + 1] = a.y; kr[c * 4 + 2] = a.z; kr[c * 4 + 3] = a.w;
            vr[c * 4] = bb.x; vr[c * 4 + 1] = bb.y; vr[c * 4 + 2] = bb.z; vr[c * 4 + 3] = bb.w;
        }
    }
    for (int i0 = 0; i0 < Q; i0 += QT) {
        __syncthreads();
        for (int e = threadIdx.x; e < QT * XD; e += XTHREADS) {
            int ii = e >> 5, d = e & 31;
            float qv = 0.f, dov = 0.f;
            if (i0 + ii < Q) {
                int64_t ro = ((int64_t)b * Q + i0 + ii) * ld + hoff + d;
                qv = __ldg(q + ro);
                dov = __ldg(gout + ro);
            }
            sQ[ii][d] = qv;
            sDO[ii][d] = dov;
        }
        {   // one warp per staged query computes delta = dO . O
            int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int ii = wid; ii < QT; ii += XTHREADS / 32) {
                float pr = 0.f;
                if (i0 + ii < Q) {
                    int64_t ro = ((int64_t)b * Q + i0 + ii) * ld + hoff + lane;
                    pr = __ldg(gout + ro) * __ldg(out + ro);
                }
                pr = warp_sum(pr);
                if (lane == 0) {
                    sDelta[ii] = pr;
                    sLse[ii] = (i0 + ii < Q) ? lse[((int64_t)b * heads + h) * Q + i0 + ii] : 0.f;
                    sUse[ii] = (i0 + ii < Q) ? ((mask != nullptr && (row_any == nullptr || row_any[(int64_t)b * Q + i0 + ii] != 0)) ? 1 : 0) : -1;
                }
            }
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 1
        for (int ii = 0; ii < QT; ++ii) {
            int use = sUse[ii];
            if (use < 0) break;
            if (use == 1 && __ldg(mask + ((int64_t)b * Q + i0 + ii) * Lk + j)) continue;
            float s = dot32(kr, sQ[ii]);
            float p = expf(s - sLse[ii]);
            float dp = dot32(vr, sDO[ii]);
            float ds = p * (dp - sDelta[ii]);
#pragma unroll
            for (int c = 0; c < XD / 4; ++c) {
                float4 dd = *reinterpret_cast<const float4*>(&sDO[ii][c * 4]);
                float4 qq = *reinterpret_cast<const float4*>(&sQ[ii][c * 4]);
                dv[c * 4 + 0] = fmaf(p, dd.x, dv[c * 4 + 0]);
                dv[c * 4 + 1] = fmaf(p, dd.y, dv[c * 4 + 1]);
                dv[c * 4 + 2] = fmaf(p, dd.z, dv[c * 4 + 2]);
                dv[c * 4 + 3] = fmaf(p, dd.w, dv[c * 4 + 3]);
                dk[c * 4 + 0] = fmaf(ds, qq.x, dk[c * 4 + 0]);
                dk[c * 4 + 1] = fmaf(ds, qq.y, dk[c * 4 + 1]);
                dk[c * 4 + 2] = fmaf(ds, qq.z, dk[c * 4 + 2]);
                dk[c * 4 + 3] = fmaf(ds, qq.w, dk[c * 4 + 3]);
            }
        }
    }
    if (active) {
        int64_t off = ((int64_t)b * Lk + j) * ld + hoff;
#pragma unroll
        for (int c = 0; c < XD / 4; ++c) {
            *reinterpret_cast<float4*>(gk + off + c * 4) = make_float4(dk[c * 4], dk[c * 4 + 1], dk[c * 4 + 2], dk[c * 4 + 3]);
            *reinterpret_cast<float4*>(gv + off + c * 4) = make_float4(dv[c * 4], dv[c * 4 + 1], dv[c * 4 + 2], dv[c * 4 + 3]);
        }
    }
}

}  // namespace pdb

// xattn_mma.cu: the same three passes on the tensor cores (mma.sync TF32, 3-pass split); PDB_XATTN=simt keeps the kernels above
namespace pdb {
int xattn_fwd_partial_mma_launch(const float* q, const float* k, const float* v, const uint8_t* mask, const int32_t* row_any,
                                 float* ws_acc, float* ws_ml, int B, int heads, int Q, int Lk, int ns, int tiles_per, int qtiles,
                                 int one_pass, cudaStream_t st);
int xattn_bwd_mma_launch(const float* q, const float* k, const float* v, const uint8_t* mask, const int32_t* row_any, const float* out,
                         const float* lse, const float* gout, float* gq, float* gk, float* gv, int B, int heads, int Q, int Lk, int ns,
                         int tiles_per, int qtiles, int one_pass, cudaStream_t st);
// 3 = fp32-accurate 3xTF32 products (default), 1 = one TF32 product (pdb_set_xattn_passes; the bf16-autocast path)
static int g_xattn_passes = 3;
static bool xattn_use_mma() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("PDB_XATTN");
        mode = (e && e[0] == 's') ? 0 : 1;
    }
    return mode == 1;
}
}

using namespace pdb;

extern "C" int64_t pdb_masked_xattn_workspace_bytes(int B, int heads, int Q, int Lk, int d) {
    if (B <= 0 || heads <= 0 || Q <= 0 || Lk <= 0 || d != XD) return -1;
    int ns = xattn_nsplit(B, heads, Q, Lk);
    return (int64_t)B * heads * ns * Q * (XD + 2) * (int64_t)sizeof(float);
}

extern "C" int pdb_masked_xattn_forward(const float* q, const float* k, const float* v, const uint8_t* mask,
                                        const int32_t* row_any, float* out, float* lse, void* workspace, int B,
                                        int heads, int Q, int Lk, int d, void* stream) {
    PDB_REQUIRE(q && k && v && out && lse && workspace, "masked_xattn_forward: null pointer");
    PDB_REQUIRE(d == XD, "masked_xattn_forward: head dim %d (only 32)", d);
    PDB_REQUIRE(B > 0 && heads > 0 && Q > 0 && Lk > 0 && B <= 65535 && heads <= 65535, "masked_xattn_forward: bad shape");
    cudaStream_t st = as_stream(stream);
    int ns = xattn_nsplit(B, heads, Q, Lk);
    int tiles = (Lk + XTK - 1) / XTK;
    int tiles_per = (tiles + ns - 1) / ns;
    int qtiles = (Q + XTHREADS - 1) / XTHREADS;
    float* ws_acc = (float*)workspace;
    float* ws_ml = ws_acc + (int64_t)B * heads * ns * Q * XD;
    if (xattn_use_mma()) {
        PDB_TRY(xattn_fwd_partial_mma_launch(q, k, v, mask, row_any, ws_acc, ws_ml, B, heads, Q, Lk, ns, tiles_per, qtiles,
                                             g_xattn_passes == 1, st));
    } else {
        dim3 grid((unsigned)(ns * qtiles), (unsigned)heads, (unsigned)B);
        xattn_fwd_partial<<<grid, XTHREADS, 0, st>>>(q, k, v, mask, row_any, ws_acc, ws_ml, heads, Q, Lk, ns, tiles_per,
                                                     qtiles);
        PDB_TRY(launched("xattn_fwd_partial"));
    }
    int64_t warps = (int64_t)B * heads * Q;
    xattn_fwd_combine<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(ws_acc, ws_ml, out, lse, B, heads, Q, ns);
    return launched("xattn_fwd_combine");
}

extern "C" int pdb_masked_xattn_backward(const float* q, const float* k, const float* v, const uint8_t* mask,
                                         const int32_t* row_any, const float* out, const float* lse,
                                         const float* grad_out, float* grad_q, float* grad_k, float* grad_v, int B,
                                         int heads, int Q, int Lk, int d, void* stream) {
    PDB_REQUIRE(q && k && v && out && lse && grad_out && grad_q && grad_k && grad_v, "masked_xattn_backward: null pointer");
    PDB_REQUIRE(d == XD, "masked_xattn_backward: head dim %d (only 32)", d);
    PDB_REQUIRE(B > 0 && heads > 0 && Q > 0 && Lk > 0 && B <= 65535 && heads <= 65535, "masked_xattn_backward: bad shape");
    cudaStream_t st = as_stream(stream);
    int ns = xattn_nsplit(B, heads, Q, Lk);
    int tiles = (Lk + XTK - 1) / XTK;
    int tiles_per = (tiles + ns - 1) / ns;
    int qtiles = (Q + XTHREADS - 1) / XTHREADS;
    cudaMemsetAsync(grad_q, 0, sizeof(float) * (size_t)B * Q * heads * XD, st);
    if (xattn_use_mma())
        return xattn_bwd_mma_launch(q, k, v, mask, row_any, out, lse, grad_out, grad_q, grad_k, grad_v, B, heads, Q, Lk, ns, tiles_per,
                                    qtiles, g_xattn_passes == 1, st);
    dim3 grid((unsigned)(ns * qtiles), (unsigned)heads, (unsigned)B);
    xattn_bwd_dq<<<grid, XTHREADS, 0, st>>>(q, k, v, mask, row_any, out, lse, grad_out, grad_q, heads, Q, Lk, ns,
                                            tiles_per);
    PDB_TRY(launched("xattn_bwd_dq"));
    dim3 grid2((unsigned)((Lk + XTHREADS - 1) / XTHREADS), (unsigned)heads, (unsigned)B);
    xattn_bwd_dkv<<<grid2, XTHREADS, 0, st>>>(q, k, v, mask, row_any, out, lse, grad_out, grad_k, grad_v, heads, Q, Lk);
    return launched("xattn_bwd_dkv");
}

// Arithmetic of the tensor-core attention kernels for the calls that follow (process-wide; a captured CUDA graph keeps what was
// set when it was captured): 3 = 3xTF32, fp32-accurate (default); 1 = one TF32 product per MMA, for torch.autocast(bfloat16)
// regions, whose reference rounds q, k, softmax(p) and v to bf16.  Returns the previous value.
extern "C" int pdb_set_xattn_passes(int passes) {
    const int prev = pdb::g_xattn_passes;
    if (passes == 1 || passes == 3) pdb::g_xattn_passes = passes;
    return prev;
}
