// Masked cross-attention core on the tensor cores: the kernels of xattn.cu (same contract, same workspace, same combine
// kernel) with the two products per (query, key) pair — and the five of the backward — as warp-level mma.sync.m16n8k8
// TF32 instructions on register-resident fragments, 3 passes per product (hi / lo operand split of gemm_tc.cu: fp32-accurate).
// Reference: nn.MultiheadAttention inside CrossAttentionLayer.forward_post (mask2former_transformer_decoder.py:84,102-114);
// SelfAttentionLayer (:44-54) runs on the same kernels with mask = NULL.
//
// Fragment bookkeeping (m16n8k8, g = lane / 4, t = lane % 4): a C fragment holds (row g | g + 8, columns 2t, 2t + 1) of an
// 8-wide tile.  Used as the A operand of a following product it is read as "k = t -> column 2t, k = t + 4 -> column 2t + 1",
// and the B operand's rows are loaded in the same order, so P -> P V, dS -> dS K, P^T -> P^T dO and dS^T -> dS^T Q need no
// shuffle and no shared-memory round trip.  The backward for K / V computes the TRANSPOSED scores S^T = K Q^T directly (rows =
// keys), which turns both of its reductions over the queries into such chained products as well.
// K / V / Q / dO tiles sit in shared memory with a row stride of 36 floats: both B-fragment access patterns
// ([8j + g][8k + t] and [8j + 2t][8n + g]) are bank-conflict free.
#include <math.h>

#include "common.cuh"

namespace pdb {

constexpr int MXD = 32;            // head dim
constexpr int MXS = 36;            // shared-memory row stride
constexpr int MXT = 64;            // keys (or queries) per shared-memory tile = 8 fragments of 8

__device__ __forceinline__ void xm_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float xm_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// d += A B with fp32 accuracy: a_hi b_hi + a_lo b_hi + a_hi b_lo (the tensor core reads the top 19 bits of each operand).
// ONE: a single TF32 product (10-bit mantissas, fp32 accumulation) — the bf16-autocast path, whose reference rounds q, k, p, v to
// 8-bit mantissas.
template <bool ONE>
__device__ __forceinline__ void xm_mma3(float (&d)[4], const float (&a)[4], const float b0, const float b1) {
    if (ONE) {
        const uint32_t ar[4] = {__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3])};
        xm_mma(d, ar, __float_as_uint(b0), __float_as_uint(b1));
        return;
    }
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ah[i] = __float_as_uint(a[i]);
        al[i] = __float_as_uint(xm_lo(a[i]));
    }
    xm_mma(d, ah, __float_as_uint(b0), __float_as_uint(b1));
    xm_mma(d, al, __float_as_uint(b0), __float_as_uint(b1));
    xm_mma(d, ah, __float_as_uint(xm_lo(b0)), __float_as_uint(xm_lo(b1)));
}
// A fragment of rows (r0, r1) of a row-major [.][ld] matrix in GLOBAL memory (NULL row = zeros), k-step ks
__device__ __forceinline__ void xm_load_a(float (&a)[4][4], const float* __restrict__ r0, const float* __restrict__ r1, int t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a[ks][0] = r0 ? __ldg(r0 + ks * 8 + t) : 0.f;
        a[ks][1] = r1 ? __ldg(r1 + ks * 8 + t) : 0.f;
        a[ks][2] = r0 ? __ldg(r0 + ks * 8 + t + 4) : 0.f;
        a[ks][3] = r1 ? __ldg(r1 + ks * 8 + t + 4) : 0.f;
    }
}
constexpr float kXmLog2e = 1.4426950408889634f, kXmLn2 = 0.6931471805599453f;
__device__ __forceinline__ void xm_scale_a(float (&a)[4][4], float f) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[ks][i] *= f;
}
// 256-thread CTAs: a 64 x 32 tile is 512 float4 = 2 per thread.  Register stage of the NEXT tile (global loads in flight while
// the current tile is computed), then a store into the other shared buffer.
struct XmStage {
    float4 k[2], v[2];
};
__device__ __forceinline__ void xm_prefetch(XmStage& st, const float* __restrict__ k, const float* __restrict__ v, int64_t rowbase, int row0,
                                            int nrows, int ld, int hoff) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int e = threadIdx.x + i * 256, jj = e >> 3, c = e & 7;
        st.k[i] = st.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + jj < nrows) {
            const int64_t off = (rowbase + row0 + jj) * ld + hoff + c * 4;
            st.k[i] = __ldg(reinterpret_cast<const float4*>(k + off));
            st.v[i] = __ldg(reinterpret_cast<const float4*>(v + off));
        }
    }
}
__device__ __forceinline__ void xm_commit(const XmStage& st, float* __restrict__ sK, float* __restrict__ sV) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int e = threadIdx.x + i * 256, jj = e >> 3, c = e & 7;
        *reinterpret_cast<float4*>(sK + jj * MXS + c * 4) = st.k[i];
        *reinterpret_cast<float4*>(sV + jj * MXS + c * 4) = st.v[i];
    }
}
// rows [row0, row0 + 64) of a (rows x heads*32) matrix -> shared tile [64][36] (rows >= nrows: zeros)
__device__ __forceinline__ void xm_load_tile(float* __restrict__ dst, const float* __restrict__ src, int64_t rowbase, int row0, int nrows,
                                             int ld, int hoff) {
    for (int e = threadIdx.x; e < MXT * (MXD / 4); e += blockDim.x) {
        const int jj = e >> 3, c = e & 7;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + jj < nrows) x = __ldg(reinterpret_cast<const float4*>(src + (rowbase + row0 + jj) * ld + hoff + c * 4));
        *reinterpret_cast<float4*>(dst + jj * MXS + c * 4) = x;
    }
}
// acc[j] (j = 0..7) += A (16 x 32, fragments a) x T^T for the 64 rows of the shared tile T ([row][36])
template <int NF = 8, bool ONE = false>
__device__ __forceinline__ void xm_scores(float (&acc)[NF][4], const float (&a)[4][4], const float* __restrict__ tile, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int j = 0; j < NF; ++j)
            xm_mma3<ONE>(acc[j], a[ks], tile[(j * 8 + g) * MXS + ks * 8 + t], tile[(j * 8 + g) * MXS + ks * 8 + t + 4]);
}
// out[n] (n = 0..3) += P (16 x 64, as the C fragments p[j] of a previous product) x T for the 64 rows of the shared tile T
template <int NF = 8, bool ONE = false>
__device__ __forceinline__ void xm_chain(float (&out)[4][4], const float (&p)[NF][4], const float* __restrict__ tile, int g, int t) {
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        const float a[4] = {p[j][0], p[j][2], p[j][1], p[j][3]};
#pragma unroll
        for (int n = 0; n < 4; ++n)
            xm_mma3<ONE>(out[n], a, tile[(j * 8 + 2 * t) * MXS + n * 8 + g], tile[(j * 8 + 2 * t + 1) * MXS + n * 8 + g]);
    }
}

// masked flags of this thread's 2 x 2 entries of every 8-key fragment of a tile: bit (2j + c) of m0 / m1 = row0 / row1, key
// j0 + 8j + 2t + c is masked or beyond Lk
__device__ __forceinline__ void xm_mask_rows(const uint8_t* __restrict__ mr0, const uint8_t* __restrict__ mr1, int j0, int Lk, int t,
                                             unsigned& m0, unsigned& m1) {
    m0 = m1 = 0u;
    const bool pair = (Lk & 1) == 0;           // row base and key index even: the two bytes can be read as one ushort
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ka = j0 + j * 8 + 2 * t;
        unsigned a0 = 0, a1 = 0, b0 = 0, b1 = 0;
        if (ka + 1 < Lk && pair) {
            if (mr0) { const unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(mr0 + ka)); a0 = w & 0xff; a1 = w >> 8; }
            if (mr1) { const unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(mr1 + ka)); b0 = w & 0xff; b1 = w >> 8; }
        } else {
            if (ka < Lk) { if (mr0) a0 = __ldg(mr0 + ka); if (mr1) b0 = __ldg(mr1 + ka); } else { a0 = b0 = 1; }
            if (ka + 1 < Lk) { if (mr0) a1 = __ldg(mr0 + ka + 1); if (mr1) b1 = __ldg(mr1 + ka + 1); } else { a1 = b1 = 1; }
        }
        m0 |= (a0 ? 1u : 0u) << (2 * j) | (a1 ? 1u : 0u) << (2 * j + 1);
        m1 |= (b0 ? 1u : 0u) << (2 * j) | (b1 ? 1u : 0u) << (2 * j + 1);
    }
}

struct XmRow {          // per-thread state of its two query rows
    int qi0, qi1;
    bool act0, act1;
    const uint8_t *mr0, *mr1;
};
__device__ __forceinline__ XmRow xm_rows(int qt, int warp, int g, int Q, int b, int Lk, const uint8_t* mask, const int32_t* row_any) {
    XmRow r;
    r.qi0 = qt * 128 + warp * 16 + g;
    r.qi1 = r.qi0 + 8;
    r.act0 = r.qi0 < Q;
    r.act1 = r.qi1 < Q;
    auto use = [&](int qi) { return mask != nullptr && (row_any == nullptr || row_any[(int64_t)b * Q + qi] != 0); };
    r.mr0 = (r.act0 && use(r.qi0)) ? mask + ((int64_t)b * Q + r.qi0) * Lk : nullptr;
    r.mr1 = (r.act1 && use(r.qi1)) ? mask + ((int64_t)b * Q + r.qi1) * Lk : nullptr;
    return r;
}

// ------------------------------------------------------------------------------------------------ forward (partial over a key split)
template <bool ONE>
__global__ void __launch_bounds__(256)
xattn_fwd_partial_mma(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const uint8_t* __restrict__ mask, const int32_t* __restrict__ row_any, float* __restrict__ ws_acc,
                      float* __restrict__ ws_ml, int heads, int Q, int Lk, int nsplit, int tiles_per) {
    __shared__ __align__(16) float sKb[2][MXT * MXS];          // double-buffered: tile t + 1 lands while tile t is computed
    __shared__ __align__(16) float sVb[2][MXT * MXS];
    const int split = blockIdx.x % nsplit, qt = blockIdx.x / nsplit;
    const int h = blockIdx.y, b = blockIdx.z;
    const int ld = heads * MXD, hoff = h * MXD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const XmRow R = xm_rows(qt, warp, g, Q, b, Lk, mask, row_any);
    float qf[4][4];
    xm_load_a(qf, R.act0 ? q + ((int64_t)b * Q + R.qi0) * ld + hoff : nullptr, R.act1 ? q + ((int64_t)b * Q + R.qi1) * ld + hoff : nullptr, t);
    xm_scale_a(qf, kXmLog2e);                  // base-2 softmax inside the kernel: exp2 is one MUFU; (m, l) leave in natural units
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int tile_begin = split * tiles_per;
    const int tile_end = min(tile_begin + tiles_per, (Lk + MXT - 1) / MXT);
    XmStage stg;
    if (tile_begin < tile_end) {
        xm_prefetch(stg, k, v, (int64_t)b * Lk, tile_begin * MXT, Lk, ld, hoff);
        xm_commit(stg, sKb[0], sVb[0]);
    }
    __syncthreads();
    for (int tl = tile_begin; tl < tile_end; ++tl) {
        const int j0 = tl * MXT;
        const float* sK = sKb[(tl - tile_begin) & 1];
        const float* sV = sVb[(tl - tile_begin) & 1];
        const bool more = tl + 1 < tile_end;
        if (more) xm_prefetch(stg, k, v, (int64_t)b * Lk, j0 + MXT, Lk, ld, hoff);
        // the barrier at the end of the iteration publishes the next tile; every path below must reach it
        bool skip;
        unsigned k0, k1;
        xm_mask_rows(R.mr0, R.mr1, j0, Lk, t, k0, k1);
        if (!R.act0) k0 = 0xffffu;
        if (!R.act1) k1 = 0xffffu;
        skip = __all_sync(0xffffffffu, (k0 & k1) == 0xffffu);                // nothing attended in this warp's 16 x 64 block
        if (!skip) {
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        xm_scores<8, ONE>(s, qf, sK, g, t);
        float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if ((k0 >> (2 * j)) & 1u) s[j][0] = -INFINITY;
            if ((k0 >> (2 * j + 1)) & 1u) s[j][1] = -INFINITY;
            if ((k1 >> (2 * j)) & 1u) s[j][2] = -INFINITY;
            if ((k1 >> (2 * j + 1)) & 1u) s[j][3] = -INFINITY;
            t0 = fmaxf(t0, fmaxf(s[j][0], s[j][1]));
            t1 = fmaxf(t1, fmaxf(s[j][2], s[j][3]));
        }
        t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1)); t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
        t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1)); t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
        const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);
        const float c0 = (m0 == -INFINITY) ? 0.f : exp2f(m0 - n0), c1 = (m1 == -INFINITY) ? 0.f : exp2f(m1 - n1);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = (s[j][0] == -INFINITY) ? 0.f : exp2f(s[j][0] - n0);
            s[j][1] = (s[j][1] == -INFINITY) ? 0.f : exp2f(s[j][1] - n0);
            s[j][2] = (s[j][2] == -INFINITY) ? 0.f : exp2f(s[j][2] - n1);
            s[j][3] = (s[j][3] == -INFINITY) ? 0.f : exp2f(s[j][3] - n1);
            a0 += s[j][0] + s[j][1];
            a1 += s[j][2] + s[j][3];
        }
        l0 = l0 * c0 + a0;             // per-thread partial of the row sum (the rescale factor is uniform over the quad)
        l1 = l1 * c1 + a1;
#pragma unroll
        for (int n = 0; n < 4; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
        m0 = n0; m1 = n1;
        xm_chain<8, ONE>(o, s, sV, g, t);
        }
        if (more) xm_commit(stg, sKb[(tl + 1 - tile_begin) & 1], sVb[(tl + 1 - tile_begin) & 1]);
        __syncthreads();
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int64_t base = (((int64_t)b * heads + h) * nsplit + split) * Q;
    if (R.act0) {
        float* oa = ws_acc + (base + R.qi0) * MXD;
#pragma unroll
        for (int n = 0; n < 4; ++n) *reinterpret_cast<float2*>(oa + n * 8 + 2 * t) = make_float2(o[n][0], o[n][1]);
        if (t == 0) { ws_ml[(base + R.qi0) * 2] = m0 * kXmLn2; ws_ml[(base + R.qi0) * 2 + 1] = l0; }
    }
    if (R.act1) {
        float* oa = ws_acc + (base + R.qi1) * MXD;
#pragma unroll
        for (int n = 0; n < 4; ++n) *reinterpret_cast<float2*>(oa + n * 8 + 2 * t) = make_float2(o[n][2], o[n][3]);
        if (t == 0) { ws_ml[(base + R.qi1) * 2] = m1 * kXmLn2; ws_ml[(base + R.qi1) * 2 + 1] = l1; }
    }
}

// ------------------------------------------------------------------------------------------------ grad_q (partial over a key split)
template <bool ONE>
__global__ void __launch_bounds__(256)
xattn_bwd_dq_mma(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                 const uint8_t* __restrict__ mask, const int32_t* __restrict__ row_any, const float* __restrict__ out,
                 const float* __restrict__ lse, const float* __restrict__ gout, float* __restrict__ gq, int heads, int Q, int Lk,
                 int nsplit, int tiles_per) {
    __shared__ __align__(16) float sKb[2][MXT * MXS];
    __shared__ __align__(16) float sVb[2][MXT * MXS];
    const int split = blockIdx.x % nsplit, qt = blockIdx.x / nsplit;
    const int h = blockIdx.y, b = blockIdx.z;
    const int ld = heads * MXD, hoff = h * MXD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const XmRow R = xm_rows(qt, warp, g, Q, b, Lk, mask, row_any);
    const int64_t r0 = ((int64_t)b * Q + R.qi0) * ld + hoff, r1 = ((int64_t)b * Q + R.qi1) * ld + hoff;
    float qf[4][4], df[4][4], of[4][4];
    xm_load_a(qf, R.act0 ? q + r0 : nullptr, R.act1 ? q + r1 : nullptr, t);
    xm_load_a(df, R.act0 ? gout + r0 : nullptr, R.act1 ? gout + r1 : nullptr, t);
    xm_load_a(of, R.act0 ? out + r0 : nullptr, R.act1 ? out + r1 : nullptr, t);
    float d0 = 0.f, d1 = 0.f;                 // delta = dO . O per row: this thread holds 8 of the 32 dims of each row
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        d0 = fmaf(df[ks][0], of[ks][0], fmaf(df[ks][2], of[ks][2], d0));
        d1 = fmaf(df[ks][1], of[ks][1], fmaf(df[ks][3], of[ks][3], d1));
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    xm_scale_a(qf, kXmLog2e);                  // P = exp(S - lse) = exp2(S log2e - lse log2e)
    const float e0 = R.act0 ? lse[((int64_t)b * heads + h) * Q + R.qi0] * kXmLog2e : 0.f;
    const float e1 = R.act1 ? lse[((int64_t)b * heads + h) * Q + R.qi1] * kXmLog2e : 0.f;
    float dq[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
    const int tile_begin = split * tiles_per;
    const int tile_end = min(tile_begin + tiles_per, (Lk + MXT - 1) / MXT);
    XmStage stg;
    if (tile_begin < tile_end) {
        xm_prefetch(stg, k, v, (int64_t)b * Lk, tile_begin * MXT, Lk, ld, hoff);
        xm_commit(stg, sKb[0], sVb[0]);
    }
    __syncthreads();
    for (int tl = tile_begin; tl < tile_end; ++tl) {
        const int j0 = tl * MXT;
        const float* sK = sKb[(tl - tile_begin) & 1];
        const float* sV = sVb[(tl - tile_begin) & 1];
        const bool more = tl + 1 < tile_end;
        if (more) xm_prefetch(stg, k, v, (int64_t)b * Lk, j0 + MXT, Lk, ld, hoff);
        unsigned k0, k1;
        xm_mask_rows(R.mr0, R.mr1, j0, Lk, t, k0, k1);
        if (!R.act0) k0 = 0xffffu;
        if (!R.act1) k1 = 0xffffu;
        if (!__all_sync(0xffffffffu, (k0 & k1) == 0xffffu)) {
        float s[8][4], dp[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
        }
        xm_scores<8, ONE>(s, qf, sK, g, t);
        xm_scores<8, ONE>(dp, df, sV, g, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) {           // dS = P (dP - delta), P = exp(S - lse) on the attended keys
            s[j][0] = ((k0 >> (2 * j)) & 1u) ? 0.f : exp2f(s[j][0] - e0) * (dp[j][0] - d0);
            s[j][1] = ((k0 >> (2 * j + 1)) & 1u) ? 0.f : exp2f(s[j][1] - e0) * (dp[j][1] - d0);
            s[j][2] = ((k1 >> (2 * j)) & 1u) ? 0.f : exp2f(s[j][2] - e1) * (dp[j][2] - d1);
            s[j][3] = ((k1 >> (2 * j + 1)) & 1u) ? 0.f : exp2f(s[j][3] - e1) * (dp[j][3] - d1);
        }
        xm_chain<8, ONE>(dq, s, sK, g, t);
        }
        if (more) xm_commit(stg, sKb[(tl + 1 - tile_begin) & 1], sVb[(tl + 1 - tile_begin) & 1]);
        __syncthreads();
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        if (R.act0) { atomicAdd(gq + r0 + n * 8 + 2 * t, dq[n][0]); atomicAdd(gq + r0 + n * 8 + 2 * t + 1, dq[n][1]); }
        if (R.act1) { atomicAdd(gq + r1 + n * 8 + 2 * t, dq[n][2]); atomicAdd(gq + r1 + n * 8 + 2 * t + 1, dq[n][3]); }
    }
}

// ------------------------------------------------------------------------------------------------ grad_k / grad_v
// CTA = KT consecutive 64-key tiles of one (b, h); warp w owns keys 16w + g (+ 8) of the current tile.  ALL queries of the (b, h)
// (Q, dO, lse, delta = dO . O, mask-use flags; QP = Q rounded up to 32 rows) are staged in shared memory once, after which the
// warps run without any block barrier: per key tile and per 32 queries, S^T = K Q^T and dP^T = V dO^T, then the two chained
// products dV += P^T dO and dK += dS^T Q.
template <bool ONE>
__global__ void __launch_bounds__(128)
xattn_bwd_dkv_mma(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                  const uint8_t* __restrict__ mask, const int32_t* __restrict__ row_any, const float* __restrict__ out,
                  const float* __restrict__ lse, const float* __restrict__ gout, float* __restrict__ gk, float* __restrict__ gv,
                  int heads, int Q, int Lk, int QP, int KT) {
    extern __shared__ __align__(16) float xm_smem[];
    float* sQ = xm_smem;
    float* sDO = sQ + QP * MXS;
    float* sLse = sDO + QP * MXS;              // lse * log2(e)
    float* sDelta = sLse + QP;
    int* sUse = reinterpret_cast<int*>(sDelta + QP);     // 1: masked row, 0: unmasked row (no mask / reset row), -1: beyond Q
    const int h = blockIdx.y, b = blockIdx.z;
    const int ld = heads * MXD, hoff = h * MXD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int e = threadIdx.x; e < QP * (MXD / 4); e += 128) {
        const int jj = e >> 3, c = e & 7;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
        if (jj < Q) {
            const int64_t off = ((int64_t)b * Q + jj) * ld + hoff + c * 4;
            x = __ldg(reinterpret_cast<const float4*>(q + off));
            y = __ldg(reinterpret_cast<const float4*>(gout + off));
        }
        *reinterpret_cast<float4*>(sQ + jj * MXS + c * 4) = x;
        *reinterpret_cast<float4*>(sDO + jj * MXS + c * 4) = y;
    }
    for (int ii = warp; ii < QP; ii += 4) {               // one warp per query: delta = dO . O, lse, mask use
        float pr = 0.f;
        if (ii < Q) {
            const int64_t ro = ((int64_t)b * Q + ii) * ld + hoff + lane;
            pr = __ldg(gout + ro) * __ldg(out + ro);
        }
        pr = warp_sum(pr);
        if (lane == 0) {
            sDelta[ii] = pr;
            sLse[ii] = (ii < Q) ? lse[((int64_t)b * heads + h) * Q + ii] * kXmLog2e : 0.f;
            sUse[ii] = (ii < Q) ? ((mask != nullptr && (row_any == nullptr || row_any[(int64_t)b * Q + ii] != 0)) ? 1 : 0) : -1;
        }
    }
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int tile = blockIdx.x * KT + kt;
        if (tile * MXT >= Lk) break;
        const int key0 = tile * MXT + warp * 16 + g, key1 = key0 + 8;
        const bool a0 = key0 < Lk, a1 = key1 < Lk;
        const int64_t o0 = ((int64_t)b * Lk + key0) * ld + hoff, o1 = ((int64_t)b * Lk + key1) * ld + hoff;
        float kf[4][4], vf[4][4];
        xm_load_a(kf, a0 ? k + o0 : nullptr, a1 ? k + o1 : nullptr, t);
        xm_load_a(vf, a0 ? v + o0 : nullptr, a1 ? v + o1 : nullptr, t);
        xm_scale_a(kf, kXmLog2e);              // S^T in log2 units (kf is only used for the scores)
        float dk[4][4], dv[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
            dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
        }
#pragma unroll 1
        for (int i0 = 0; i0 < QP; i0 += 32) {
            // masked flags of this thread's entries: bit (2j + c) of f0 / f1 = (key0 / key1, query i0 + 8j + 2t + c)
            unsigned f0 = 0u, f1 = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int qq = i0 + j * 8 + 2 * t + c;
                    const int use = sUse[qq];
                    bool x0 = use < 0 || !a0, x1 = use < 0 || !a1;
                    if (use == 1) {
                        const uint8_t* mrow = mask + ((int64_t)b * Q + qq) * Lk;
                        if (a0 && __ldg(mrow + key0)) x0 = true;
                        if (a1 && __ldg(mrow + key1)) x1 = true;
                    }
                    f0 |= (x0 ? 1u : 0u) << (2 * j + c);
                    f1 |= (x1 ? 1u : 0u) << (2 * j + c);
                }
            if (__all_sync(0xffffffffu, (f0 & f1) == 0xffu)) continue;
            const float* tq = sQ + i0 * MXS;
            const float* td = sDO + i0 * MXS;
            float st[4][4], dp[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
                dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
            }
            xm_scores<4, ONE>(st, kf, tq, g, t);          // S^T = K Q^T
            xm_scores<4, ONE>(dp, vf, td, g, t);          // dP^T = V dO^T
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int qa = i0 + j * 8 + 2 * t;
                const float la = sLse[qa], lb = sLse[qa + 1];
                const float da = sDelta[qa], db = sDelta[qa + 1];
                st[j][0] = ((f0 >> (2 * j)) & 1u) ? 0.f : exp2f(st[j][0] - la);
                st[j][1] = ((f0 >> (2 * j + 1)) & 1u) ? 0.f : exp2f(st[j][1] - lb);
                st[j][2] = ((f1 >> (2 * j)) & 1u) ? 0.f : exp2f(st[j][2] - la);
                st[j][3] = ((f1 >> (2 * j + 1)) & 1u) ? 0.f : exp2f(st[j][3] - lb);
                dp[j][0] = st[j][0] * (dp[j][0] - da);
                dp[j][1] = st[j][1] * (dp[j][1] - db);
                dp[j][2] = st[j][2] * (dp[j][2] - da);
                dp[j][3] = st[j][3] * (dp[j][3] - db);
            }
            xm_chain<4, ONE>(dv, st, td, g, t);           // dV += P^T dO
            xm_chain<4, ONE>(dk, dp, tq, g, t);           // dK += dS^T Q
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (a0) {
                *reinterpret_cast<float2*>(gk + o0 + n * 8 + 2 * t) = make_float2(dk[n][0], dk[n][1]);
                *reinterpret_cast<float2*>(gv + o0 + n * 8 + 2 * t) = make_float2(dv[n][0], dv[n][1]);
            }
            if (a1) {
                *reinterpret_cast<float2*>(gk + o1 + n * 8 + 2 * t) = make_float2(dk[n][2], dk[n][3]);
                *reinterpret_cast<float2*>(gv + o1 + n * 8 + 2 * t) = make_float2(dv[n][2], dv[n][3]);
            }
        }
    }
}

// launchers used by xattn.cu's C ABI (same grids as the SIMT kernels: 128 queries per CTA in the forward / grad_q pass)
int xattn_fwd_partial_mma_launch(const float* q, const float* k, const float* v, const uint8_t* mask, const int32_t* row_any,
                                 float* ws_acc, float* ws_ml, int B, int heads, int Q, int Lk, int ns, int tiles_per, int qtiles,
                                 int one_pass, cudaStream_t st) {
    dim3 grid((unsigned)(ns * qtiles), (unsigned)heads, (unsigned)B);
    if (one_pass) xattn_fwd_partial_mma<true><<<grid, 256, 0, st>>>(q, k, v, mask, row_any, ws_acc, ws_ml, heads, Q, Lk, ns, tiles_per);
    else xattn_fwd_partial_mma<false><<<grid, 256, 0, st>>>(q, k, v, mask, row_any, ws_acc, ws_ml, heads, Q, Lk, ns, tiles_per);
    return launched("xattn_fwd_partial_mma");
}
int xattn_bwd_mma_launch(const float* q, const float* k, const float* v, const uint8_t* mask, const int32_t* row_any, const float* out,
                         const float* lse, const float* gout, float* gq, float* gk, float* gv, int B, int heads, int Q, int Lk, int ns,
                         int tiles_per, int qtiles, int one_pass, cudaStream_t st) {
    dim3 grid((unsigned)(ns * qtiles), (unsigned)heads, (unsigned)B);
    if (one_pass) xattn_bwd_dq_mma<true><<<grid, 256, 0, st>>>(q, k, v, mask, row_any, out, lse, gout, gq, heads, Q, Lk, ns, tiles_per);
    else xattn_bwd_dq_mma<false><<<grid, 256, 0, st>>>(q, k, v, mask, row_any, out, lse, gout, gq, heads, Q, Lk, ns, tiles_per);
    PDB_TRY(launched("xattn_bwd_dq_mma"));
    const int tiles = (Lk + MXT - 1) / MXT;
    const int QP = (Q + 31) / 32 * 32;
    int KT = (int)(((int64_t)tiles * heads * B) / (4 * kNumSMs));       // key tiles per CTA: keep >= ~600 CTAs in flight
    KT = KT < 1 ? 1 : (KT > 8 ? 8 : KT);
    const size_t smem = sizeof(float) * ((size_t)2 * QP * MXS + 3 * QP);
    PDB_REQUIRE(smem <= 200 * 1024, "masked_xattn_backward: %d queries do not fit shared memory", Q);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(xattn_bwd_dkv_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(xattn_bwd_dkv_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "xattn_bwd_dkv_mma: smem attribute: %s", cudaGetErrorString(e));
        attr = smem;
    }
    dim3 grid2((unsigned)((tiles + KT - 1) / KT), (unsigned)heads, (unsigned)B);
    if (one_pass) xattn_bwd_dkv_mma<true><<<grid2, 128, smem, st>>>(q, k, v, mask, row_any, out, lse, gout, gk, gv, heads, Q, Lk, QP, KT);
    else xattn_bwd_dkv_mma<false><<<grid2, 128, smem, st>>>(q, k, v, mask, row_any, out, lse, gout, gk, gv, heads, Q, Lk, QP, KT);
    return launched("xattn_bwd_dkv_mma");
}

}  // namespace pdb
