// Swin window attention, forward (frozen backbone: SURVEY.md section 8 f2) — replaces, inside WindowAttention.forward
// (modeling/backbone/swin.py:78-176 of the reference), q @ k^T * scale + relative-position bias (+ shift mask) ->
// softmax -> @ v and the head transpose, which the reference runs as separate matmul / add / softmax / matmul kernels
// on (B*nW, heads, N, N) score tensors.
//   qkv  (Bw, N, 3, heads, 32) f32 (output of the qkv Linear)       bias (heads, N, N) f32
//   mask (nW, N, N) f32 additive or NULL (window of row bw is bw % nW)      out (Bw, N, heads*32) f32
// One CTA per (window, head); thread i owns query row i: q_i and the 32-wide output accumulator live in registers,
// K and V of the window (N x 32 each) in shared memory and are read as broadcasts; the softmax is computed online
// (running max / sum, flash-attention style) so the N x N scores are never stored.
#include "common.cuh"

namespace pdb {

constexpr int kWinD = 32;

// One key/value step of the online softmax for a query whose scaled q and accumulator live in registers.  The running
// reference exponent m is re-based lazily (p stays below 2^16 and the row sum below 2^24, far from fp32 overflow); the
// normalised result does not depend on the choice of m.  Base-2 softmax: q carries scale * log2(e), `add` (bias + mask)
// is already multiplied by log2(e).
__device__ __forceinline__ void attend_step(const float (&q)[kWinD], float (&acc)[kWinD], const float* __restrict__ kj,
                                            const float* __restrict__ vj, float add, float& m, float& l) {
    const float4* kp = reinterpret_cast<const float4*>(kj);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;          // four independent chains instead of one 32-long FMA chain
#pragma unroll
    for (int c = 0; c < kWinD / 4; ++c) {
        const float4 kv = kp[c];
        s0 = fmaf(q[4 * c], kv.x, s0); s1 = fmaf(q[4 * c + 1], kv.y, s1);
        s2 = fmaf(q[4 * c + 2], kv.z, s2); s3 = fmaf(q[4 * c + 3], kv.w, s3);
    }
    const float s = ((s0 + s1) + (s2 + s3)) + add;
    if (s > m + 16.f) {                 // lazy running maximum: only re-base when the exponent would grow past 2^16
        const float r = exp2f(m - s);   // exp2f(-inf) = 0 on the first key
        l *= r;
#pragma unroll
        for (int c = 0; c < kWinD; ++c) acc[c] *= r;
        m = s;
    }
    const float p = exp2f(s - m);
    l += p;
    const float4* vp = reinterpret_cast<const float4*>(vj);
#pragma unroll
    for (int c = 0; c < kWinD / 4; ++c) {
        const float4 vv = vp[c];
        acc[4 * c] = fmaf(p, vv.x, acc[4 * c]); acc[4 * c + 1] = fmaf(p, vv.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(p, vv.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(p, vv.w, acc[4 * c + 3]);
    }
}

__global__ void __launch_bounds__(256)
window_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ mask,
                        float* __restrict__ out, int N, int heads, int nW, float scale) {
    extern __shared__ __align__(16) float s_kv[];      // K [N][32] | V [N][32]
    float* s_k = s_kv;
    float* s_v = s_kv + N * kWinD;
    const int bw = blockIdx.x / heads;
    const int h = blockIdx.x - bw * heads;
    const int C3 = 3 * heads * kWinD;
    const float* base = qkv + (int64_t)bw * N * C3 + h * kWinD;
    for (int i = threadIdx.x; i < N * (kWinD / 4); i += blockDim.x) {
        const int row = i >> 3, c4 = (i & 7) * 4;
        const float* src = base + (int64_t)row * C3 + c4;
        *reinterpret_cast<float4*>(s_k + row * kWinD + c4) = __ldg(reinterpret_cast<const float4*>(src + heads * kWinD));
        *reinterpret_cast<float4*>(s_v + row * kWinD + c4) = __ldg(reinterpret_cast<const float4*>(src + 2 * heads * kWinD));
    }
    __syncthreads();
    const int i = threadIdx.x;
    if (i >= N) return;
    float q[kWinD], acc[kWinD];
    constexpr float kLog2e = 1.4426950408889634f;
    const float qs = scale * kLog2e;
    {
        const float4* qp = reinterpret_cast<const float4*>(base + (int64_t)i * C3);
#pragma unroll
        for (int c = 0; c < kWinD / 4; ++c) {
            const float4 v = __ldg(qp + c);
            q[4 * c] = v.x * qs; q[4 * c + 1] = v.y * qs; q[4 * c + 2] = v.z * qs; q[4 * c + 3] = v.w * qs;
        }
    }
#pragma unroll
    for (int c = 0; c < kWinD; ++c) acc[c] = 0.f;
    const float* brow = bias + ((int64_t)h * N + i) * N;
    const float* mrow = mask ? mask + ((int64_t)(bw % nW) * N + i) * N : nullptr;
    float m = -INFINITY, l = 0.f;
    // each thread streams its own bias (and mask) row: 8 columns = one 32-byte sector per load pair, so the row reads
    // cost 4 L1 wavefronts per key instead of 32
    const bool vec = (N % 8) == 0;
    for (int j0 = 0; j0 < N; j0 += 8) {
        float add[8];
        if (vec) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(brow + j0)), b1 = __ldg(reinterpret_cast<const float4*>(brow + j0 + 4));
            add[0] = b0.x; add[1] = b0.y; add[2] = b0.z; add[3] = b0.w; add[4] = b1.x; add[5] = b1.y; add[6] = b1.z; add[7] = b1.w;
            if (mrow) {
                const float4 m0 = __ldg(reinterpret_cast<const float4*>(mrow + j0)), m1 = __ldg(reinterpret_cast<const float4*>(mrow + j0 + 4));
                add[0] += m0.x; add[1] += m0.y; add[2] += m0.z; add[3] += m0.w; add[4] += m1.x; add[5] += m1.y; add[6] += m1.z; add[7] += m1.w;
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) add[u] = j0 + u < N ? __ldg(brow + j0 + u) + (mrow ? __ldg(mrow + j0 + u) : 0.f) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (j0 + u < N) attend_step(q, acc, s_k + (j0 + u) * kWinD, s_v + (j0 + u) * kWinD, add[u] * kLog2e, m, l);
    }
    const float inv = 1.f / l;
    float4* op = reinterpret_cast<float4*>(out + ((int64_t)bw * N + i) * heads * kWinD + h * kWinD);
#pragma unroll
    for (int c = 0; c < kWinD / 4; ++c)
        op[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_window_attention_forward(const float* qkv, const float* bias, const float* mask, float* out, int Bw, int N,
                                            int heads, int d, int nW, float scale, void* stream) {
    PDB_REQUIRE(qkv && bias && out, "window_attention: null pointer");
    PDB_REQUIRE(d == kWinD, "window_attention: head dim %d (only 32, Swin's C / heads)", d);
    PDB_REQUIRE(Bw > 0 && heads > 0 && N > 0 && N <= 256, "window_attention: bad sizes (N <= 256)");
    PDB_REQUIRE(!mask || (nW > 0 && Bw % nW == 0), "window_attention: Bw must be a multiple of the mask's window count");
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "window_attention: alignment");
    PDB_REQUIRE((int64_t)Bw * heads < (1ll << 31), "window_attention: too many (window, head) pairs");
    const int threads = ((N + 31) / 32) * 32;
    const size_t smem = sizeof(float) * 2 * N * kWinD;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr = true;
    }
    window_attention_kernel<<<(unsigned)(Bw * heads), threads, smem, as_stream(stream)>>>(qkv, bias, mask, out, N, heads,
                                                                                       mask ? nW : 1, scale);
    return launched("window_attention");
}

// ------------------------------------------------------------------------------------------------
// Whole shifted-window attention of a Swin block on the un-partitioned token grid: folds F.pad, torch.roll,
// window_partition, the (nW, N, N) shift mask, window_reverse, the inverse roll and the crop of
// SwinTransformerBlock.forward (swin.py:239-300) into the index arithmetic of the attention kernel.
//   qkv (B, H, W, 3, heads, 32) f32 = qkv Linear applied to norm1(x) in token order; padded tokens (the reference
//   pads the NORMALISED map with zeros, so their qkv is the Linear's bias) take qkv_bias (3*heads*32) or zeros;
//   bias (heads, N, N) relative-position bias, N = ws*ws;  out (B, H, W, heads*32).
// CTA = (image, window of the shifted frame, head); token t of the window sits at shifted-frame (wy*ws + t/ws,
// wx*ws + t%ws) = original pixel ((ys + shift) % Hp, (xs + shift) % Wp); the mask is -100 between tokens of
// different regions (0 | Hp-ws | Hp-shift bands per axis), as the reference's img_mask construction.
// ------------------------------------------------------------------------------------------------
namespace pdb {

__global__ void __launch_bounds__(256)
swin_window_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ bias,
                             float* __restrict__ out, int H, int W, int heads, int ws, int shift, int Hp, int Wp, float scale) {
    extern __shared__ __align__(16) float s_kv[];      // K [N][32] | V [N][32] | region id [N]
    const int N = ws * ws;
    float* s_k = s_kv;
    float* s_v = s_kv + N * kWinD;
    int* s_id = reinterpret_cast<int*>(s_kv + 2 * N * kWinD);
    const int nwx = Wp / ws, nwy = Hp / ws;
    int r = blockIdx.x;
    const int h = r % heads; r /= heads;
    const int wx = r % nwx; r /= nwx;
    const int wy = r % nwy;
    const int b = r / nwy;
    const int C = heads * kWinD, C3 = 3 * C;
    // token -> source row of qkv (or -1 for a padded token)
    auto source = [&](int t, int& region) -> int64_t {
        const int ys = wy * ws + t / ws, xs = wx * ws + t % ws;
        const int hr = ys < Hp - ws ? 0 : (ys < Hp - shift ? 1 : 2);
        const int wr = xs < Wp - ws ? 0 : (xs < Wp - shift ? 1 : 2);
        region = shift > 0 ? hr * 3 + wr : 0;
        int yo = ys + shift, xo = xs + shift;
        if (yo >= Hp) yo -= Hp;
        if (xo >= Wp) xo -= Wp;
        return (yo < H && xo < W) ? ((int64_t)b * H + yo) * W + xo : -1;
    };
    for (int i = threadIdx.x; i < N * (kWinD / 4); i += blockDim.x) {
        const int row = i >> 3, c4 = (i & 7) * 4;
        int region;
        const int64_t src = source(row, region);
        float4 kk, vv;
        if (src >= 0) {
            const float* p = qkv + src * C3 + h * kWinD + c4;
            kk = __ldg(reinterpret_cast<const float4*>(p + C));
            vv = __ldg(reinterpret_cast<const float4*>(p + 2 * C));
        } else if (qkv_bias) {
            kk = __ldg(reinterpret_cast<const float4*>(qkv_bias + C + h * kWinD + c4));
            vv = __ldg(reinterpret_cast<const float4*>(qkv_bias + 2 * C + h * kWinD + c4));
        } else {
            kk = vv = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        *reinterpret_cast<float4*>(s_k + row * kWinD + c4) = kk;
        *reinterpret_cast<float4*>(s_v + row * kWinD + c4) = vv;
        if ((i & 7) == 0) s_id[row] = region;
    }
    __syncthreads();
    const int i = threadIdx.x;
    if (i >= N) return;
    int my_region;
    const int64_t src = source(i, my_region);
    if (src < 0) return;                                  // padded query: its output row is cropped away
    float q[kWinD], acc[kWinD];
    constexpr float kLog2e = 1.4426950408889634f;
    const float qs = scale * kLog2e;
    {
        const float4* qp = reinterpret_cast<const float4*>(qkv + src * C3 + h * kWinD);
#pragma unroll
        for (int c = 0; c < kWinD / 4; ++c) {
            const float4 v = __ldg(qp + c);
            q[4 * c] = v.x * qs; q[4 * c + 1] = v.y * qs; q[4 * c + 2] = v.z * qs; q[4 * c + 3] = v.w * qs;
        }
    }
#pragma unroll
    for (int c = 0; c < kWinD; ++c) acc[c] = 0.f;
    const float* brow = bias + ((int64_t)h * N + i) * N;
    float m = -INFINITY, l = 0.f;
    const bool vec = (N % 8) == 0;
    for (int j0 = 0; j0 < N; j0 += 8) {
        float add[8];
        if (vec) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(brow + j0)), b1 = __ldg(reinterpret_cast<const float4*>(brow + j0 + 4));
            add[0] = b0.x; add[1] = b0.y; add[2] = b0.z; add[3] = b0.w; add[4] = b1.x; add[5] = b1.y; add[6] = b1.z; add[7] = b1.w;
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) add[u] = j0 + u < N ? __ldg(brow + j0 + u) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (j0 + u < N)
                attend_step(q, acc, s_k + (j0 + u) * kWinD, s_v + (j0 + u) * kWinD,
                            (add[u] + (s_id[j0 + u] != my_region ? -100.0f : 0.f)) * kLog2e, m, l);
    }
    const float inv = 1.f / l;
    float4* op = reinterpret_cast<float4*>(out + src * C + h * kWinD);
#pragma unroll
    for (int c = 0; c < kWinD / 4; ++c)
        op[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
}

}  // namespace pdb

extern "C" int pdb_swin_window_attention_forward(const float* qkv, const float* qkv_bias, const float* bias, float* out, int B,
                                                 int H, int W, int heads, int d, int ws, int shift, float scale, void* stream) {
    PDB_REQUIRE(qkv && bias && out, "swin_window_attention: null pointer");
    PDB_REQUIRE(d == kWinD, "swin_window_attention: head dim %d (only 32)", d);
    PDB_REQUIRE(B > 0 && H > 0 && W > 0 && heads > 0 && ws > 0 && ws * ws <= 256 && shift >= 0 && shift < ws,
                "swin_window_attention: bad sizes (ws*ws <= 256, 0 <= shift < ws)");
    const int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
    const int64_t ctas = (int64_t)B * (Hp / ws) * (Wp / ws) * heads;
    PDB_REQUIRE(ctas < (1ll << 31), "swin_window_attention: too many windows");
    const int N = ws * ws;
    const int threads = ((N + 31) / 32) * 32;
    const size_t smem = sizeof(float) * 2 * N * kWinD + sizeof(int) * N;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        cudaFuncSetAttribute(swin_window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr = true;
    }
    swin_window_attention_kernel<<<(unsigned)ctas, threads, smem, as_stream(stream)>>>(qkv, qkv_bias, bias, out, H, W, heads, ws,
                                                                                    shift, Hp, Wp, scale);
    return launched("swin_window_attention");
}
