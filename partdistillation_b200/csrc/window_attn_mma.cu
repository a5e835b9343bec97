// Shifted-window attention of a Swin block on the tensor cores (frozen backbone, forward only; reference
// modeling/backbone/swin.py:78-176,239-300).  Same contract and index arithmetic as swin_window_attention_kernel
// (window_attn.cu: pad / roll / partition / shift mask / reverse / crop folded into the token mapping); the two products of
// a (window, head) — S = Q K^T (N x N x 32) and O = softmax(S) V (N x 32 x N), N = ws^2 = 144 for Swin-B — run as warp-level
// mma.sync.m16n8k8 TF32 instructions instead of 2 x N^2 x 32 scalar FMAs (the FFMA kernel is the largest kernel of the C3
// step and the second largest of the C2 step).  These are 16-row register-resident tiles of a 144 x 144
// problem per CTA, not a 128-row tcgen05 tile: mma.sync is the instruction that fits.
//
// PASSES = 3: fp32 accuracy from TF32 tensor cores by the hi / lo operand split of gemm_tc.cu (x = hi + lo, hi = the 19 bits
// the tensor core reads, lo = x - hi: a_hi b_hi + a_lo b_hi + a_hi b_lo; the dropped lo lo term is 2^-22 relative).
// PASSES = 1: single TF32 pass (10-bit mantissa), used under bf16 autocast, where the reference runs these matmuls in bf16.
//
// One CTA per (image, window, head), one warp per 16 query rows.  K and V of the window live in shared memory with a row
// stride of 36 floats (conflict-free B-fragment reads, see below); Q fragments, the 16 x N score fragments, the softmax and the
// 16 x 32 output fragments stay in registers.  The C-fragment of S (thread holds columns 2t, 2t+1 of every 8-wide key tile)
// feeds the A-fragment of the second product directly by permuting the contraction index: "k = t" is key 8j + 2t and
// "k = t + 4" is key 8j + 2t + 1, and the V rows of the B-fragment are read in the same order — no shuffles, no smem round trip.
#include <cuda_bf16.h>
#include "common.cuh"

namespace pdb {

constexpr int kMD = 32;            // head dim
constexpr int kKS = 36;            // shared-memory row stride (floats) of K and V

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lo_bits(float x) {          // x - trunc_tf32(x), exact in fp32
    return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}

template <int PASSES, int NT>      // NT = N / 8 key tiles (18 for ws = 12)
__global__ void __launch_bounds__(NT * 16, NT * 16 <= 288 ? 2 : 1)
swin_window_attention_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ bias,
                                 float* __restrict__ out, int H, int W, int heads, int ws, int shift, int Hp, int Wp, float scale,
                                 int out_bf16) {
    constexpr int N = NT * 8;
    extern __shared__ __align__(16) float s_kv[];      // K [N][36] | V [N][36] | region id [N]
    float* s_k = s_kv;
    float* s_v = s_kv + N * kKS;
    int* s_id = reinterpret_cast<int*>(s_kv + 2 * N * kKS);
    const int nwx = Wp / ws, nwy = Hp / ws;
    int r = blockIdx.x;
    const int h = r % heads; r /= heads;
    const int wx = r % nwx; r /= nwx;
    const int wy = r % nwy;
    const int b = r / nwy;
    const int C = heads * kMD, C3 = 3 * C;
    // token -> source row of qkv (or -1 for a padded token), as swin_window_attention_kernel
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
    for (int i = threadIdx.x; i < N * (kMD / 4); i += blockDim.x) {
        const int row = i >> 3, c4 = (i & 7) * 4;
        int region;
        const int64_t src = source(row, region);
        float4 kk, vv;
        if (src >= 0) {
            const float* p = qkv + src * C3 + h * kMD + c4;
            kk = __ldg(reinterpret_cast<const float4*>(p + C));
            vv = __ldg(reinterpret_cast<const float4*>(p + 2 * C));
        } else if (qkv_bias) {
            kk = __ldg(reinterpret_cast<const float4*>(qkv_bias + C + h * kMD + c4));
            vv = __ldg(reinterpret_cast<const float4*>(qkv_bias + 2 * C + h * kMD + c4));
        } else {
            kk = vv = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        *reinterpret_cast<float4*>(s_k + row * kKS + c4) = kk;
        *reinterpret_cast<float4*>(s_v + row * kKS + c4) = vv;
        if ((i & 7) == 0) s_id[row] = region;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = warp * 16 + g, row1 = row0 + 8;            // the two query rows (tokens) of this thread
    int reg0, reg1;
    const int64_t src0 = source(row0, reg0), src1 = source(row1, reg1);
    constexpr float kLog2e = 1.4426950408889634f;
    const float qs = scale * kLog2e;

    // ---- Q fragments: a0 (row0, k0 + t), a1 (row1, k0 + t), a2 (row0, k0 + t + 4), a3 (row1, k0 + t + 4); padded query rows
    // read the Linear's bias like padded keys (their output is never stored)
    float qf[4][4];
    {
        const float* q0 = src0 >= 0 ? qkv + src0 * C3 + h * kMD : (qkv_bias ? qkv_bias + h * kMD : nullptr);
        const float* q1 = src1 >= 0 ? qkv + src1 * C3 + h * kMD : (qkv_bias ? qkv_bias + h * kMD : nullptr);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            qf[ks][0] = q0 ? __ldg(q0 + ks * 8 + t) * qs : 0.f;
            qf[ks][1] = q1 ? __ldg(q1 + ks * 8 + t) * qs : 0.f;
            qf[ks][2] = q0 ? __ldg(q0 + ks * 8 + t + 4) * qs : 0.f;
            qf[ks][3] = q1 ? __ldg(q1 + ks * 8 + t + 4) * qs : 0.f;
        }
    }

    // ---- S = (Q scale log2e) K^T: 16 x N per warp, NT tiles of 8 keys, 4 k-steps of 8 dims
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ah[i] = __float_as_uint(qf[ks][i]);
            al[i] = lo_bits(qf[ks][i]);
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            // B fragment: b0 = K[8j + g][8ks + t], b1 = K[8j + g][8ks + t + 4]: bank (4g + t) mod 32 — conflict-free
            const float k0 = s_k[(j * 8 + g) * kKS + ks * 8 + t], k1 = s_k[(j * 8 + g) * kKS + ks * 8 + t + 4];
            mma_tf32(s[j], ah, __float_as_uint(k0), __float_as_uint(k1));
            if (PASSES == 3) {
                mma_tf32(s[j], al, __float_as_uint(k0), __float_as_uint(k1));
                mma_tf32(s[j], ah, lo_bits(k0), lo_bits(k1));
            }
        }
    }

    // ---- + relative-position bias (+ shift mask), base-2 softmax over the row (thread: columns 8j + 2t, 8j + 2t + 1 of rows
    // row0 / row1; the 4 lanes of a quad hold one row)
    const float* b0p = bias + ((int64_t)h * N + row0) * N + 2 * t;
    const float* b1p = bias + ((int64_t)h * N + row1) * N + 2 * t;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const float2 ba = __ldg(reinterpret_cast<const float2*>(b0p + j * 8));
        const float2 bb = __ldg(reinterpret_cast<const float2*>(b1p + j * 8));
        const int ida = s_id[j * 8 + 2 * t], idb = s_id[j * 8 + 2 * t + 1];
        s[j][0] += (ba.x + (ida != reg0 ? -100.f : 0.f)) * kLog2e;
        s[j][1] += (ba.y + (idb != reg0 ? -100.f : 0.f)) * kLog2e;
        s[j][2] += (bb.x + (ida != reg1 ? -100.f : 0.f)) * kLog2e;
        s[j][3] += (bb.y + (idb != reg1 ? -100.f : 0.f)) * kLog2e;
        m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
        m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s[j][0] = exp2f(s[j][0] - m0); s[j][1] = exp2f(s[j][1] - m0);
        s[j][2] = exp2f(s[j][2] - m1); s[j][3] = exp2f(s[j][3] - m1);
        l0 += s[j][0] + s[j][1];
        l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- O = P V: A fragment of k-step j = the C fragment of key tile j with the contraction index permuted
    // (a0 = P[row0][8j + 2t], a1 = P[row1][8j + 2t], a2 = P[row0][8j + 2t + 1], a3 = P[row1][8j + 2t + 1]); the B fragment reads
    // V rows in the same order: b0 = V[8j + 2t][8n + g], b1 = V[8j + 2t + 1][8n + g]: banks (8t + g), (8t + 4 + g) mod 32
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        uint32_t ph[4], pl[4];
        ph[0] = __float_as_uint(s[j][0]); ph[1] = __float_as_uint(s[j][2]); ph[2] = __float_as_uint(s[j][1]); ph[3] = __float_as_uint(s[j][3]);
        if (PASSES == 3) {
            pl[0] = lo_bits(s[j][0]); pl[1] = lo_bits(s[j][2]); pl[2] = lo_bits(s[j][1]); pl[3] = lo_bits(s[j][3]);
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float v0 = s_v[(j * 8 + 2 * t) * kKS + n * 8 + g], v1 = s_v[(j * 8 + 2 * t + 1) * kKS + n * 8 + g];
            mma_tf32(o[n], ph, __float_as_uint(v0), __float_as_uint(v1));
            if (PASSES == 3) {
                mma_tf32(o[n], pl, __float_as_uint(v0), __float_as_uint(v1));
                mma_tf32(o[n], ph, lo_bits(v0), lo_bits(v1));
            }
        }
    }

    // ---- normalise and store: thread holds out[row0][8n + 2t, + 1] and out[row1][...]
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const int64_t col = h * kMD + n * 8 + 2 * t;
        if (out_bf16) {         // the consumer is the bf16 projection GEMM of the autocast path: no separate conversion pass
            __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out);
            if (src0 >= 0) *reinterpret_cast<__nv_bfloat162*>(ob + src0 * C + col) = __floats2bfloat162_rn(o[n][0] * i0, o[n][1] * i0);
            if (src1 >= 0) *reinterpret_cast<__nv_bfloat162*>(ob + src1 * C + col) = __floats2bfloat162_rn(o[n][2] * i1, o[n][3] * i1);
        } else {
            if (src0 >= 0) *reinterpret_cast<float2*>(out + src0 * C + col) = make_float2(o[n][0] * i0, o[n][1] * i0);
            if (src1 >= 0) *reinterpret_cast<float2*>(out + src1 * C + col) = make_float2(o[n][2] * i1, o[n][3] * i1);
        }
    }
}

template <int PASSES, int NT>
static int launch_mma(const float* qkv, const float* qkv_bias, const float* bias, float* out, int B, int H, int W, int heads, int ws,
                      int shift, float scale, int out_bf16, cudaStream_t st) {
    const int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
    const int64_t ctas = (int64_t)B * (Hp / ws) * (Wp / ws) * heads;
    PDB_REQUIRE(ctas < (1ll << 31), "swin_window_attention: too many windows");
    constexpr int N = NT * 8;
    const size_t smem = sizeof(float) * 2 * N * kKS + sizeof(int) * N;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(swin_window_attention_mma_kernel<PASSES, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "swin_window_attention_tc: smem attribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    swin_window_attention_mma_kernel<PASSES, NT><<<(unsigned)ctas, NT * 16, smem, st>>>(qkv, qkv_bias, bias, out, H, W, heads, ws, shift,
                                                                                      Hp, Wp, scale, out_bf16);
    return launched("swin_window_attention_mma");
}

}  // namespace pdb

using namespace pdb;

// passes: 3 = fp32-accurate (3xTF32), 1 = single TF32 pass (bf16-autocast path).  Window sizes 12 (N = 144), 8 (N = 64) and 4
// (N = 16: the micro trunks of the tests); other sizes: PDB_ERR_UNSUPPORTED (callers use pdb_swin_window_attention_forward).
extern "C" int pdb_swin_window_attention_forward_tc(const float* qkv, const float* qkv_bias, const float* bias, void* out, int B,
                                                    int H, int W, int heads, int d, int ws, int shift, float scale, int passes,
                                                    int out_bf16, void* stream) {
    PDB_REQUIRE(qkv && bias && out, "swin_window_attention_tc: null pointer");
    PDB_REQUIRE(d == kMD, "swin_window_attention_tc: head dim %d (only 32)", d);
    PDB_REQUIRE(B > 0 && H > 0 && W > 0 && heads > 0 && shift >= 0 && shift < ws, "swin_window_attention_tc: bad sizes");
    PDB_REQUIRE(passes == 1 || passes == 3, "swin_window_attention_tc: passes %d (1 or 3)", passes);
    cudaStream_t st = as_stream(stream);
#define PDB_WIN_CASE(WS, NT_)                                                                                              \
    if (ws == WS)                                                                                                          \
        return passes == 3 ? launch_mma<3, NT_>(qkv, qkv_bias, bias, (float*)out, B, H, W, heads, ws, shift, scale, out_bf16, st) \
                           : launch_mma<1, NT_>(qkv, qkv_bias, bias, (float*)out, B, H, W, heads, ws, shift, scale, out_bf16, st);
    PDB_WIN_CASE(12, 18)
    PDB_WIN_CASE(8, 8)
    PDB_WIN_CASE(4, 2)
#undef PDB_WIN_CASE
    return fail(PDB_ERR_UNSUPPORTED, "swin_window_attention_tc: window size %d (12, 8 or 4)", ws);
}
