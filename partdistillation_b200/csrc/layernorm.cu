// LayerNorm forward (optionally fused with the residual add in front of it) — replaces nn.LayerNorm / F.layer_norm of
// the encoder layers (msdeformattn.py:129-133), the decoder layers (mask2former_transformer_decoder.py:44-54,102-114,
// 167-171) and the Swin blocks (swin.py:239-300).  PyTorch's kernel runs at ~10 % of the HBM roofline for the short rows
// of this model (C = 128 .. 1024); here one warp owns one row, holds it in registers (float4 loads), and reduces mean
// and the centred sum of squares with shuffles, so the op is a single pass over the data.
//   z = x (+ residual);  y = (z - mean(z)) * rstd(z) * weight + bias
//   x, residual, y, sum_out: (rows, C) f32;  mean, rstd: (rows) f32 (saved for backward, as native_layer_norm)
//   sum_out (optional): z itself, for pre-norm residual streams that keep using the sum.
// C % 4 == 0 and C <= 2048.
#include <cuda_bf16.h>
#include "common.cuh"

namespace pdb {

template <int VEC>      // float4 per lane: C <= VEC * 128
__global__ void __launch_bounds__(256)
layer_norm_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ res, const float4* __restrict__ w,
                      const float4* __restrict__ b, float4* __restrict__ y, float4* __restrict__ sum_out,
                      float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int C4, float inv_c, float eps,
                      const float* __restrict__ res_scale, int64_t rows_per_sample, int y_bf16) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    // stochastic depth (swin_transformer.py:131-134: x = shortcut + drop_path(branch)): the branch of sample b enters scaled by
    // res_scale[b] = keep_b / (1 - p)
    const float sc = res_scale ? __ldg(res_scale + row / rows_per_sample) : 1.f;
    const float4* xr = x + row * C4;
    float4 v[VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = lane + i * 32;
        if (c < C4) {
            v[i] = __ldg(xr + c);
            if (res) {
                const float4 r = __ldg(res + row * C4 + c);
                v[i].x = fmaf(sc, r.x, v[i].x); v[i].y = fmaf(sc, r.y, v[i].y);
                v[i].z = fmaf(sc, r.z, v[i].z); v[i].w = fmaf(sc, r.w, v[i].w);
            }
            if (sum_out) sum_out[row * C4 + c] = v[i];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        } else {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        if (lane + i * 32 < C4) {
            const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + bb * bb) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = lane + i * 32;
        if (c < C4) {
            const float4 ww = __ldg(w + c), bv = __ldg(b + c);
            float4 o;
            o.x = (v[i].x - mean) * rstd * ww.x + bv.x;
            o.y = (v[i].y - mean) * rstd * ww.y + bv.y;
            o.z = (v[i].z - mean) * rstd * ww.z + bv.z;
            o.w = (v[i].w - mean) * rstd * ww.w + bv.w;
            if (y_bf16) {       // autocast: the consumer is a bf16 GEMM; same rounding as the cast it replaces
                const __nv_bfloat162 lo2 = __floats2bfloat162_rn(o.x, o.y), hi2 = __floats2bfloat162_rn(o.z, o.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo2);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi2);
                reinterpret_cast<uint2*>(y)[row * C4 + c] = pk;
            } else {
                y[row * C4 + c] = o;
            }
        }
    }
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_layer_norm_forward_scaled(const float* x, const float* residual, const float* res_scale, int64_t rows_per_sample,
                                             const float* weight, const float* bias, void* y, float* sum_out, float* mean,
                                             float* rstd, int64_t rows, int C, float eps, int y_bf16, void* stream) {
    PDB_REQUIRE(x && weight && bias && y && mean && rstd, "layer_norm: null pointer");
    PDB_REQUIRE(!res_scale || (residual && rows_per_sample > 0), "layer_norm: res_scale needs a residual and rows_per_sample > 0");
    PDB_REQUIRE(rows >= 0 && C > 0 && C % 4 == 0 && C <= 2048, "layer_norm: C=%d must be a multiple of 4, at most 2048", C);
    uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(weight) |
                   reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(sum_out);
    PDB_REQUIRE((al & 15) == 0, "layer_norm: buffers must be 16-byte aligned");
    if (rows == 0) return PDB_OK;
    const int C4 = C / 4;
    const int64_t blocks = (rows + 7) / 8;
    PDB_REQUIRE(blocks < (1ll << 31), "layer_norm: too many rows");
    cudaStream_t st = as_stream(stream);
#define PDB_LN_LAUNCH(V)                                                                                                   \
    layer_norm_fwd_kernel<V><<<(unsigned)blocks, 256, 0, st>>>(                                                             \
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(residual), reinterpret_cast<const float4*>(weight), \
        reinterpret_cast<const float4*>(bias), reinterpret_cast<float4*>(y), reinterpret_cast<float4*>(sum_out), mean, rstd, rows, \
        C4, 1.f / (float)C, eps, res_scale, rows_per_sample, y_bf16)
    if (C4 <= 32) PDB_LN_LAUNCH(1);
    else if (C4 <= 64) PDB_LN_LAUNCH(2);
    else if (C4 <= 128) PDB_LN_LAUNCH(4);
    else if (C4 <= 256) PDB_LN_LAUNCH(8);
    else PDB_LN_LAUNCH(16);
#undef PDB_LN_LAUNCH
    return launched("layer_norm_forward");
}

extern "C" int pdb_layer_norm_forward(const float* x, const float* residual, const float* weight, const float* bias, float* y,
                                      float* sum_out, float* mean, float* rstd, int64_t rows, int C, float eps, void* stream) {
    return pdb_layer_norm_forward_scaled(x, residual, nullptr, 0, weight, bias, y, sum_out, mean, rstd, rows, C, eps, 0, stream);
}
