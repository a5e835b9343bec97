// Optimizer step on the flat gradient buffer: what the reference does with
// FullModelGradientClippingOptimizer = clip_grad_norm_(all parameters, CLIP_VALUE) followed by AdamW with per-parameter
// lr / weight decay (base_trainer.py:65-147), as two passes over one flat fp32 buffer instead of one kernel launch per
// parameter group:
//   pdb_grad_sumsq   : sum_i (grad_scale * g_i)^2  in float64 (block partials + one atomicAdd(double) per block)
//   pdb_adamw_flat   : coef = min(1, clip / (sqrt(sumsq) + 1e-6));  g = grad_scale * coef * g;  decoupled weight decay,
//                      moments, bias-corrected update — torch.optim.AdamW arithmetic, element for element.
// Parameters occupy 4-element-aligned segments of the flat buffers; segment s = [seg_start[s], seg_start[s+1]) has its
// own lr / weight decay (seg tables live in device memory; a thread finds its segment by binary search).
#include <algorithm>
#include <cmath>
#include "common.cuh"

namespace pdb {

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float4* __restrict__ g, int64_t n4, float scale, double* __restrict__ out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(g + i);
        float a = v.x * scale, b = v.y * scale, c = v.z * scale, d = v.w * scale;
        acc += (double)(a * a + b * b) + (double)(c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += part[w];
        atomicAdd(out, s);
    }
}

__global__ void __launch_bounds__(256)
adamw_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                  int64_t n4, const int64_t* __restrict__ seg_start, const float* __restrict__ seg_lr,
                  const float* __restrict__ seg_wd, int num_segs, float beta1, float beta2, float eps,
                  const int64_t* __restrict__ step_ptr, float grad_scale, float clip_norm, const double* __restrict__ sumsq) {
    // bias corrections from the device-side step counter (so a CUDA graph of the training step stays valid)
    const double step = (double)*step_ptr;
    const float bc1 = (float)(1.0 - pow((double)beta1, step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
    float coef = grad_scale;
    if (clip_norm > 0.f && sumsq) {
        float total = (float)sqrt(*sumsq);
        coef *= fminf(clip_norm / (total + 1e-6f), 1.0f);
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        int lo = 0, hi = num_segs - 1;              // last segment with seg_start <= e
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (__ldg(seg_start + mid) <= e) lo = mid; else hi = mid - 1;
        }
        const float lr = __ldg(seg_lr + lo), wd = __ldg(seg_wd + lo);
        float4 pp = p[i], gg = __ldg(g + i), mm = m[i], vv = v[i];
        float* P = &pp.x; float* G = &gg.x; float* Mo = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = G[k] * coef;
            float x = P[k] * (1.f - lr * wd);
            const float m1 = Mo[k] + (1.f - beta1) * (gr - Mo[k]);          // lerp, as torch's fused kernel
            const float v1 = beta2 * V[k] + (1.f - beta2) * gr * gr;
            const float denom = sqrtf(v1) / bc2_sqrt + eps;
            x -= (lr / bc1) * (m1 / denom);
            P[k] = x; Mo[k] = m1; V[k] = v1;
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_grad_sumsq(const float* grad, int64_t n, float grad_scale, double* out, void* stream) {
    PDB_REQUIRE(grad && out && n >= 0 && n % 4 == 0, "grad_sumsq: null pointer or n not a multiple of 4");
    PDB_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad_sumsq: buffer must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(out, 0, sizeof(double), st);
    if (n == 0) return PDB_OK;
    int64_t n4 = n / 4;
    int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 4 * kNumSMs);
    grad_sumsq_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(grad), n4, grad_scale, out);
    return launched("grad_sumsq");
}

extern "C" int pdb_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              const int64_t* seg_start, const float* seg_lr, const float* seg_wd, int num_segs, float beta1,
                              float beta2, float eps, const int64_t* step, float grad_scale, float clip_norm, const double* sumsq,
                              void* stream) {
    PDB_REQUIRE(param && grad && exp_avg && exp_avg_sq && seg_start && seg_lr && seg_wd, "adamw_flat: null pointer");
    PDB_REQUIRE(step, "adamw_flat: null step pointer");
    PDB_REQUIRE(n > 0 && n % 4 == 0 && num_segs > 0, "adamw_flat: bad sizes (n %% 4 == 0)");
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adamw_flat: buffers must be 16-byte aligned");
    int64_t n4 = n / 4;
    int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 8 * kNumSMs);
    adamw_flat_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg),
        reinterpret_cast<float4*>(exp_avg_sq), n4, seg_start, seg_lr, seg_wd, num_segs, beta1, beta2, eps, step,
        grad_scale, clip_norm, sumsq);
    return launched("adamw_flat");
}
