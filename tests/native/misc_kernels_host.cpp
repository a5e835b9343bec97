// TEST INFRASTRUCTURE ONLY.  Host builds (through cuda_on_cpu.h) of three small kernel files of partdistillation_b200/csrc:
// optim.cu (flat gradient norm + AdamW), grouping.cu (one-pass pixel-group affinity) and attn_mask.cu (attention-mask
// build / reset).  tests/test_misc_kernels_host_cpu.py cuts their `namespace pdb` blocks into *_section.inc; the entry
// points restate the launchers of the three files (C ABI argument order minus the stream).
#include "pdb_common_host.h"

#include "../../partdistillation_b200/csrc/grouping_resized.cuh"     // group_scores_kernel (stage 1 of the grouping)
#include "optim_section.inc"
#include "grouping_section.inc"
#include "attn_mask_section.inc"

using namespace pdb;
using cpu_cuda::launch;

extern "C" int host_grad_sumsq(const float* grad, int64_t n, float grad_scale, double* out) {
    if (!(grad && out && n >= 0 && n % 4 == 0)) return -1;
    *out = 0.0;                                                             // cudaMemsetAsync in the launcher
    if (n == 0) return 0;
    const int64_t n4 = n / 4;
    const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 4 * kNumSMs);
    launch(dim3((unsigned)blocks), dim3(256), [&] { grad_sumsq_kernel(reinterpret_cast<const float4*>(grad), n4, grad_scale, out); });
    return 0;
}

extern "C" int host_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               const int64_t* seg_start, const float* seg_lr, const float* seg_wd, int num_segs, float beta1,
                               float beta2, float eps, const int64_t* step, float grad_scale, float clip_norm,
                               const double* sumsq) {
    if (!(n > 0 && n % 4 == 0 && num_segs > 0 && step)) return -1;
    const int64_t n4 = n / 4;
    const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 8 * kNumSMs);
    launch(dim3((unsigned)blocks), dim3(256), [&] {
        adamw_flat_kernel(reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg),
                          reinterpret_cast<float4*>(exp_avg_sq), n4, seg_start, seg_lr, seg_wd, num_segs, beta1, beta2, eps, step,
                          grad_scale, clip_norm, sumsq);
    });
    return 0;
}

extern "C" int host_group_affinity(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels, int C, int Kc,
                                   int h, int w, int H, int W, int metric) {
    if (!(C > 0 && Kc > 0 && Kc <= kMaxCentroids)) return -1;
    const size_t smem = sizeof(float) * ((size_t)C * Kc + Kc);
    launch(dim3((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8)), dim3(256), smem,
           [&] { group_affinity_kernel(feat, centroids, mask, labels, C, Kc, h, w, H, W, metric, 0); });
    return 0;
}

extern "C" int host_group_scores(const float* feat, const float* centroids, float* scores, int C, int Kc, int h, int w, int metric) {
    if (!(C > 0 && Kc > 0 && Kc <= kMaxGroupCentroids)) return -1;
    const int hw = h * w;
    launch(dim3((unsigned)((hw + 31) / 32)), dim3(256), [&] { group_scores_kernel(feat, centroids, scores, C, Kc, hw, metric); });
    return 0;
}

extern "C" int host_attn_mask_build(const float* logits, uint8_t* mask, int32_t* row_any, int B, int Q, int H, int W, int h,
                                    int w) {
    const float rh = (float)H / (float)h, rw = (float)W / (float)w;
    launch(dim3((unsigned)((h * w + 255) / 256), (unsigned)(B * Q)), dim3(256),
           [&] { attn_mask_kernel(logits, mask, row_any, H, W, h, w, rh, rw); });
    return 0;
}

extern "C" int host_attn_mask_reset_rows(uint8_t* mask, const int32_t* row_any, int rows, int64_t hw) {
    launch(dim3((unsigned)((hw + 255) / 256), (unsigned)rows), dim3(256), [&] { attn_mask_reset_kernel(mask, row_any, hw); });
    return 0;
}
