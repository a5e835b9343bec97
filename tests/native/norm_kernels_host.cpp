// TEST INFRASTRUCTURE ONLY.  The LayerNorm (csrc/layernorm.cu) and GroupNorm (csrc/groupnorm.cu) kernels compiled for the
// host through cuda_on_cpu.h.  tests/test_norm_kernels_host_cpu.py cuts each file's `namespace pdb { ... }` block into
// layernorm_section.inc / groupnorm_section.inc; the entry points restate the launchers at the end of the two files.
#include "pdb_common_host.h"

#include "layernorm_section.inc"
#include "groupnorm_section.inc"

using namespace pdb;
using cpu_cuda::launch;

extern "C" int host_layer_norm_forward_scaled(const float* x, const float* residual, const float* res_scale, int64_t rows_per_sample,
                                              const float* weight, const float* bias, void* y, float* sum_out, float* mean,
                                              float* rstd, int64_t rows, int C, float eps, int y_bf16) {
    if (!(rows >= 0 && C > 0 && C % 4 == 0 && C <= 2048)) return -1;
    if (res_scale && !(residual && rows_per_sample > 0)) return -1;
    if (rows == 0) return 0;
    const int C4 = C / 4;
    const int64_t blocks = (rows + 7) / 8;
#define HOST_LN_LAUNCH(V)                                                                                                   \
    launch(dim3((unsigned)blocks), dim3(256), [&] {                                                                         \
        layer_norm_fwd_kernel<V>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(residual),             \
                                 reinterpret_cast<const float4*>(weight), reinterpret_cast<const float4*>(bias),            \
                                 reinterpret_cast<float4*>(y), reinterpret_cast<float4*>(sum_out), mean, rstd, rows, C4,    \
                                 1.f / (float)C, eps, res_scale, rows_per_sample, y_bf16);                                          \
    })
    if (C4 <= 32) HOST_LN_LAUNCH(1);
    else if (C4 <= 64) HOST_LN_LAUNCH(2);
    else if (C4 <= 128) HOST_LN_LAUNCH(4);
    else if (C4 <= 256) HOST_LN_LAUNCH(8);
    else HOST_LN_LAUNCH(16);
#undef HOST_LN_LAUNCH
    return 0;
}

extern "C" int host_layer_norm_forward(const float* x, const float* residual, const float* weight, const float* bias, float* y,
                                       float* sum_out, float* mean, float* rstd, int64_t rows, int C, float eps) {
    return host_layer_norm_forward_scaled(x, residual, nullptr, 0, weight, bias, y, sum_out, mean, rstd, rows, C, eps, 0);
}

extern "C" int host_group_norm_forward(const float* x, const float* weight, const float* bias, float* y, double* stats,
                                       float* mean, float* rstd, int B, int64_t HW, int C, int G, float eps, int relu) {
    PDB_TRY(gn_check("group_norm_forward", B, HW, C, G));
    const int C4 = C / 4, qpg = (C / G) / 4;
    int64_t chunks, ppc;
    gn_chunks(B, HW, C4, chunks, ppc);
    launch(dim3((unsigned)chunks, (unsigned)B), dim3(GN_THREADS),
           [&] { gn_stats_kernel(reinterpret_cast<const float4*>(x), stats, HW, C4, G, qpg, ppc); });
    launch(dim3((unsigned)((B * G + 127) / 128)), dim3(128),
           [&] { gn_finalize_kernel(stats, mean, rstd, B * G, 1.0 / ((double)HW * (C / G)), (double)eps); });
    const int64_t pixels = (int64_t)B * HW;
    const unsigned blocks = gn_apply_blocks(pixels, C4);
    const float4 *x4 = reinterpret_cast<const float4*>(x), *w4 = reinterpret_cast<const float4*>(weight);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
    float4* y4 = reinterpret_cast<float4*>(y);
    if (relu)
        launch(dim3(blocks), dim3(GN_THREADS), [&] { gn_apply_kernel<true>(x4, w4, b4, mean, rstd, y4, HW, pixels, C4, G, qpg); });
    else
        launch(dim3(blocks), dim3(GN_THREADS), [&] { gn_apply_kernel<false>(x4, w4, b4, mean, rstd, y4, HW, pixels, C4, G, qpg); });
    return 0;
}

extern "C" int host_group_norm_backward(const float* dy, const float* x, const float* weight, const float* bias,
                                        const float* mean, const float* rstd, double* chan_sums, float* coef, float* dx, int B,
                                        int64_t HW, int C, int G, int relu) {
    PDB_TRY(gn_check("group_norm_backward", B, HW, C, G));
    const int C4 = C / 4, qpg = (C / G) / 4;
    int64_t chunks, ppc;
    gn_chunks(B, HW, C4, chunks, ppc);
    const dim3 sgrid((unsigned)chunks, (unsigned)B);
    const float4 *dy4 = reinterpret_cast<const float4*>(dy), *x4 = reinterpret_cast<const float4*>(x);
    const float4 *w4 = reinterpret_cast<const float4*>(weight), *b4 = reinterpret_cast<const float4*>(bias);
    if (relu)
        launch(sgrid, dim3(GN_THREADS), [&] { gn_bwd_stats_kernel<true>(dy4, x4, w4, b4, mean, rstd, chan_sums, HW, C4, G, qpg, ppc); });
    else
        launch(sgrid, dim3(GN_THREADS), [&] { gn_bwd_stats_kernel<false>(dy4, x4, w4, b4, mean, rstd, chan_sums, HW, C4, G, qpg, ppc); });
    if (!dx) return 0;
    launch(dim3((unsigned)((B * G + 127) / 128)), dim3(128),
           [&] { gn_bwd_group_kernel(chan_sums, weight, coef, B * G, G, C, C / G, 1.0 / ((double)HW * (C / G))); });
    const int64_t pixels = (int64_t)B * HW;
    const unsigned blocks = gn_apply_blocks(pixels, C4);
    float4* dx4 = reinterpret_cast<float4*>(dx);
    if (relu)
        launch(dim3(blocks), dim3(GN_THREADS),
               [&] { gn_bwd_apply_kernel<true>(dy4, x4, w4, b4, mean, rstd, coef, dx4, HW, pixels, C4, G, qpg); });
    else
        launch(dim3(blocks), dim3(GN_THREADS),
               [&] { gn_bwd_apply_kernel<false>(dy4, x4, w4, b4, mean, rstd, coef, dx4, HW, pixels, C4, G, qpg); });
    return 0;
}
