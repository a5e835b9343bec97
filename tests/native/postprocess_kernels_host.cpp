// TEST INFRASTRUCTURE ONLY.  The post-processing kernels (partdistillation_b200/csrc/postprocess_kernels.cuh) compiled
// for the host through cuda_on_cpu.h, behind entry points with the signatures of include/pdb200.h (minus the stream) and
// the launch geometry of postprocess.cu.  tests/test_postprocess_host_cpu.py runs them against the oracle.
#include "cuda_on_cpu.h"
#include "../../partdistillation_b200/csrc/postprocess_kernels.cuh"
#include "../../partdistillation_b200/csrc/grouping_resized.cuh"

using namespace pdb;
using cpu_cuda::launch;

extern "C" int host_postprocess_masks(const float* logits, const int32_t* sel, const float* scores, const uint8_t* gate,
                                      uint32_t* bits, int32_t* label, uint32_t* score_bits, float score_thr, int Q, int K,
                                      int h, int w, int Hp, int Wp, int Hi, int Wi, int Ho, int Wo) {
    if (label == nullptr && score_bits == nullptr) scores = nullptr;
    const int Ww = (Wo + 31) / 32;
    const float s1h = (float)h / (float)Hp, s1w = (float)w / (float)Wp;
    const float s2h = (float)Hi / (float)Ho, s2w = (float)Wi / (float)Wo;
    const dim3 grid = row_word_grid(Ho, Wo, 1), block = row_word_block();
    if (Hi == Ho && Wi == Wo)
        launch(grid, block, [&] {
            postprocess_masks_kernel<false>(logits, sel, scores, gate, bits, label, score_bits, score_thr, K, h, w, Hi, Wi,
                                            Ho, Wo, Ww, s1h, s1w, s2h, s2w);
        });
    else
        launch(grid, block, [&] {
            postprocess_masks_kernel<true>(logits, sel, scores, gate, bits, label, score_bits, score_thr, K, h, w, Hi, Wi,
                                           Ho, Wo, Ww, s1h, s1w, s2h, s2w);
        });
    return 0;
}

extern "C" int host_resize_masks_u8(const uint8_t* masks, uint8_t* out, int G, int Hp, int Wp, int Hi, int Wi, int Ho,
                                    int Wo) {
    const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
    launch(per_pixel_grid(Ho, Wo, G), dim3(256), [&] { resize_masks_u8_kernel(masks, out, Hp, Wp, Hi, Wi, Ho, Wo, sh, sw); });
    return 0;
}

extern "C" int host_pack_bits(const uint8_t* in, uint32_t* bits, int R, int Ho, int Wo) {
    const int Ww = (Wo + 31) / 32;
    launch(row_word_grid(Ho, Wo, R), row_word_block(), [&] { pack_bits_kernel(in, bits, Ho, Wo, Ww); });
    return 0;
}

extern "C" int host_unpack_bits(const uint32_t* bits, const int32_t* rows, uint8_t* out, int R, int Ho, int Wo) {
    const int Ww = (Wo + 31) / 32;
    launch(row_word_grid(Ho, Wo, R), row_word_block(), [&] { unpack_bits_kernel(bits, rows, out, Ho, Wo, Ww); });
    return 0;
}

extern "C" int host_bits_popcount(const uint32_t* bits, int64_t* counts, int rows, int64_t words) {
    launch(chunk_grid(words, rows), dim3(256),
           [&] { bits_popcount_kernel(bits, reinterpret_cast<unsigned long long*>(counts), words); });
    return 0;
}

extern "C" int host_bits_intersect(const uint32_t* a, const uint32_t* b, int64_t* inter, int Ka, int Kb, int64_t words) {
    launch(chunk_grid(words, Ka), dim3(256),
           [&] { bits_intersect_kernel(a, b, reinterpret_cast<unsigned long long*>(inter), Kb, words); });
    return 0;
}

extern "C" int host_group_scores(const float* feat, const float* centroids, float* scores, int C, int Kc, int h, int w, int metric) {
    if (!(C > 0 && Kc > 0 && Kc <= kMaxGroupCentroids)) return -1;
    const int hw = h * w;
    launch(dim3((unsigned)((hw + 31) / 32)), dim3(256), [&] { group_scores_kernel(feat, centroids, scores, C, Kc, hw, metric); });
    return 0;
}

extern "C" int host_group_affinity_resized(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels,
                                           int C, int Kc, int h, int w, int Hp, int Wp, int Hi, int Wi, int Ho, int Wo,
                                           int metric) {
    const float s1h = (float)h / (float)Hp, s1w = (float)w / (float)Wp;
    const float s2h = (float)Hi / (float)Ho, s2w = (float)Wi / (float)Wo;
    const size_t smem = group_affinity_smem(C, Kc);
    if (Hi == Ho && Wi == Wo)
        launch(group_affinity_grid(Ho, Wo), dim3(256), smem, [&] {
            group_affinity_resized_kernel<false>(feat, centroids, mask, labels, C, Kc, h, w, Hi, Wi, Ho, Wo, s1h, s1w, s2h, s2w, metric);
        });
    else
        launch(group_affinity_grid(Ho, Wo), dim3(256), smem, [&] {
            group_affinity_resized_kernel<true>(feat, centroids, mask, labels, C, Kc, h, w, Hi, Wi, Ho, Wo, s1h, s1w, s2h, s2w, metric);
        });
    return 0;
}
