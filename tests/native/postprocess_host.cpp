// TEST INFRASTRUCTURE ONLY.  Compiles partdistillation_b200/csrc/postprocess_math.cuh — the per-pixel arithmetic the
// post-processing kernels run on the device — for the HOST, so that tests/test_postprocess_host_cpu.py can compare it
// with the oracle in the build container (no GPU there).  Built by the test into a temporary directory; never part
// of libpdb200.so and never loaded by the product.
#include "../../partdistillation_b200/csrc/postprocess_math.cuh"

using namespace pdb;

// logits (K, h, w) -> out (K, Ho, Wo): the value pdb_postprocess_masks thresholds, with the entry point's own choice
// between the one-pass and the two-pass kernel and its scale expressions.
extern "C" void pp_host_resize(const float* logits, float* out, int K, int h, int w, int Hp, int Wp, int Hi, int Wi,
                               int Ho, int Wo) {
    const float s1h = (float)h / (float)Hp, s1w = (float)w / (float)Wp;
    const float s2h = (float)Hi / (float)Ho, s2w = (float)Wi / (float)Wo;
    const bool two = !(Hi == Ho && Wi == Wo);
    for (int oy = 0; oy < Ho; ++oy)
        for (int ox = 0; ox < Wo; ++ox) {
            PixelTaps t = two ? make_pixel_taps<true>(oy, ox, h, w, Hi, Wi, s1h, s1w, s2h, s2w)
                              : make_pixel_taps<false>(oy, ox, h, w, Hi, Wi, s1h, s1w, s2h, s2w);
            for (int k = 0; k < K; ++k) {
                const float* plane = logits + (int64_t)k * h * w;
                out[((int64_t)k * Ho + oy) * Wo + ox] = two ? sample_pixel<true>(plane, t) : sample_pixel<false>(plane, t);
            }
        }
}

// masks (G, Hp, Wp) 0/1 -> out (G, Ho, Wo) 0/1, as pdb_resize_masks_u8
extern "C" void pp_host_resize_masks(const uint8_t* masks, uint8_t* out, int G, int Hp, int Wp, int Hi, int Wi, int Ho,
                                     int Wo) {
    const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
    for (int g = 0; g < G; ++g)
        for (int oy = 0; oy < Ho; ++oy)
            for (int ox = 0; ox < Wo; ++ox)
                out[((int64_t)g * Ho + oy) * Wo + ox] =
                    resized_mask_bit(masks + (int64_t)g * Hp * Wp, Wp, Hi, Wi, oy, ox, sh, sw) ? 1 : 0;
}
