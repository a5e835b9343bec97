// Inference post-processing of the ProposalModel eval branch (SURVEY.md §8 row f4;
// reference proposal_model.py:220-302,381-432 + detectron2 sem_seg_postprocess + pycocotools rleIou).
//
// The reference materialises, per image, the (Q, Hpad, Wpad) fp32 up-sampled logits (F.interpolate, :225-230), a
// second (Q, Hout, Wout) fp32 copy (sem_seg_postprocess, :243), the gated copy (:375), sigmoid / product / top-k
// maps (:260-265) and finally moves every bool mask to the host to run-length encode it for the IoU
// (utils/utils.py:35-42).  Here the two bilinear passes are composed on the fly per output pixel straight from the
// (Q, h, w) logits, the result lives as ONE BIT per (query, pixel) (32 pixels of a row per word, written by a warp
// ballot), areas are popcounts and the pairwise intersections are popc(a & b) sums — the fp32 maps never exist.
//
// Numerics: every bilinear pass evaluates ATen's upsample_bilinear2d expression
//   h0*(w0*a + w1*b) + h1*(w0*c + w1*d),   src = scale*(dst + 0.5) - 0.5 clamped at 0,  scale = (float)in/out,
// with every product and sum rounded separately (__fmul_rn / __fadd_rn: no FMA contraction, so the bits do not
// depend on the compiler).  When the output size equals the cropped size the second pass has lambda = 0 and is the
// identity, bit for bit; it is skipped.
#include "common.cuh"
#include "postprocess_kernels.cuh"

using namespace pdb;

extern "C" int pdb_postprocess_masks(const float* logits, const int32_t* sel, const float* scores, const uint8_t* gate,
                                     uint32_t* bits, int32_t* label, uint32_t* score_bits, float score_thr, int Q,
                                     int K, int h, int w, int Hp, int Wp, int Hi, int Wi, int Ho, int Wo,
                                     void* stream) {
    PDB_REQUIRE(logits && sel, "postprocess_masks: null pointer");
    PDB_REQUIRE(bits || label || score_bits, "postprocess_masks: neither bits nor label requested");
    PDB_REQUIRE((label == nullptr && score_bits == nullptr) || scores != nullptr,
                "postprocess_masks: label map / score bits needs the scores");
    if (label == nullptr && score_bits == nullptr) scores = nullptr;     // the kernel keys the sigmoid work on `scores`
    PDB_REQUIRE(Q > 0 && K > 0 && h > 0 && w > 0 && Hp > 0 && Wp > 0 && Ho > 0 && Wo > 0,
                "postprocess_masks: non-positive dimension");
    PDB_REQUIRE(Hi > 0 && Wi > 0 && Hi <= Hp && Wi <= Wp, "postprocess_masks: image size (%d, %d) outside the padded size (%d, %d)",
                Hi, Wi, Hp, Wp);
    PDB_REQUIRE((int64_t)Q * h * w < (1ll << 40), "postprocess_masks: logits too large");
    PDB_REQUIRE((int64_t)h * w < (1ll << 31), "postprocess_masks: plane too large");
    const int Ww = (Wo + 31) / 32;
    PDB_REQUIRE((Ho + 7) / 8 <= 65535, "postprocess_masks: output height %d exceeds grid.y", Ho);
    // area_pixel_compute_scale(align_corners=False, no explicit scale factor): (float)in / out
    const float s1h = (float)h / (float)Hp, s1w = (float)w / (float)Wp;
    const float s2h = (float)Hi / (float)Ho, s2w = (float)Wi / (float)Wo;
    const dim3 grid = row_word_grid(Ho, Wo, 1), block = row_word_block();
    if (Hi == Ho && Wi == Wo)
        postprocess_masks_kernel<false><<<grid, block, 0, as_stream(stream)>>>(
            logits, sel, scores, gate, bits, label, score_bits, score_thr, K, h, w, Hi, Wi, Ho, Wo, Ww, s1h, s1w, s2h, s2w);
    else
        postprocess_masks_kernel<true><<<grid, block, 0, as_stream(stream)>>>(
            logits, sel, scores, gate, bits, label, score_bits, score_thr, K, h, w, Hi, Wi, Ho, Wo, Ww, s1h, s1w, s2h, s2w);
    return launched("postprocess_masks");
}

extern "C" int pdb_resize_masks_u8(const uint8_t* masks, uint8_t* out, int G, int Hp, int Wp, int Hi, int Wi, int Ho,
                                   int Wo, void* stream) {
    PDB_REQUIRE(masks && out, "resize_masks_u8: null pointer");
    PDB_REQUIRE(G > 0 && G <= 65535 && Hp > 0 && Wp > 0 && Ho > 0 && Wo > 0, "resize_masks_u8: bad shape");
    PDB_REQUIRE(Hi > 0 && Wi > 0 && Hi <= Hp && Wi <= Wp, "resize_masks_u8: image size outside the padded size");
    const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
    resize_masks_u8_kernel<<<per_pixel_grid(Ho, Wo, G), 256, 0, as_stream(stream)>>>(masks, out, Hp, Wp, Hi, Wi, Ho, Wo, sh, sw);
    return launched("resize_masks_u8");
}

extern "C" int pdb_pack_bits(const uint8_t* in, uint32_t* bits, int R, int Ho, int Wo, void* stream) {
    PDB_REQUIRE(in && bits, "pack_bits: null pointer");
    PDB_REQUIRE(R > 0 && R <= 65535 && Ho > 0 && Wo > 0 && (Ho + 7) / 8 <= 65535, "pack_bits: bad shape");
    const int Ww = (Wo + 31) / 32;
    pack_bits_kernel<<<row_word_grid(Ho, Wo, R), row_word_block(), 0, as_stream(stream)>>>(in, bits, Ho, Wo, Ww);
    return launched("pack_bits");
}

extern "C" int pdb_unpack_bits(const uint32_t* bits, const int32_t* rows, uint8_t* out, int R, int Ho, int Wo,
                               void* stream) {
    PDB_REQUIRE(bits && out, "unpack_bits: null pointer");
    PDB_REQUIRE(R > 0 && R <= 65535 && Ho > 0 && Wo > 0 && (Ho + 7) / 8 <= 65535, "unpack_bits: bad shape");
    const int Ww = (Wo + 31) / 32;
    unpack_bits_kernel<<<row_word_grid(Ho, Wo, R), row_word_block(), 0, as_stream(stream)>>>(bits, rows, out, Ho, Wo, Ww);
    return launched("unpack_bits");
}

extern "C" int pdb_bits_popcount(const uint32_t* bits, int64_t* counts, int rows, int64_t words, void* stream) {
    PDB_REQUIRE(bits && counts, "bits_popcount: null pointer");
    PDB_REQUIRE(rows > 0 && rows <= 65535 && words > 0, "bits_popcount: bad shape");
    bits_popcount_kernel<<<chunk_grid(words, rows), 256, 0, as_stream(stream)>>>(bits, reinterpret_cast<unsigned long long*>(counts), words);
    return launched("bits_popcount");
}

extern "C" int pdb_bits_intersect(const uint32_t* a, const uint32_t* b, int64_t* inter, int Ka, int Kb, int64_t words,
                                  void* stream) {
    PDB_REQUIRE(a && b && inter, "bits_intersect: null pointer");
    PDB_REQUIRE(Ka > 0 && Ka <= 65535 && Kb > 0 && words > 0, "bits_intersect: bad shape");
    bits_intersect_kernel<<<chunk_grid(words, Ka), 256, 0, as_stream(stream)>>>(a, b, reinterpret_cast<unsigned long long*>(inter), Kb, words);
    return launched("bits_intersect");
}
