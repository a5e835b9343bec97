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
#include "postprocess_math.cuh"

namespace pdb {

// grid (Ww, ceil(Ho / 8)), block (32, 8): a warp owns 32 consecutive pixels of one output row = one packed word.
template <bool TWO_STAGE>
__global__ void __launch_bounds__(256)
postprocess_masks_kernel(const float* __restrict__ logits, const int32_t* __restrict__ sel,
                         const float* __restrict__ scores, const uint8_t* __restrict__ gate,
                         uint32_t* __restrict__ bits, int32_t* __restrict__ label,
                         uint32_t* __restrict__ score_bits, float score_thr,
                         int K, int h, int w, int Hi, int Wi, int Ho, int Wo, int Ww,
                         float s1h, float s1w, float s2h, float s2w) {
    const int ox = blockIdx.x * 32 + threadIdx.x;
    const int oy = blockIdx.y * 8 + threadIdx.y;
    if (oy >= Ho) return;                                   // whole warp leaves together (threadIdx.y is per warp)
    const bool inside = ox < Wo;
    const int cx = inside ? ox : Wo - 1;                    // lanes past the row end compute a valid pixel, then drop it

    const PixelTaps taps = make_pixel_taps<TWO_STAGE>(oy, cx, h, w, Hi, Wi, s1h, s1w, s2h, s2w);
    const bool open = inside && (gate == nullptr || gate[(int64_t)oy * Wo + ox] != 0);
    const int64_t plane = (int64_t)h * w;
    const int64_t row_words = (int64_t)Ho * Ww;
    const int64_t word = (int64_t)oy * Ww + blockIdx.x;

    unsigned any_word = 0;
    float best = 0.f;
    int best_k = 0;
    for (int k = 0; k < K; ++k) {
        const float* src = logits + (int64_t)__ldg(sel + k) * plane;
        float v = sample_pixel<TWO_STAGE>(src, taps);
        if (!open) v = mul_rn(v, 0.f);                   // masks_per_image * object_target_mask (:375)
        const bool on = inside && (v > 0.f);
        const unsigned wbits = __ballot_sync(0xffffffffu, on);
        any_word |= wbits;
        if (bits != nullptr && threadIdx.x == 0) bits[(int64_t)k * row_words + word] = wbits;
        if (scores != nullptr) {
            // scores[:, None, None] * masks.sigmoid() -> topk(1, dim=0)[1]  (:262-264); first maximum wins
            float s = mul_rn(__ldg(scores + k), 1.0f / (1.0f + expf(-v)));
            if (k == 0 || s > best) { best = s; best_k = k; }
            if (score_bits != nullptr) {        // (predmask > thr) of part_distillation_model.py:379,385,391
                const unsigned sb = __ballot_sync(0xffffffffu, inside && (s > score_thr));
                if (threadIdx.x == 0) score_bits[(int64_t)k * row_words + word] = sb;
            }
        }
    }
    if (bits != nullptr && threadIdx.x == 0) bits[(int64_t)K * row_words + word] = any_word;   // topk(1, dim=0)[0] > 0 (:259)
    if (label != nullptr && inside) label[(int64_t)oy * Wo + ox] = best_k;
}

// Ground-truth masks: zero-padded bool (G, Hp, Wp) -> crop (Hi, Wi) -> bilinear as fp32 -> .bool()  (:244-245)
__global__ void __launch_bounds__(256)
resize_masks_u8_kernel(const uint8_t* __restrict__ masks, uint8_t* __restrict__ out, int Hp, int Wp, int Hi, int Wi,
                       int Ho, int Wo, float sh, float sw) {
    const int g = blockIdx.y;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (int64_t)Ho * Wo) return;
    const int oy = (int)(o / Wo), ox = (int)(o - (int64_t)oy * Wo);
    const uint8_t* src = masks + (int64_t)g * Hp * Wp;
    out[(int64_t)g * Ho * Wo + o] = resized_mask_bit(src, Wp, Hi, Wi, oy, ox, sh, sw) ? 1 : 0;
}

// grid (Ww, ceil(Ho / 8), R), block (32, 8)
__global__ void __launch_bounds__(256)
pack_bits_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ bits, int Ho, int Wo, int Ww) {
    const int ox = blockIdx.x * 32 + threadIdx.x;
    const int oy = blockIdx.y * 8 + threadIdx.y;
    if (oy >= Ho) return;
    const int64_t r = blockIdx.z;
    const bool on = ox < Wo && in[(r * Ho + oy) * Wo + ox] != 0;
    const unsigned wbits = __ballot_sync(0xffffffffu, on);
    if (threadIdx.x == 0) bits[(r * Ho + oy) * Ww + blockIdx.x] = wbits;
}

__global__ void __launch_bounds__(256)
unpack_bits_kernel(const uint32_t* __restrict__ bits, const int32_t* __restrict__ rows, uint8_t* __restrict__ out,
                   int Ho, int Wo, int Ww) {
    const int ox = blockIdx.x * 32 + threadIdx.x;
    const int oy = blockIdx.y * 8 + threadIdx.y;
    if (oy >= Ho || ox >= Wo) return;
    const int64_t r = blockIdx.z;
    const int64_t src = rows ? rows[r] : r;
    const uint32_t wbits = __ldg(bits + (src * Ho + oy) * Ww + blockIdx.x);
    out[(r * Ho + oy) * Wo + ox] = (wbits >> threadIdx.x) & 1u;
}

constexpr int kWordsPerThread = 8;
constexpr int kChunkWords = 256 * kWordsPerThread;

// grid (chunks, rows): counts[row] += popcount of the row's words in this chunk
__global__ void __launch_bounds__(256)
bits_popcount_kernel(const uint32_t* __restrict__ bits, unsigned long long* __restrict__ counts, int64_t words) {
    __shared__ int warp_part[8];
    const uint32_t* row = bits + (int64_t)blockIdx.y * words;
    const int64_t base = (int64_t)blockIdx.x * kChunkWords;
    int n = 0;
#pragma unroll
    for (int i = 0; i < kWordsPerThread; ++i) {
        int64_t idx = base + i * 256 + threadIdx.x;
        if (idx < words) n += __popc(__ldg(row + idx));
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += warp_part[i];
        if (t) atomicAdd(counts + blockIdx.y, (unsigned long long)t);
    }
}

constexpr int kInterTile = 64;      // rows of b handled per shared-memory accumulator pass

// grid (chunks, Ka): inter[i, j] += sum over this chunk of popc(a_i & b_j)
__global__ void __launch_bounds__(256)
bits_intersect_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                      unsigned long long* __restrict__ inter, int Kb, int64_t words) {
    __shared__ int acc[kInterTile];
    const int i = blockIdx.y;
    const int64_t base = (int64_t)blockIdx.x * kChunkWords;
    uint32_t wa[kWordsPerThread];
    bool any = false;
#pragma unroll
    for (int t = 0; t < kWordsPerThread; ++t) {
        int64_t idx = base + t * 256 + threadIdx.x;
        wa[t] = idx < words ? __ldg(a + (int64_t)i * words + idx) : 0u;
        any |= wa[t] != 0u;
    }
    if (!__syncthreads_or(any)) return;                     // this chunk of a_i is empty: nothing to add
    for (int j0 = 0; j0 < Kb; j0 += kInterTile) {
        const int nj = min(kInterTile, Kb - j0);
        if (threadIdx.x < kInterTile) acc[threadIdx.x] = 0;
        __syncthreads();
        for (int j = 0; j < nj; ++j) {
            const uint32_t* rb = b + (int64_t)(j0 + j) * words;
            int n = 0;
#pragma unroll
            for (int t = 0; t < kWordsPerThread; ++t) {
                int64_t idx = base + t * 256 + threadIdx.x;
                if (wa[t] != 0u) n += __popc(wa[t] & __ldg(rb + idx));     // wa != 0 implies idx < words
            }
            n = __reduce_add_sync(0xffffffffu, n);
            if ((threadIdx.x & 31) == 0 && n) atomicAdd(acc + j, n);
        }
        __syncthreads();
        if (threadIdx.x < nj && acc[threadIdx.x])
            atomicAdd(inter + (int64_t)i * Kb + j0 + threadIdx.x, (unsigned long long)acc[threadIdx.x]);
        __syncthreads();
    }
}

}  // namespace pdb

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
    const unsigned gy = (unsigned)((Ho + 7) / 8);
    PDB_REQUIRE(gy <= 65535u, "postprocess_masks: output height %d exceeds grid.y", Ho);
    // area_pixel_compute_scale(align_corners=False, no explicit scale factor): (float)in / out
    const float s1h = (float)h / (float)Hp, s1w = (float)w / (float)Wp;
    const float s2h = (float)Hi / (float)Ho, s2w = (float)Wi / (float)Wo;
    dim3 grid((unsigned)Ww, gy), block(32, 8);
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
    dim3 grid((unsigned)(((int64_t)Ho * Wo + 255) / 256), (unsigned)G);
    resize_masks_u8_kernel<<<grid, 256, 0, as_stream(stream)>>>(masks, out, Hp, Wp, Hi, Wi, Ho, Wo, sh, sw);
    return launched("resize_masks_u8");
}

extern "C" int pdb_pack_bits(const uint8_t* in, uint32_t* bits, int R, int Ho, int Wo, void* stream) {
    PDB_REQUIRE(in && bits, "pack_bits: null pointer");
    PDB_REQUIRE(R > 0 && R <= 65535 && Ho > 0 && Wo > 0 && (Ho + 7) / 8 <= 65535, "pack_bits: bad shape");
    const int Ww = (Wo + 31) / 32;
    dim3 grid((unsigned)Ww, (unsigned)((Ho + 7) / 8), (unsigned)R), block(32, 8);
    pack_bits_kernel<<<grid, block, 0, as_stream(stream)>>>(in, bits, Ho, Wo, Ww);
    return launched("pack_bits");
}

extern "C" int pdb_unpack_bits(const uint32_t* bits, const int32_t* rows, uint8_t* out, int R, int Ho, int Wo,
                               void* stream) {
    PDB_REQUIRE(bits && out, "unpack_bits: null pointer");
    PDB_REQUIRE(R > 0 && R <= 65535 && Ho > 0 && Wo > 0 && (Ho + 7) / 8 <= 65535, "unpack_bits: bad shape");
    const int Ww = (Wo + 31) / 32;
    dim3 grid((unsigned)Ww, (unsigned)((Ho + 7) / 8), (unsigned)R), block(32, 8);
    unpack_bits_kernel<<<grid, block, 0, as_stream(stream)>>>(bits, rows, out, Ho, Wo, Ww);
    return launched("unpack_bits");
}

extern "C" int pdb_bits_popcount(const uint32_t* bits, int64_t* counts, int rows, int64_t words, void* stream) {
    PDB_REQUIRE(bits && counts, "bits_popcount: null pointer");
    PDB_REQUIRE(rows > 0 && rows <= 65535 && words > 0, "bits_popcount: bad shape");
    dim3 grid((unsigned)((words + kChunkWords - 1) / kChunkWords), (unsigned)rows);
    bits_popcount_kernel<<<grid, 256, 0, as_stream(stream)>>>(bits, reinterpret_cast<unsigned long long*>(counts), words);
    return launched("bits_popcount");
}

extern "C" int pdb_bits_intersect(const uint32_t* a, const uint32_t* b, int64_t* inter, int Ka, int Kb, int64_t words,
                                  void* stream) {
    PDB_REQUIRE(a && b && inter, "bits_intersect: null pointer");
    PDB_REQUIRE(Ka > 0 && Ka <= 65535 && Kb > 0 && words > 0, "bits_intersect: bad shape");
    dim3 grid((unsigned)((words + kChunkWords - 1) / kChunkWords), (unsigned)Ka);
    bits_intersect_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, reinterpret_cast<unsigned long long*>(inter), Kb, words);
    return launched("bits_intersect");
}
