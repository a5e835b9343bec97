// Kernels of the inference post-processing (launched by postprocess.cu).  They live in a header so that
// tests/native/postprocess_kernels_host.cpp can compile the SAME kernel bodies for the host through a small
// CUDA-threading shim (tests/native/cuda_on_cpu.h: one OS thread per CUDA thread, barriers for the warp / block
// collectives) and run them against the oracle where there is no GPU.  Launch geometry is defined here too, next to
// the indexing that depends on it.
#pragma once
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


// ---- launch geometry (shared by the launchers in postprocess.cu and the host harness of the tests)
inline dim3 row_word_block() { return dim3(32, 8, 1); }                                   // a warp = one packed word
inline dim3 row_word_grid(int Ho, int Wo, int R) { return dim3((unsigned)((Wo + 31) / 32), (unsigned)((Ho + 7) / 8), (unsigned)R); }
inline dim3 per_pixel_grid(int Ho, int Wo, int G) { return dim3((unsigned)(((int64_t)Ho * Wo + 255) / 256), (unsigned)G, 1); }
inline dim3 chunk_grid(int64_t words, int rows) { return dim3((unsigned)((words + kChunkWords - 1) / kChunkWords), (unsigned)rows, 1); }

}  // namespace pdb
