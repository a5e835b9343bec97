// Pixel grouping at an evaluation size that differs from the padded batch size: the reference up-samples the backbone
// features to the padded size, crops the padding and resizes again to (height, width) (pixel_grouping_model.py:139-160 /
// proposal_generation_model.py:139-155 via detectron2 sem_seg_postprocess) before measure_distance + topk.  Both bilinear
// passes are composed per output pixel with the helpers of postprocess_math.cuh (same arithmetic as the mask
// post-processing); everything else is group_affinity_kernel (grouping.cu).  Header-only so the tests can compile the
// kernel for the host (tests/native/cuda_on_cpu.h).
#pragma once
#include "postprocess_math.cuh"

#ifndef PDB_DYNAMIC_SMEM
#define PDB_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif

namespace pdb {

constexpr int kMaxGroupCentroids = 16;

// Stage 1 of the grouping (linearity of the bilinear up-sampling): the affinity of an up-sampled feature with a centroid is
// the up-sampled affinity, sum_c bilinear(f_c) c_kc = bilinear(sum_c f_c c_kc), so the contraction over the C channels is
// done ONCE at feature resolution (C h w reads instead of 4 C reads per output pixel) and the per-pixel kernels interpolate
// Kc score maps.  scores[k][p] = <f_p, c_k> (dot) or 2 <f_p, c_k> - |c_k|^2 (l2; interpolation weights sum to 1, so the
// constant passes through both bilinear passes unchanged).
// grid ceil(h w / 32), block 256 = 32 pixels x 8 channel groups (channel c belongs to group c % 8: a warp reads 32
// consecutive pixels of one channel plane = one 128-byte line); the 8 partial sums of a pixel are reduced in shared memory.
__global__ void __launch_bounds__(256)
group_scores_kernel(const float* __restrict__ feat, const float* __restrict__ centroids, float* __restrict__ scores, int C,
                    int Kc, int hw, int l2) {
    __shared__ float s_part[8][kMaxGroupCentroids][33];
    feat += (int64_t)blockIdx.y * C * hw;                  // blockIdx.y = image of the batch
    centroids += (int64_t)blockIdx.y * Kc * C;
    scores += (int64_t)blockIdx.y * Kc * hw;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    float acc[kMaxGroupCentroids], nrm[kMaxGroupCentroids];
#pragma unroll
    for (int k = 0; k < kMaxGroupCentroids; ++k) acc[k] = nrm[k] = 0.f;
    if (p < hw) {
        for (int c = grp; c < C; c += 8) {
            const float v = __ldg(feat + (int64_t)c * hw + p);
#pragma unroll
            for (int k = 0; k < kMaxGroupCentroids; ++k)
                if (k < Kc) {
                    const float ck = __ldg(centroids + k * C + c);
                    acc[k] = fmaf(v, ck, acc[k]);
                    nrm[k] = fmaf(ck, ck, nrm[k]);
                }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxGroupCentroids; ++k)
        if (k < Kc) s_part[grp][k][lane] = l2 ? 2.f * acc[k] - nrm[k] : acc[k];
    __syncthreads();
    // thread (grp, lane): centroids grp, grp + 8 of pixel lane
    for (int k = grp; k < Kc; k += 8) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += s_part[g][k][lane];
        if (p < hw) scores[(int64_t)k * hw + p] = s;
    }
}

// grid (ceil(Wo / 32), ceil(Ho / 8)), block 256, dynamic shared memory (C * Kc + Kc) floats
template <bool TWO_STAGE>
__global__ void __launch_bounds__(256)
group_affinity_resized_kernel(const float* __restrict__ feat, const float* __restrict__ centroids,
                              const uint8_t* __restrict__ mask, int32_t* __restrict__ labels, int C, int Kc, int h, int w,
                              int Hi, int Wi, int Ho, int Wo, float s1h, float s1w, float s2h, float s2w, int l2) {
    PDB_DYNAMIC_SMEM(float, s_cent);        // [C][Kc] + |c_k|^2 [Kc]
    float* s_norm = s_cent + C * Kc;
    for (int i = threadIdx.x; i < C * Kc; i += blockDim.x) {
        const int c = i / Kc, k = i - c * Kc;
        s_cent[i] = __ldg(centroids + k * C + c);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kc; k += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(s_cent[c * Kc + k], s_cent[c * Kc + k], s);
        s_norm[k] = s;
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= Wo || y >= Ho) return;
    const int64_t pix = (int64_t)y * Wo + x;
    if (!mask[pix]) {
        labels[pix] = 0;
        return;
    }
    const PixelTaps taps = make_pixel_taps<TWO_STAGE>(y, x, h, w, Hi, Wi, s1h, s1w, s2h, s2w);
    float score[kMaxGroupCentroids];
#pragma unroll
    for (int k = 0; k < kMaxGroupCentroids; ++k) score[k] = 0.f;
    const int64_t plane = (int64_t)h * w;
    for (int c = 0; c < C; ++c) {
        const float v = sample_pixel<TWO_STAGE>(feat + c * plane, taps);
        const float* cc = s_cent + c * Kc;
#pragma unroll
        for (int k = 0; k < kMaxGroupCentroids; ++k)
            if (k < Kc) score[k] = fmaf(v, cc[k], score[k]);
    }
    int best = 0;
    float bv = l2 ? 2.f * score[0] - s_norm[0] : score[0];
#pragma unroll
    for (int k = 1; k < kMaxGroupCentroids; ++k) {
        if (k < Kc) {
            const float s = l2 ? 2.f * score[k] - s_norm[k] : score[k];
            if (s > bv) { bv = s; best = k; }
        }
    }
    labels[pix] = best + 1;
}

inline dim3 group_affinity_grid(int Ho, int Wo) { return dim3((unsigned)((Wo + 31) / 32), (unsigned)((Ho + 7) / 8), 1); }
inline size_t group_affinity_smem(int C, int Kc) { return sizeof(float) * ((size_t)C * Kc + Kc); }

}  // namespace pdb
