// Pixel grouping: per-pixel part label = argmax over centroids of the affinity between the (bilinearly up-sampled)
// backbone feature and the k-means centroids — replaces, for every pixel of the object mask,
//   F.interpolate(features, image size, bilinear, align_corners=False)          pixel_grouping_model.py:139-144
//   feature[:, mask].T.contiguous().cpu();  measure_distance(...).topk(1)        :197-208
//   label map scatter                                                             :210-211
// of PixelGroupingModel.generate_part_segments (the reference moves the full-resolution (C, H, W) feature map to the
// host for this).  The feature map stays at backbone resolution in HBM/L2 (C*h*w*4 bytes instead of C*H*W*4) and is
// interpolated on the fly, ATen's align_corners=False source-index rule (src = (dst + 0.5) * in/out - 0.5, clamped
// at 0; second tap clamped to in - 1).
//   dot:  score_k = <f, c_k>          l2:  score_k = 2 <f, c_k> - |c_k|^2   (the -|f|^2 term does not change argmax)
// One thread per output pixel; the C x Kc centroid table sits in shared memory (broadcast reads); neighbouring
// pixels share their 4 source texels, so the feature reads are L1 hits after the first touch.
#include "common.cuh"
#include "grouping_resized.cuh"

namespace pdb {

constexpr int kMaxCentroids = 16;

__global__ void __launch_bounds__(256)
group_affinity_kernel(const float* __restrict__ feat, const float* __restrict__ centroids, const uint8_t* __restrict__ mask,
                      int32_t* __restrict__ labels, int C, int Kc, int h, int w, int H, int W, int l2, int64_t cent_stride) {
    extern __shared__ float s_cent[];       // [C][Kc] + |c_k|^2 [Kc]
    float* s_norm = s_cent + C * Kc;
    feat += (int64_t)blockIdx.z * C * h * w;                // blockIdx.z = image of the batch (cent_stride 0: shared centroids)
    centroids += (int64_t)blockIdx.z * cent_stride;
    mask += (int64_t)blockIdx.z * H * W;
    labels += (int64_t)blockIdx.z * H * W;
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
    if (x >= W || y >= H) return;
    const int64_t pix = (int64_t)y * W + x;
    if (!mask[pix]) {
        labels[pix] = 0;
        return;
    }
    const float sy = fmaxf(((float)y + 0.5f) * ((float)h / (float)H) - 0.5f, 0.f);
    const float sx = fmaxf(((float)x + 0.5f) * ((float)w / (float)W) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly1 = sy - (float)y0, lx1 = sx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const int o00 = y0 * w + x0, o01 = y0 * w + x1, o10 = y1 * w + x0, o11 = y1 * w + x1;
    float score[kMaxCentroids];
#pragma unroll
    for (int k = 0; k < kMaxCentroids; ++k) score[k] = 0.f;
    const int plane = h * w;
    for (int c = 0; c < C; ++c) {
        const float* fp = feat + (int64_t)c * plane;
        // ATen's upsample_bilinear2d: h0lambda * (w0lambda * a + w1lambda * b) + h1lambda * (w0lambda * c + w1lambda * d)
        const float v = ly0 * (lx0 * __ldg(fp + o00) + lx1 * __ldg(fp + o01)) + ly1 * (lx0 * __ldg(fp + o10) + lx1 * __ldg(fp + o11));
        const float* cc = s_cent + c * Kc;
#pragma unroll
        for (int k = 0; k < kMaxCentroids; ++k)
            if (k < Kc) score[k] = fmaf(v, cc[k], score[k]);
    }
    int best = 0;
    float bv = l2 ? 2.f * score[0] - s_norm[0] : score[0];
#pragma unroll
    for (int k = 1; k < kMaxCentroids; ++k) {
        if (k < Kc) {
            const float s = l2 ? 2.f * score[k] - s_norm[k] : score[k];
            if (s > bv) { bv = s; best = k; }
        }
    }
    labels[pix] = best + 1;
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_group_affinity(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels, int C,
                                  int Kc, int h, int w, int H, int W, int metric, void* stream) {
    PDB_REQUIRE(feat && centroids && mask && labels, "group_affinity: null pointer");
    PDB_REQUIRE(C > 0 && Kc > 0 && Kc <= kMaxCentroids && h > 0 && w > 0 && H > 0 && W > 0,
                "group_affinity: bad sizes (1 <= Kc <= %d)", kMaxCentroids);
    PDB_REQUIRE(metric == 0 || metric == 1, "group_affinity: metric %d (0 = dot, 1 = l2)", metric);
    size_t smem = sizeof(float) * ((size_t)C * Kc + Kc);
    PDB_REQUIRE(smem <= 200 * 1024, "group_affinity: centroid table of %zu bytes does not fit shared memory", smem);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(group_affinity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "group_affinity: smem attribute: %s", cudaGetErrorString(e));
        attr_smem = smem;
    }
    dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8));
    group_affinity_kernel<<<grid, 256, smem, as_stream(stream)>>>(feat, centroids, mask, labels, C, Kc, h, w, H, W, metric, 0);
    return launched("group_affinity");
}

// B images of one geometry in two launches: score maps at feature resolution (pdb_group_scores per image, grid.y = B), then
// the per-pixel interpolation + argmax on the Kc score maps (identity centroids).  scores: workspace (B, Kc, h, w) floats.
extern "C" int pdb_group_affinity_batched(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels,
                                          float* scores, const float* identity, int B, int C, int Kc, int h, int w, int H, int W,
                                          int metric, void* stream) {
    PDB_REQUIRE(feat && centroids && mask && labels && scores && identity, "group_affinity_batched: null pointer");
    PDB_REQUIRE(B > 0 && B <= 65535 && C > 0 && Kc > 0 && Kc <= kMaxCentroids && h > 0 && w > 0 && H > 0 && W > 0,
                "group_affinity_batched: bad sizes (1 <= Kc <= %d, B <= 65535)", kMaxCentroids);
    PDB_REQUIRE(metric == 0 || metric == 1, "group_affinity_batched: metric %d (0 = dot, 1 = l2)", metric);
    const int hw = h * w;
    group_scores_kernel<<<dim3((unsigned)((hw + 31) / 32), (unsigned)B), 256, 0, as_stream(stream)>>>(feat, centroids, scores, C, Kc,
                                                                                                     hw, metric);
    PDB_TRY(launched("group_scores"));
    const size_t smem = sizeof(float) * ((size_t)Kc * Kc + Kc);
    dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)B);
    group_affinity_kernel<<<grid, 256, smem, as_stream(stream)>>>(scores, identity, mask, labels, Kc, Kc, h, w, H, W, 0, 0);
    return launched("group_affinity");
}

extern "C" int pdb_group_scores(const float* feat, const float* centroids, float* scores, int C, int Kc, int h, int w, int metric,
                                void* stream) {
    PDB_REQUIRE(feat && centroids && scores, "group_scores: null pointer");
    PDB_REQUIRE(C > 0 && Kc > 0 && Kc <= kMaxGroupCentroids && h > 0 && w > 0, "group_scores: bad sizes (1 <= Kc <= %d)",
                kMaxGroupCentroids);
    PDB_REQUIRE(metric == 0 || metric == 1, "group_scores: metric %d (0 = dot, 1 = l2)", metric);
    const int hw = h * w;
    group_scores_kernel<<<(unsigned)((hw + 31) / 32), 256, 0, as_stream(stream)>>>(feat, centroids, scores, C, Kc, hw, metric);
    return launched("group_scores");
}

// General geometry: features (C, h, w) -> bilinear -> padded (Hp, Wp) -> crop (Hi, Wi) -> bilinear -> (Ho, Wo); mask and
// labels at (Ho, Wo).  With Hp == Hi == Ho and Wp == Wi == Wo this computes what pdb_group_affinity computes (up to the
// FMA contraction the older kernel leaves to the compiler).
extern "C" int pdb_group_affinity_resized(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels,
                                          int C, int Kc, int h, int w, int Hp, int Wp, int Hi, int Wi, int Ho, int Wo,
                                          int metric, void* stream) {
    PDB_REQUIRE(feat && centroids && mask && labels, "group_affinity_resized: null pointer");
    PDB_REQUIRE(C > 0 && Kc > 0 && Kc <= kMaxGroupCentroids && h > 0 && w > 0 && Hp > 0 && Wp > 0 && Ho > 0 && Wo > 0,
                "group_affinity_resized: bad sizes (1 <= Kc <= %d)", kMaxGroupCentroids);
    PDB_REQUIRE(Hi > 0 && Wi > 0 && Hi <= Hp && Wi <= Wp, "group_affinity_resized: image size outside the padded size");
    PDB_REQUIRE(metric == 0 || metric == 1, "group_affinity_resized: metric %d (0 = dot, 1 = l2)", metric);
    PDB_REQUIRE((Ho + 7) / 8 <= 65535, "group_affinity_resized: output height %d exceeds grid.y", Ho);
    const size_t smem = group_affinity_smem(C, Kc);
    PDB_REQUIRE(smem <= 200 * 1024, "group_affinity_resized: centroid table of %zu bytes does not fit shared memory", smem);
    const bool two = !(Hi == Ho && Wi == Wo);
    static size_t attr_smem[2] = {0, 0};
    if (smem > 48 * 1024 && smem > attr_smem[two]) {
        cudaError_t e = two ? cudaFuncSetAttribute(group_affinity_resized_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(group_affinity_resized_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PDB_ERR_LAUNCH, "group_affinity_resized: smem attribute: %s", cudaGetErrorString(e));
        attr_smem[two] = smem;
    }
    const float s1h = (float)h / (float)Hp, s1w = (float)w / (float)Wp;
    const float s2h = (float)Hi / (float)Ho, s2w = (float)Wi / (float)Wo;
    const dim3 grid = group_affinity_grid(Ho, Wo);
    if (two)
        group_affinity_resized_kernel<true><<<grid, 256, smem, as_stream(stream)>>>(feat, centroids, mask, labels, C, Kc, h, w, Hi, Wi,
                                                                                  Ho, Wo, s1h, s1w, s2h, s2w, metric);
    else
        group_affinity_resized_kernel<false><<<grid, 256, smem, as_stream(stream)>>>(feat, centroids, mask, labels, C, Kc, h, w, Hi, Wi,
                                                                                   Ho, Wo, s1h, s1w, s2h, s2w, metric);
    return launched("group_affinity_resized");
}
