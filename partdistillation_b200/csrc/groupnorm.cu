// GroupNorm (+ ReLU) over channels-last maps, forward and backward.
//
// Replaces nn.GroupNorm(32, C) and the F.relu behind it on the pixel decoder's input_proj / lateral / output convolutions
// (reference: msdeformattn.py:249-287 through detectron2's Conv2d(norm=get_norm("GN", C), activation=F.relu)).  The
// convolutions of this library produce pixel-major (B, HW, C) maps (they are GEMMs over pixels); ATen's GroupNorm wants NCHW,
// so every call cost a transposing copy in, a 64-CTA statistics kernel (one CTA per (image, group): 0.4 ms on a 134 MB map)
// and a transposing copy back out for the next convolution.  Here the statistics pass runs on ~600 CTAs over the pixel-major
// rows (float4 = 4 channels of one group per thread, double accumulation, one f64 atomic per (CTA, group)), and the
// normalisation (+ReLU) is one more streaming pass in the same layout: 3 passes over the map instead of ~8.
//
// Thread mapping (all kernels): C4 = C / 4 float4 per pixel, 256 threads = (256 / C4) pixel rows x C4 channel quads, so a
// warp reads whole 128-byte row segments; a channel quad lies inside one group (C / G is a multiple of 4).
#include <algorithm>
#include "common.cuh"

namespace pdb {

constexpr int GN_THREADS = 256;

// y = (x - mean) * rstd * w + b, the one expression both the forward and the ReLU mask of the backward evaluate
__device__ __forceinline__ float gn_norm(float x, float mean, float rstd, float w, float b) { return (x - mean) * rstd * w + b; }

// sum / sum of squares per (image, group) -> stats[(b * G + g) * 2 + {0, 1}] (f64, zero-filled by the caller)
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const float4* __restrict__ x, double* __restrict__ stats, int64_t HW, int C4, int G, int quads_per_group,
                int64_t pixels_per_chunk) {
    __shared__ double red[2][GN_THREADS];
    const int b = blockIdx.y;
    const int cq = threadIdx.x % C4, r = threadIdx.x / C4, R = GN_THREADS / C4;
    const int64_t p0 = (int64_t)blockIdx.x * pixels_per_chunk;
    const int64_t p1 = min(HW, p0 + pixels_per_chunk);
    const float4* xb = x + (int64_t)b * HW * C4;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t p = p0 + r; p < p1; p += R) {
        const float4 v = __ldg(xb + p * C4 + cq);
        s1 += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
        s2 += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
    }
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < C4) {            // totals of channel quad t over the R rows
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < R; ++i) {
            t1 += red[0][i * C4 + threadIdx.x];
            t2 += red[1][i * C4 + threadIdx.x];
        }
        red[0][threadIdx.x] = t1;      // (only rows 0 of red are rewritten, by the thread that just finished reading that column)
        red[1][threadIdx.x] = t2;
    }
    __syncthreads();
    if (threadIdx.x < G) {
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < quads_per_group; ++i) {
            t1 += red[0][threadIdx.x * quads_per_group + i];
            t2 += red[1][threadIdx.x * quads_per_group + i];
        }
        atomicAdd(stats + ((int64_t)b * G + threadIdx.x) * 2, t1);
        atomicAdd(stats + ((int64_t)b * G + threadIdx.x) * 2 + 1, t2);
    }
}

__global__ void gn_finalize_kernel(const double* __restrict__ stats, float* __restrict__ mean, float* __restrict__ rstd, int n_groups,
                                   double inv_n, double eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups) return;
    const double m = stats[2 * i] * inv_n;
    const double var = fmax(stats[2 * i + 1] * inv_n - m * m, 0.0);
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + eps));
}

// y = GroupNorm(x) (ReLU).  Each thread keeps its channel quad (the grid stride is a multiple of C4).
template <bool RELU>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const float4* __restrict__ x, const float4* __restrict__ weight, const float4* __restrict__ bias,
                const float* __restrict__ mean, const float* __restrict__ rstd, float4* __restrict__ y, int64_t HW, int64_t pixels,
                int C4, int G, int quads_per_group) {
    const int64_t tid = (int64_t)blockIdx.x * GN_THREADS + threadIdx.x;
    const int cq = (int)(tid % C4);
    const int g = cq / quads_per_group;
    const float4 w = __ldg(weight + cq), bb = __ldg(bias + cq);
    const int64_t pstride = (int64_t)gridDim.x * GN_THREADS / C4;
    for (int64_t p = tid / C4; p < pixels; p += pstride) {
        const int64_t sg = (p / HW) * G + g;
        const float m = __ldg(mean + sg), rs = __ldg(rstd + sg);
        const float4 v = __ldg(x + p * C4 + cq);
        float4 o;
        o.x = gn_norm(v.x, m, rs, w.x, bb.x);
        o.y = gn_norm(v.y, m, rs, w.y, bb.y);
        o.z = gn_norm(v.z, m, rs, w.z, bb.z);
        o.w = gn_norm(v.w, m, rs, w.w, bb.w);
        if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        y[p * C4 + cq] = o;
    }
}

// per (image, channel): sum_hw dy' * xhat and sum_hw dy'  (dy' = dy, or dy where the forward output was positive)
template <bool RELU>
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_stats_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ weight,
                    const float4* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ rstd,
                    double* __restrict__ chan_sums, int64_t HW, int C4, int G, int quads_per_group, int64_t pixels_per_chunk) {
    __shared__ double red[8][GN_THREADS];
    const int b = blockIdx.y;
    const int cq = threadIdx.x % C4, r = threadIdx.x / C4, R = GN_THREADS / C4;
    const int g = cq / quads_per_group;
    const float m = __ldg(mean + (int64_t)b * G + g), rs = __ldg(rstd + (int64_t)b * G + g);
    const float4 w = __ldg(weight + cq), bb = __ldg(bias + cq);
    const int64_t p0 = (int64_t)blockIdx.x * pixels_per_chunk;
    const int64_t p1 = min(HW, p0 + pixels_per_chunk);
    const float4* xb = x + (int64_t)b * HW * C4;
    const float4* db = dy + (int64_t)b * HW * C4;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int64_t p = p0 + r; p < p1; p += R) {
        const float4 v = __ldg(xb + p * C4 + cq);
        float4 d = __ldg(db + p * C4 + cq);
        if (RELU) {
            if (!(gn_norm(v.x, m, rs, w.x, bb.x) > 0.f)) d.x = 0.f;
            if (!(gn_norm(v.y, m, rs, w.y, bb.y) > 0.f)) d.y = 0.f;
            if (!(gn_norm(v.z, m, rs, w.z, bb.z) > 0.f)) d.z = 0.f;
            if (!(gn_norm(v.w, m, rs, w.w, bb.w) > 0.f)) d.w = 0.f;
        }
        a0 += (double)(d.x * ((v.x - m) * rs)); s0 += (double)d.x;
        a1 += (double)(d.y * ((v.y - m) * rs)); s1 += (double)d.y;
        a2 += (double)(d.z * ((v.z - m) * rs)); s2 += (double)d.z;
        a3 += (double)(d.w * ((v.w - m) * rs)); s3 += (double)d.w;
    }
    red[0][threadIdx.x] = a0; red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2; red[3][threadIdx.x] = a3;
    red[4][threadIdx.x] = s0; red[5][threadIdx.x] = s1; red[6][threadIdx.x] = s2; red[7][threadIdx.x] = s3;
    __syncthreads();
    if (threadIdx.x < C4) {
        double* out = chan_sums + ((int64_t)b * C4 + threadIdx.x) * 8;      // channel c = 4 * quad + j: [c][0] = A, [c][1] = S
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double ta = 0.0, ts = 0.0;
            for (int i = 0; i < R; ++i) {
                ta += red[j][i * C4 + threadIdx.x];
                ts += red[4 + j][i * C4 + threadIdx.x];
            }
            atomicAdd(out + 2 * j, ta);
            atomicAdd(out + 2 * j + 1, ts);
        }
    }
}

// c1 = mean_group(dy' * w * xhat), c2 = mean_group(dy' * w) per (image, group)
__global__ void gn_bwd_group_kernel(const double* __restrict__ chan_sums, const float* __restrict__ weight, float* __restrict__ coef,
                                    int n_groups, int G, int C, int cpg, double inv_n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups) return;
    const int b = i / G, g = i % G;
    double c1 = 0.0, c2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
        const int c = g * cpg + j;
        const double wv = (double)weight[c];
        c1 += wv * chan_sums[((int64_t)b * C + c) * 2];
        c2 += wv * chan_sums[((int64_t)b * C + c) * 2 + 1];
    }
    coef[2 * i] = (float)(c1 * inv_n);
    coef[2 * i + 1] = (float)(c2 * inv_n);
}

// dx = rstd * (dy' * w - c2 - xhat * c1)
template <bool RELU>
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_apply_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ weight,
                    const float4* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ coef, float4* __restrict__ dx, int64_t HW, int64_t pixels, int C4, int G,
                    int quads_per_group) {
    const int64_t tid = (int64_t)blockIdx.x * GN_THREADS + threadIdx.x;
    const int cq = (int)(tid % C4);
    const int g = cq / quads_per_group;
    const float4 w = __ldg(weight + cq), bb = __ldg(bias + cq);
    const int64_t pstride = (int64_t)gridDim.x * GN_THREADS / C4;
    for (int64_t p = tid / C4; p < pixels; p += pstride) {
        const int64_t sg = (p / HW) * G + g;
        const float m = __ldg(mean + sg), rs = __ldg(rstd + sg);
        const float c1 = __ldg(coef + 2 * sg), c2 = __ldg(coef + 2 * sg + 1);
        const float4 v = __ldg(x + p * C4 + cq);
        float4 d = __ldg(dy + p * C4 + cq);
        if (RELU) {
            if (!(gn_norm(v.x, m, rs, w.x, bb.x) > 0.f)) d.x = 0.f;
            if (!(gn_norm(v.y, m, rs, w.y, bb.y) > 0.f)) d.y = 0.f;
            if (!(gn_norm(v.z, m, rs, w.z, bb.z) > 0.f)) d.z = 0.f;
            if (!(gn_norm(v.w, m, rs, w.w, bb.w) > 0.f)) d.w = 0.f;
        }
        float4 o;
        o.x = rs * (d.x * w.x - c2 - ((v.x - m) * rs) * c1);
        o.y = rs * (d.y * w.y - c2 - ((v.y - m) * rs) * c1);
        o.z = rs * (d.z * w.z - c2 - ((v.z - m) * rs) * c1);
        o.w = rs * (d.w * w.w - c2 - ((v.w - m) * rs) * c1);
        dx[p * C4 + cq] = o;
    }
}

static int gn_check(const char* what, int B, int64_t HW, int C, int G) {
    PDB_REQUIRE(B > 0 && HW > 0 && C > 0 && G > 0, "%s: non-positive dimension", what);
    PDB_REQUIRE(C % G == 0 && (C / G) % 4 == 0, "%s: channels per group (%d / %d) must be a multiple of 4", what, C, G);
    PDB_REQUIRE(C / 4 <= GN_THREADS && GN_THREADS % (C / 4) == 0, "%s: C / 4 = %d must divide %d", what, C / 4, GN_THREADS);
    PDB_REQUIRE(G <= GN_THREADS, "%s: at most %d groups", what, GN_THREADS);
    PDB_REQUIRE(B < 65536 && (int64_t)B * HW < (1ll << 40), "%s: batch / map too large", what);
    return PDB_OK;
}

// chunks of pixels per image so that ~4 CTAs per SM run; a chunk is a multiple of the rows one CTA covers per iteration
static void gn_chunks(int B, int64_t HW, int C4, int64_t& chunks, int64_t& pixels_per_chunk) {
    const int R = GN_THREADS / C4;
    const int64_t want = std::max<int64_t>(1, (4 * kNumSMs + B - 1) / B);
    pixels_per_chunk = ((HW + want - 1) / want + R - 1) / R * R;
    chunks = (HW + pixels_per_chunk - 1) / pixels_per_chunk;
}

static unsigned gn_apply_blocks(int64_t pixels, int C4) {
    const int64_t total = pixels * C4;
    return (unsigned)std::min<int64_t>((total + GN_THREADS - 1) / GN_THREADS, 8 * kNumSMs);
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_group_norm_forward(const float* x, const float* weight, const float* bias, float* y, double* stats, float* mean,
                                      float* rstd, int B, int64_t HW, int C, int G, float eps, int relu, void* stream) {
    PDB_REQUIRE(x && weight && bias && y && stats && mean && rstd, "group_norm_forward: null pointer");
    PDB_TRY(gn_check("group_norm_forward", B, HW, C, G));
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(weight) |
                  reinterpret_cast<uintptr_t>(bias)) & 15) == 0, "group_norm_forward: buffers must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int C4 = C / 4, qpg = (C / G) / 4;
    int64_t chunks, ppc;
    gn_chunks(B, HW, C4, chunks, ppc);
    gn_stats_kernel<<<dim3((unsigned)chunks, (unsigned)B), GN_THREADS, 0, st>>>(reinterpret_cast<const float4*>(x), stats, HW, C4, G,
                                                                              qpg, ppc);
    PDB_TRY(launched("group_norm_forward(stats)"));
    gn_finalize_kernel<<<(B * G + 127) / 128, 128, 0, st>>>(stats, mean, rstd, B * G, 1.0 / ((double)HW * (C / G)), (double)eps);
    PDB_TRY(launched("group_norm_forward(finalize)"));
    const int64_t pixels = (int64_t)B * HW;
    const unsigned blocks = gn_apply_blocks(pixels, C4);
    if (relu)
        gn_apply_kernel<true><<<blocks, GN_THREADS, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(weight),
                                                             reinterpret_cast<const float4*>(bias), mean, rstd,
                                                             reinterpret_cast<float4*>(y), HW, pixels, C4, G, qpg);
    else
        gn_apply_kernel<false><<<blocks, GN_THREADS, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(weight),
                                                              reinterpret_cast<const float4*>(bias), mean, rstd,
                                                              reinterpret_cast<float4*>(y), HW, pixels, C4, G, qpg);
    return launched("group_norm_forward(apply)");
}

extern "C" int pdb_group_norm_backward(const float* dy, const float* x, const float* weight, const float* bias, const float* mean,
                                       const float* rstd, double* chan_sums, float* coef, float* dx, int B, int64_t HW, int C, int G,
                                       int relu, void* stream) {
    PDB_REQUIRE(dy && x && weight && bias && mean && rstd && chan_sums && coef, "group_norm_backward: null pointer");
    PDB_TRY(gn_check("group_norm_backward", B, HW, C, G));
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                  reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                "group_norm_backward: buffers must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int C4 = C / 4, qpg = (C / G) / 4;
    int64_t chunks, ppc;
    gn_chunks(B, HW, C4, chunks, ppc);
    const dim3 sgrid((unsigned)chunks, (unsigned)B);
    const float4 *dy4 = reinterpret_cast<const float4*>(dy), *x4 = reinterpret_cast<const float4*>(x);
    const float4 *w4 = reinterpret_cast<const float4*>(weight), *b4 = reinterpret_cast<const float4*>(bias);
    if (relu)
        gn_bwd_stats_kernel<true><<<sgrid, GN_THREADS, 0, st>>>(dy4, x4, w4, b4, mean, rstd, chan_sums, HW, C4, G, qpg, ppc);
    else
        gn_bwd_stats_kernel<false><<<sgrid, GN_THREADS, 0, st>>>(dy4, x4, w4, b4, mean, rstd, chan_sums, HW, C4, G, qpg, ppc);
    PDB_TRY(launched("group_norm_backward(stats)"));
    if (!dx) return PDB_OK;
    gn_bwd_group_kernel<<<(B * G + 127) / 128, 128, 0, st>>>(chan_sums, weight, coef, B * G, G, C, C / G, 1.0 / ((double)HW * (C / G)));
    PDB_TRY(launched("group_norm_backward(group)"));
    const int64_t pixels = (int64_t)B * HW;
    const unsigned blocks = gn_apply_blocks(pixels, C4);
    if (relu)
        gn_bwd_apply_kernel<true><<<blocks, GN_THREADS, 0, st>>>(dy4, x4, w4, b4, mean, rstd, coef, reinterpret_cast<float4*>(dx), HW,
                                                                 pixels, C4, G, qpg);
    else
        gn_bwd_apply_kernel<false><<<blocks, GN_THREADS, 0, st>>>(dy4, x4, w4, b4, mean, rstd, coef, reinterpret_cast<float4*>(dx), HW,
                                                                  pixels, C4, G, qpg);
    return launched("group_norm_backward(apply)");
}
