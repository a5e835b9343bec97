// Loss-side kernels: point sampling, Hungarian-matcher cost matrices, batched LSAP, the fused
// point-sampled BCE + dice mask loss, and PartDistillation's gathered-row fp64 classifier.
// Reference call sites: matcher.py:100-168, criterion.py:147-207 (detectron2 point_sample ==
// F.grid_sample(2*coords-1, bilinear, zeros, align_corners=False)), scipy linear_sum_assignment,
// part_distillation_transformer_decoder.py:107,215-238.
#include "common.cuh"
#include <math.h>

namespace pdb {

constexpr int kMaxBatch = 255;
struct Offsets { int v[kMaxBatch + 1]; };

// bilinear tap of a point (cx, cy) in [0,1]^2 on an H x W map, grid_sample(align_corners=False)
struct Bilin {
    int off[4];
    float w[4];   // 0 when the corner is outside the map
    int px[4], py[4];   // clamped corner coordinates (the bit-packed sources address words, not bytes)
};
__device__ __forceinline__ Bilin bilin_setup(float cx, float cy, int H, int W) {
    float gx = __fsub_rn(__fmul_rn(2.f, cx), 1.f);
    float gy = __fsub_rn(__fmul_rn(2.f, cy), 1.f);
    float x = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
    float y = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
    x = fminf(fmaxf(x, -2.f), (float)W + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)H + 1.f);
    float x0f = floorf(x), y0f = floorf(y);
    float wx1 = x - x0f, wy1 = y - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    int x0 = (int)x0f, y0 = (int)y0f;
    Bilin r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int xi = x0 + (c & 1), yi = y0 + (c >> 1);
        bool ok = xi >= 0 && xi < W && yi >= 0 && yi < H;
        r.px[c] = min(max(xi, 0), W - 1);
        r.py[c] = min(max(yi, 0), H - 1);
        r.off[c] = r.py[c] * W + r.px[c];
        r.w[c] = ok ? ((c & 1) ? wx1 : wx0) * ((c >> 1) ? wy1 : wy0) : 0.f;
    }
    return r;
}
__device__ __forceinline__ float sample_f32(const float* __restrict__ m, const Bilin& t) {
    // same association as ATen: nw + ne + sw + se
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (t.w[c] != 0.f) s += __ldg(m + t.off[c]) * t.w[c];
    return s;
}
__device__ __forceinline__ float sample_u8(const uint8_t* __restrict__ m, const Bilin& t) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (t.w[c] != 0.f) s += (__ldg(m + t.off[c]) ? 1.f : 0.f) * t.w[c];
    return s;
}

// maps at one bit per pixel: int32 words (rows x ceil(W / 32)), bit x % 32 of word x / 32 (the layout of pdb_pack_bits): the
// ground-truth masks are sampled straight from the words they arrived in
__device__ __forceinline__ float sample_bits(const uint32_t* __restrict__ m, const Bilin& t, int Ww) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (t.w[c] != 0.f) s += (float)((__ldg(m + t.py[c] * Ww + (t.px[c] >> 5)) >> (t.px[c] & 31)) & 1u) * t.w[c];
    return s;
}
__device__ __forceinline__ float sample_gt(const void* __restrict__ gt, int64_t gi, const Bilin& t, int Hg, int Wg, int bits) {
    if (bits) {
        const int Ww = (Wg + 31) >> 5;
        return sample_bits((const uint32_t*)gt + gi * Hg * Ww, t, Ww);
    }
    return sample_u8((const uint8_t*)gt + gi * Hg * Wg, t);
}

__global__ void point_sample_fwd(const void* __restrict__ src, int src_u8, const int32_t* __restrict__ map_index,
                                 const float* __restrict__ coords, const int32_t* __restrict__ coord_index,
                                 float* __restrict__ out, int R, int P, int H, int W) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)R * P) return;
    int r = (int)(idx / P), p = (int)(idx - (int64_t)r * P);
    int mp = map_index ? map_index[r] : r;
    int cr = coord_index ? coord_index[r] : r;
    float v = 0.f;
    if (mp >= 0) {
        const float* c = coords + ((int64_t)cr * P + p) * 2;
        Bilin t = bilin_setup(__ldg(c), __ldg(c + 1), H, W);
        v = src_u8 ? sample_gt(src, mp, t, H, W, src_u8 == 2)       // src_u8: 0 = f32, 1 = uint8, 2 = bit-packed words
                   : sample_f32((const float*)src + (int64_t)mp * H * W, t);
    }
    out[idx] = v;
}

__global__ void point_sample_bwd(const float* __restrict__ gout, const int32_t* __restrict__ map_index,
                                 const float* __restrict__ coords, const int32_t* __restrict__ coord_index,
                                 float* __restrict__ gsrc, int R, int P, int H, int W) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)R * P) return;
    int r = (int)(idx / P), p = (int)(idx - (int64_t)r * P);
    int mp = map_index ? map_index[r] : r;
    if (mp < 0) return;
    int cr = coord_index ? coord_index[r] : r;
    const float* c = coords + ((int64_t)cr * P + p) * 2;
    Bilin t = bilin_setup(__ldg(c), __ldg(c + 1), H, W);
    float g = gout[idx];
    float* m = gsrc + (int64_t)mp * H * W;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (t.w[k] != 0.f) atomicAdd(m + t.off[k], g * t.w[k]);
}

__device__ __forceinline__ float softplus_f(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + expf(-z)); }

// block-wide sum of one value; result valid in thread 0.  `red` has >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
    }
    return t;
}

// one CTA per (image, query): cost row over that image's targets
constexpr int CK = 8;   // targets per pass
__global__ void __launch_bounds__(512)
matcher_cost_kernel(const float* __restrict__ pred_pts, const float* __restrict__ tgt_pts,
                    const float* __restrict__ cls_prob, const int32_t* __restrict__ tgt_label, Offsets off,
                    float* __restrict__ cost, int Q, int Kc, int P, float w_class, float w_mask, float w_dice) {
    __shared__ float red[32];
    const int q = blockIdx.x, b = blockIdx.y;
    const int k0 = off.v[b], Kb = off.v[b + 1] - k0;
    if (Kb <= 0) return;
    const float* pr = pred_pts + ((int64_t)b * Q + q) * P;
    float* crow = cost + (int64_t)Q * k0 + (int64_t)q * Kb;
    for (int kb = 0; kb < Kb; kb += CK) {
        const int nk = min(CK, Kb - kb);
        float am[CK], ast[CK], at[CK], as = 0.f;
#pragma unroll
        for (int k = 0; k < CK; ++k) { am[k] = 0.f; ast[k] = 0.f; at[k] = 0.f; }
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            float o = __ldg(pr + p);
            float pos = softplus_f(-o), neg = softplus_f(o), s = sigmoid_f(o);
            as += s;
#pragma unroll
            for (int k = 0; k < CK; ++k) {
                if (k < nk) {
                    float t = __ldg(tgt_pts + (int64_t)(k0 + kb + k) * P + p);
                    am[k] += pos * t + neg * (1.f - t);
                    ast[k] = fmaf(s, t, ast[k]);
                    at[k] += t;
                }
            }
        }
        float S = block_sum(as, red);
#pragma unroll
        for (int k = 0; k < CK; ++k) {
            if (k >= nk) break;     // uniform across the block
            float m = block_sum(am[k], red);
            float st = block_sum(ast[k], red);
            float T = block_sum(at[k], red);
            if (threadIdx.x == 0) {
                float cmask = m / (float)P;
                float cdice = 1.f - (2.f * st + 1.f) / (S + T + 1.f);
                // label < 0: a padding slot (the trainer pads every image's targets to a multiple of its bucket so that batches with
                // different target counts replay one CUDA graph).  Its column costs the same for every query, so the optimal
                // assignment of the real targets is the one of the unpadded problem (K_padded <= Q: a free query always exists).
                const int lab = tgt_label[k0 + kb + k];
                float ccls = lab < 0 ? 0.f : -__ldg(cls_prob + ((int64_t)b * Q + q) * Kc + lab);
                crow[kb + k] = lab < 0 ? 0.f : w_mask * cmask + w_class * ccls + w_dice * cdice;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// batched rectangular LSAP (shortest augmenting path, float64, SciPy tie-breaking), one warp/image
// ------------------------------------------------------------------------------------------------
constexpr int LS_MAXR = 256;    // min(Q, K)
constexpr int LS_MAXC = 1024;   // max(Q, K)

struct Cand { double v; int first; int lastu; };
__device__ __forceinline__ Cand cand_merge(Cand a, Cand b) {
    if (b.v < a.v) return b;
    if (a.v < b.v) return a;
    Cand r;
    r.v = a.v;
    r.first = (a.first < 0) ? b.first : ((b.first < 0) ? a.first : min(a.first, b.first));
    r.lastu = max(a.lastu, b.lastu);
    return r;
}

__global__ void __launch_bounds__(32)
lsap_kernel(const float* __restrict__ cost, Offsets off, int64_t* __restrict__ pred_idx, int64_t* __restrict__ tgt_idx,
            int Q) {
    __shared__ double u[LS_MAXR], v[LS_MAXC], shortest[LS_MAXC];
    __shared__ int path[LS_MAXC], row4col[LS_MAXC], remaining[LS_MAXC], col4row[LS_MAXR];
    __shared__ unsigned char SR[LS_MAXR], SC[LS_MAXC];
    __shared__ float pc[LS_MAXR];
    __shared__ int prow[LS_MAXR], pcol[LS_MAXR];
    const int b = blockIdx.x, lane = threadIdx.x;
    const int k0 = off.v[b], K = off.v[b + 1] - k0;
    if (K <= 0) return;
    const float* C = cost + (int64_t)Q * k0;            // (Q, K) row-major
    const bool tr = Q > K;                               // SciPy transposes when rows > cols
    const int nr = tr ? K : Q, nc = tr ? Q : K;
    // element (i, j) of the working matrix
    auto at = [&](int i, int j) -> double { return (double)(tr ? C[(int64_t)j * K + i] : C[(int64_t)i * K + j]); };

    for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = lane; j < nc; j += 32) { v[j] = 0.0; row4col[j] = -1; }
    __syncwarp();
    for (int cur = 0; cur < nr; ++cur) {
        for (int j = lane; j < nc; j += 32) { shortest[j] = INFINITY; path[j] = -1; SC[j] = 0; remaining[j] = nc - j - 1; }
        for (int i = lane; i < nr; i += 32) SR[i] = 0;
        __syncwarp();
        int num_remaining = nc, i = cur, sink = -1;
        double min_val = 0.0;
        while (sink == -1) {
            if (lane == 0) SR[i] = 1;
            const double ui = u[i];
            Cand best; best.v = INFINITY; best.first = -1; best.lastu = -1;
            for (int it = lane; it < num_remaining; it += 32) {
                int j = remaining[it];
                double r = min_val + at(i, j) - ui - v[j];
                double sj = shortest[j];
                if (r < sj) { path[j] = i; shortest[j] = r; sj = r; }
                Cand c; c.v = sj; c.first = it; c.lastu = (row4col[j] == -1) ? it : -1;
                best = cand_merge(best, c);     // positions visited in increasing order per lane
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Cand other;
                other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
                other.first = __shfl_xor_sync(0xffffffffu, best.first, o);
                other.lastu = __shfl_xor_sync(0xffffffffu, best.lastu, o);
                best = cand_merge(best, other);
            }
            int index = best.lastu >= 0 ? best.lastu : best.first;
            if (index < 0) index = 0;            // NaN / inf costs: keep the loop finite
            min_val = best.v;
            __syncwarp();
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            __syncwarp();
            if (lane == 0) { SC[j] = 1; remaining[index] = remaining[num_remaining - 1]; }
            --num_remaining;
            __syncwarp();
        }
        if (lane == 0) u[cur] += min_val;
        for (int r = lane; r < nr; r += 32)
            if (SR[r] && r != cur) u[r] += min_val - shortest[col4row[r]];
        for (int j = lane; j < nc; j += 32)
            if (SC[j]) v[j] -= min_val - shortest[j];
        __syncwarp();
        if (lane == 0) {
            int j = sink;
            while (true) {
                int ii = path[j];
                row4col[j] = ii;
                int t = col4row[ii];
                col4row[ii] = j;
                j = t;
                if (ii == cur) break;
            }
        }
        __syncwarp();
    }
    // pairs in SciPy's return order: sorted by the row index of the ORIGINAL (Q, K) matrix
    if (!tr) {
        for (int q = lane; q < nr; q += 32) { prow[q] = q; pcol[q] = col4row[q]; }
    } else {
        for (int k = lane; k < nr; k += 32) {
            int qk = col4row[k], rank = 0;
            for (int k2 = 0; k2 < nr; ++k2) rank += (col4row[k2] < qk) ? 1 : 0;
            prow[rank] = qk; pcol[rank] = k;
        }
    }
    __syncwarp();
    for (int p = lane; p < nr; p += 32) pc[p] = C[(int64_t)prow[p] * K + pcol[p]];
    __syncwarp();
    // ascending matched cost (matcher.py:162-163), stable in SciPy order
    for (int p = lane; p < nr; p += 32) {
        float c = pc[p];
        int rank = 0;
        for (int p2 = 0; p2 < nr; ++p2) rank += (pc[p2] < c || (pc[p2] == c && p2 < p)) ? 1 : 0;
        pred_idx[k0 + rank] = prow[p];
        tgt_idx[k0 + rank] = pcol[p];
    }
    for (int p = nr + lane; p < K; p += 32) { pred_idx[k0 + p] = -1; tgt_idx[k0 + p] = -1; }
}

// ------------------------------------------------------------------------------------------------
// fused point-sampled BCE + dice
// ------------------------------------------------------------------------------------------------
// one CTA of 1024 threads per matched pair: there are only a handful of pairs per layer, so the block is as wide as it
// can be (12544 points -> 12 per thread)
// With few pairs (B = 2: 12 per decoder output) the points of a pair are additionally split over blockIdx.y (`chunk` points per
// CTA); the partial sums go to sums[(i * gridDim.y + s) * 4 ..] and point_loss_fwd_final adds them in split order (deterministic).
__global__ void __launch_bounds__(1024)
point_loss_fwd(const float* __restrict__ pred, const int64_t* __restrict__ pred_index, const void* __restrict__ gt,
               const int64_t* __restrict__ gt_index, const float* __restrict__ coords, float* __restrict__ sums, int P,
               int H, int W, int Hg, int Wg, int gt_bits, int chunk) {
    __shared__ float red[32];
    const int i = blockIdx.x;
    const int64_t pi = pred_index[i], gi = gt_index[i];
    float* out = sums + ((int64_t)i * gridDim.y + blockIdx.y) * 4;
    if (pi < 0 || gi < 0) {
        if (threadIdx.x < 4) out[threadIdx.x] = 0.f;
        return;
    }
    const float* pm = pred + pi * H * W;
    float a_bce = 0.f, a_st = 0.f, a_s = 0.f, a_t = 0.f;
    const int p_end = min(P, (int)(blockIdx.y + 1) * chunk);
    for (int p = blockIdx.y * chunk + threadIdx.x; p < p_end; p += blockDim.x) {
        const float* c = coords + ((int64_t)i * P + p) * 2;
        float cx = __ldg(c), cy = __ldg(c + 1);
        float x = sample_f32(pm, bilin_setup(cx, cy, H, W));
        float t = sample_gt(gt, gi, bilin_setup(cx, cy, Hg, Wg), Hg, Wg, gt_bits);
        // ATen binary_cross_entropy_with_logits: (1-t)*x + max(-x,0) + log(1+exp(-|x|))
        a_bce += (1.f - t) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
        float s = sigmoid_f(x);
        a_st = fmaf(s, t, a_st);
        a_s += s;
        a_t += t;
    }
    float r0 = block_sum(a_bce, red), r1 = block_sum(a_st, red), r2 = block_sum(a_s, red), r3 = block_sum(a_t, red);
    if (threadIdx.x == 0) {
        out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3;
    }
}

__global__ void point_loss_fwd_final(const float* __restrict__ partial, float* __restrict__ sums, int n4, int S) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;       // (pair, component)
    if (e >= n4) return;
    const int i = e >> 2, k = e & 3;
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += partial[((int64_t)i * S + s) * 4 + k];
    sums[e] = a;
}

__global__ void __launch_bounds__(1024)
point_loss_bwd(const float* __restrict__ pred, const int64_t* __restrict__ pred_index, const void* __restrict__ gt,
               const int64_t* __restrict__ gt_index, const float* __restrict__ coords, const float* __restrict__ sums,
               const float* __restrict__ g_bce, const float* __restrict__ g_dice, float* __restrict__ gpred, int P,
               int H, int W, int Hg, int Wg, int gt_bits, int chunk) {
    const int i = blockIdx.x;
    const int64_t pi = pred_index[i], gi = gt_index[i];
    if (pi < 0 || gi < 0) return;
    const float* pm = pred + pi * H * W;
    float* gp = gpred + pi * H * W;
    const float st = sums[i * 4 + 1], S = sums[i * 4 + 2], T = sums[i * 4 + 3];
    const float den = S + T + 1.f, num = 2.f * st + 1.f;
    const float gb = g_bce[i] / (float)P, gd = g_dice[i];
    const int p_end = min(P, (int)(blockIdx.y + 1) * chunk);
    for (int p = blockIdx.y * chunk + threadIdx.x; p < p_end; p += blockDim.x) {
        const float* c = coords + ((int64_t)i * P + p) * 2;
        float cx = __ldg(c), cy = __ldg(c + 1);
        Bilin tp = bilin_setup(cx, cy, H, W);
        float x = sample_f32(pm, tp);
        float t = sample_gt(gt, gi, bilin_setup(cx, cy, Hg, Wg), Hg, Wg, gt_bits);
        float s = sigmoid_f(x);
        // d dice / d s_p = -(2 t den - num) / den^2 ; d s / d x = s (1 - s)
        float dx = gb * (s - t) + gd * (-(2.f * t * den - num) / (den * den)) * s * (1.f - s);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (tp.w[k] != 0.f) atomicAdd(gp + tp.off[k], dx * tp.w[k]);
    }
}

// ------------------------------------------------------------------------------------------------
// PartDistillation classifier: only P+1 rows of the (P*O+1, C) fp64 weight are used per image
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t class_row(const int32_t* obj, int b, int j, int Pn, int64_t Ncls) {
    return j < Pn ? (int64_t)obj[b] * Pn + j : Ncls - 1;
}

__global__ void class_rows_fwd(const float* __restrict__ x, const double* __restrict__ weight,
                               const double* __restrict__ bias, const int32_t* __restrict__ obj,
                               double* __restrict__ out, int B, int Q, int C, int Pn, int64_t Ncls) {
    // one warp per (b, q, j)
    int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (wi >= (int64_t)B * Q * (Pn + 1)) return;
    int j = (int)(wi % (Pn + 1));
    int64_t bq = wi / (Pn + 1);
    int b = (int)(bq / Q);
    int64_t row = class_row(obj, b, j, Pn, Ncls);
    const float* xr = x + bq * C;
    const double* wr = weight + row * C;
    double acc = 0.0;
    for (int c = lane; c < C; c += 32) acc += (double)xr[c] * wr[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[wi] = acc + bias[row];
}

__global__ void class_rows_bwd_x(const double* __restrict__ weight, const int32_t* __restrict__ obj,
                                 const double* __restrict__ gout, float* __restrict__ gx, int B, int Q, int C, int Pn,
                                 int64_t Ncls) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * Q * C) return;
    int c = (int)(idx % C);
    int64_t bq = idx / C;
    int b = (int)(bq / Q);
    double acc = 0.0;
    for (int j = 0; j <= Pn; ++j) acc += gout[bq * (Pn + 1) + j] * weight[class_row(obj, b, j, Pn, Ncls) * C + c];
    gx[idx] = (float)acc;
}

__global__ void class_rows_bwd_w(const float* __restrict__ x, const int32_t* __restrict__ obj,
                                 const double* __restrict__ gout, double* __restrict__ gw, double* __restrict__ gb,
                                 int B, int Q, int C, int Pn, int64_t Ncls) {
    // thread per (b, j, c); c == 0 also accumulates the bias
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * (Pn + 1) * C) return;
    int c = (int)(idx % C);
    int64_t bj = idx / C;
    int j = (int)(bj % (Pn + 1));
    int b = (int)(bj / (Pn + 1));
    int64_t row = class_row(obj, b, j, Pn, Ncls);
    double acc = 0.0, accb = 0.0;
    for (int q = 0; q < Q; ++q) {
        double g = gout[((int64_t)b * Q + q) * (Pn + 1) + j];
        acc += g * (double)x[((int64_t)b * Q + q) * C + c];
        accb += g;
    }
    atomicAdd(gw + row * C + c, acc);
    if (c == 0) atomicAdd(gb + row, accb);
}

static int fill_offsets(const int32_t* tgt_offset, int B, Offsets& off, const char* who) {
    PDB_REQUIRE(tgt_offset, "%s: null tgt_offset", who);
    PDB_REQUIRE(B > 0 && B <= kMaxBatch, "%s: B=%d outside [1,%d]", who, B, kMaxBatch);
    for (int b = 0; b <= B; ++b) {
        off.v[b] = tgt_offset[b];
        PDB_REQUIRE(b == 0 ? off.v[0] == 0 : off.v[b] >= off.v[b - 1], "%s: tgt_offset not a prefix sum", who);
    }
    return PDB_OK;
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_point_sample_forward(const void* src, int src_dtype, const int32_t* map_index, const float* coords,
                                        const int32_t* coord_index, float* out, int R, int P, int H, int W,
                                        void* stream) {
    PDB_REQUIRE(src && coords && out, "point_sample_forward: null pointer");
    PDB_REQUIRE(R >= 0 && P > 0 && H > 0 && W > 0, "point_sample_forward: bad shape");
    PDB_REQUIRE(src_dtype >= 0 && src_dtype <= 2, "point_sample_forward: src_dtype %d (0 f32, 1 uint8, 2 bit-packed)", src_dtype);
    if (R == 0) return PDB_OK;
    int64_t total = (int64_t)R * P;
    point_sample_fwd<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(src, src_dtype, map_index, coords,
                                                                                      coord_index, out, R, P, H, W);
    return launched("point_sample_forward");
}

extern "C" int pdb_point_sample_backward(const float* grad_out, const int32_t* map_index, const float* coords,
                                         const int32_t* coord_index, float* grad_src, int R, int P, int H, int W,
                                         void* stream) {
    PDB_REQUIRE(grad_out && coords && grad_src, "point_sample_backward: null pointer");
    PDB_REQUIRE(R >= 0 && P > 0 && H > 0 && W > 0, "point_sample_backward: bad shape");
    if (R == 0) return PDB_OK;
    int64_t total = (int64_t)R * P;
    point_sample_bwd<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(grad_out, map_index, coords,
                                                                                      coord_index, grad_src, R, P, H, W);
    return launched("point_sample_backward");
}

extern "C" int pdb_matcher_cost(const float* pred_pts, const float* tgt_pts, const float* cls_prob,
                                const int32_t* tgt_label, const int32_t* tgt_offset, float* cost, int B, int Q, int Kc,
                                int P, float w_class, float w_mask, float w_dice, void* stream) {
    PDB_REQUIRE(pred_pts && tgt_pts && cls_prob && tgt_label && cost, "matcher_cost: null pointer");
    PDB_REQUIRE(Q > 0 && Q <= 65535 && Kc > 0 && P > 0, "matcher_cost: bad shape");
    Offsets off;
    PDB_TRY(fill_offsets(tgt_offset, B, off, "matcher_cost"));
    if (off.v[B] == 0) return PDB_OK;
    dim3 grid((unsigned)Q, (unsigned)B);
    matcher_cost_kernel<<<grid, 512, 0, as_stream(stream)>>>(pred_pts, tgt_pts, cls_prob, tgt_label, off, cost, Q, Kc,
                                                             P, w_class, w_mask, w_dice);
    return launched("matcher_cost");
}

extern "C" int pdb_lsap_batched(const float* cost, const int32_t* tgt_offset, int64_t* pred_idx, int64_t* tgt_idx,
                                int B, int Q, void* stream) {
    PDB_REQUIRE(cost && pred_idx && tgt_idx, "lsap_batched: null pointer");
    PDB_REQUIRE(Q > 0, "lsap_batched: Q=%d", Q);
    Offsets off;
    PDB_TRY(fill_offsets(tgt_offset, B, off, "lsap_batched"));
    for (int b = 0; b < B; ++b) {
        int K = off.v[b + 1] - off.v[b];
        int nr = K < Q ? K : Q, nc = K < Q ? Q : K;
        PDB_REQUIRE(nr <= LS_MAXR && nc <= LS_MAXC, "lsap_batched: image %d is %d x %d (limits: min<=%d, max<=%d)", b,
                    Q, K, LS_MAXR, LS_MAXC);
    }
    if (off.v[B] == 0) return PDB_OK;
    lsap_kernel<<<(unsigned)B, 32, 0, as_stream(stream)>>>(cost, off, pred_idx, tgt_idx, Q);
    return launched("lsap_batched");
}

// splits > 1: the P points of every pair are cut into `splits` chunks handled by different CTAs (few pairs: B = 2); `partial`
// (Nm * splits * 4 floats) receives their sums and a second launch adds them in order.  splits <= 1: one CTA per pair.
extern "C" int pdb_point_loss_forward(const float* pred, const int64_t* pred_index, const void* gt,
                                      const int64_t* gt_index, const float* coords, float* sums, float* partial, int splits,
                                      int Nm, int P, int H, int W, int Hg, int Wg, int gt_bits, void* stream) {
    PDB_REQUIRE(pred && pred_index && gt && gt_index && coords && sums, "point_loss_forward: null pointer");
    PDB_REQUIRE(Nm >= 0 && P > 0 && H > 0 && W > 0 && Hg > 0 && Wg > 0, "point_loss_forward: bad shape");
    PDB_REQUIRE(splits <= 1 || partial, "point_loss_forward: splits > 1 needs the partial-sum buffer");
    if (Nm == 0) return PDB_OK;
    if (splits <= 1) {
        point_loss_fwd<<<(unsigned)Nm, 1024, 0, as_stream(stream)>>>(pred, pred_index, gt, gt_index, coords, sums, P, H, W,
                                                                    Hg, Wg, gt_bits, P);
        return launched("point_loss_forward");
    }
    const int chunk = (P + splits - 1) / splits;
    point_loss_fwd<<<dim3((unsigned)Nm, (unsigned)splits), 256, 0, as_stream(stream)>>>(pred, pred_index, gt, gt_index, coords,
                                                                                       partial, P, H, W, Hg, Wg, gt_bits, chunk);
    PDB_TRY(launched("point_loss_forward(partial)"));
    point_loss_fwd_final<<<(unsigned)((Nm * 4 + 127) / 128), 128, 0, as_stream(stream)>>>(partial, sums, Nm * 4, splits);
    return launched("point_loss_forward(final)");
}

extern "C" int pdb_point_loss_backward(const float* pred, const int64_t* pred_index, const void* gt,
                                       const int64_t* gt_index, const float* coords, const float* sums,
                                       const float* g_bce, const float* g_dice, float* grad_pred, int splits, int Nm, int P,
                                       int H, int W, int Hg, int Wg, int gt_bits, void* stream) {
    PDB_REQUIRE(pred && pred_index && gt && gt_index && coords && sums && g_bce && g_dice && grad_pred,
                "point_loss_backward: null pointer");
    PDB_REQUIRE(Nm >= 0 && P > 0 && H > 0 && W > 0 && Hg > 0 && Wg > 0, "point_loss_backward: bad shape");
    if (Nm == 0) return PDB_OK;
    if (splits <= 1) {
        point_loss_bwd<<<(unsigned)Nm, 1024, 0, as_stream(stream)>>>(pred, pred_index, gt, gt_index, coords, sums, g_bce,
                                                                    g_dice, grad_pred, P, H, W, Hg, Wg, gt_bits, P);
    } else {
        point_loss_bwd<<<dim3((unsigned)Nm, (unsigned)splits), 256, 0, as_stream(stream)>>>(
            pred, pred_index, gt, gt_index, coords, sums, g_bce, g_dice, grad_pred, P, H, W, Hg, Wg, gt_bits,
            (P + splits - 1) / splits);
    }
    return launched("point_loss_backward");
}

extern "C" int pdb_class_rows_forward(const float* x, const double* weight, const double* bias, const int32_t* obj,
                                      double* out, int B, int Q, int C, int Pn, int64_t Ncls, void* stream) {
    PDB_REQUIRE(x && weight && bias && obj && out, "class_rows_forward: null pointer");
    PDB_REQUIRE(B > 0 && Q > 0 && C > 0 && Pn > 0 && Ncls > Pn, "class_rows_forward: bad shape");
    int64_t warps = (int64_t)B * Q * (Pn + 1);
    class_rows_fwd<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(x, weight, bias, obj, out, B, Q,
                                                                                         C, Pn, Ncls);
    return launched("class_rows_forward");
}

extern "C" int pdb_class_rows_backward(const float* x, const double* weight, const int32_t* obj, const double* grad_out,
                                       float* grad_x, double* grad_weight, double* grad_bias, int B, int Q, int C,
                                       int Pn, int64_t Ncls, void* stream) {
    PDB_REQUIRE(x && weight && obj && grad_out && grad_x && grad_weight && grad_bias, "class_rows_backward: null pointer");
    PDB_REQUIRE(B > 0 && Q > 0 && C > 0 && Pn > 0 && Ncls > Pn, "class_rows_backward: bad shape");
    cudaStream_t st = as_stream(stream);
    int64_t n1 = (int64_t)B * Q * C;
    class_rows_bwd_x<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(weight, obj, grad_out, grad_x, B, Q, C, Pn, Ncls);
    PDB_TRY(launched("class_rows_backward(x)"));
    int64_t n2 = (int64_t)B * (Pn + 1) * C;
    class_rows_bwd_w<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(x, obj, grad_out, grad_weight, grad_bias, B, Q, C,
                                                                    Pn, Ncls);
    return launched("class_rows_backward(w)");
}
