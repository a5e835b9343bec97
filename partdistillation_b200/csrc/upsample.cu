// FPN top-down step of the pixel decoder (msdeformattn.py:352-356): y = lateral + F.interpolate(x, size=lateral.shape[-2:],
// mode="bilinear", align_corners=False) on channels-last maps, forward in one pass and the input gradient as a gather.
// ATen's upsample_bilinear2d_nhwc kernels take 277 us (forward) and 253 us (backward, atomics) for (2, 256, 128, 128) ->
// 256 x 256 — 0.5 TB/s — and the add is a third pass over the 134 MB map.
//   forward : out[b][oy][ox][c] = lat[b][oy][ox][c] + sum of the 4 taps of x            (x read through L2: 4 outputs per input)
//   backward: gx[b][iy][ix][c]  = sum over the output pixels whose taps touch (iy, ix) of weight * g   (no atomics, deterministic)
// Source index and weights follow ATen's area_pixel_compute_source_index (align_corners = false, scale = in / out, negative
// source clamped to 0, upper neighbour clamped to the last row / column).
#include "common.cuh"

namespace pdb {

struct Tap {
    int i0, i1;
    float w0, w1;
};

__device__ __forceinline__ Tap up_tap(int o, float scale, int in) {
    float s = scale * ((float)o + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    Tap t;
    t.i0 = (int)s;
    t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0);
    t.w1 = s - (float)t.i0;
    t.w0 = 1.f - t.w1;
    return t;
}

// thread = (output pixel, float4 of channels); x: (B, h, w, C4) with batch stride xs (float4 units)
__global__ void __launch_bounds__(256)
upsample_add_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ lat, float4* __restrict__ out, int64_t total,
                        int h, int w, int H, int W, int C4, int64_t xs, float sy, float sx) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C4);
    int64_t p = idx / C4;
    const int ox = (int)(p % W);
    p /= W;
    const int oy = (int)(p % H), b = (int)(p / H);
    const Tap ty = up_tap(oy, sy, h), tx = up_tap(ox, sx, w);
    const float4* xb = x + (int64_t)b * xs + c;
    const float4 a = __ldg(xb + ((int64_t)ty.i0 * w + tx.i0) * C4), bq = __ldg(xb + ((int64_t)ty.i0 * w + tx.i1) * C4);
    const float4 cq = __ldg(xb + ((int64_t)ty.i1 * w + tx.i0) * C4), d = __ldg(xb + ((int64_t)ty.i1 * w + tx.i1) * C4);
    float4 r;
    r.x = ty.w0 * (tx.w0 * a.x + tx.w1 * bq.x) + ty.w1 * (tx.w0 * cq.x + tx.w1 * d.x);
    r.y = ty.w0 * (tx.w0 * a.y + tx.w1 * bq.y) + ty.w1 * (tx.w0 * cq.y + tx.w1 * d.y);
    r.z = ty.w0 * (tx.w0 * a.z + tx.w1 * bq.z) + ty.w1 * (tx.w0 * cq.z + tx.w1 * d.z);
    r.w = ty.w0 * (tx.w0 * a.w + tx.w1 * bq.w) + ty.w1 * (tx.w0 * cq.w + tx.w1 * d.w);
    if (lat) {
        const float4 l = __ldg(lat + idx);
        r.x += l.x; r.y += l.y; r.z += l.z; r.w += l.w;
    }
    out[idx] = r;
}

// weight with which output index o reads input index i (0 if it does not)
__device__ __forceinline__ float up_weight(int o, int i, float scale, int in) {
    const Tap t = up_tap(o, scale, in);
    return (t.i0 == i ? t.w0 : 0.f) + (t.i1 == i ? t.w1 : 0.f);
}

// thread = (input pixel, float4 of channels): gathers the output pixels in [lo, hi] x [lo, hi] that can touch it
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float4* __restrict__ g, float4* __restrict__ gx, int64_t total, int h, int w, int H, int W, int C4,
                    float sy, float sx, float ry, float rx) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C4);
    int64_t p = idx / C4;
    const int ix = (int)(p % w);
    p /= w;
    const int iy = (int)(p % h), b = (int)(p / h);
    // output o reads inputs floor(s), floor(s) + 1 with s = scale * (o + 0.5) - 0.5: candidates are s in (i - 1, i + 1), i.e.
    // o in ((i - 0.5) / scale - 0.5, (i + 1.5) / scale - 0.5), widened by one on each side against rounding; the border rows also
    // collect the clamped sources
    const int oy0 = max(0, (int)floorf(((float)iy - 0.5f) * ry - 0.5f) - 1), oy1 = min(H - 1, (int)ceilf(((float)iy + 1.5f) * ry - 0.5f) + 1);
    const int ox0 = max(0, (int)floorf(((float)ix - 0.5f) * rx - 0.5f) - 1), ox1 = min(W - 1, (int)ceilf(((float)ix + 1.5f) * rx - 0.5f) + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* gb = g + (int64_t)b * H * W * C4 + c;
    for (int oy = oy0; oy <= oy1; ++oy) {
        const float wy = up_weight(oy, iy, sy, h);
        if (wy == 0.f) continue;
        for (int ox = ox0; ox <= ox1; ++ox) {
            const float wgt = wy * up_weight(ox, ix, sx, w);
            if (wgt == 0.f) continue;
            const float4 v = __ldg(gb + ((int64_t)oy * W + ox) * C4);
            acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
            acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
        }
    }
    gx[idx] = acc;
}

}  // namespace pdb

using namespace pdb;

// x: (B, h, w, C) f32 pixel-major with batch stride x_batch_stride (elements; h * w * C when contiguous); lateral: (B, H, W, C)
// contiguous or NULL; out: (B, H, W, C) contiguous.  C % 4 == 0, 16-byte aligned bases, x_batch_stride % 4 == 0.
extern "C" int pdb_upsample_add_forward(const float* x, const float* lateral, float* out, int B, int h, int w, int H, int W, int C,
                                        int64_t x_batch_stride, void* stream) {
    PDB_REQUIRE(x && out, "upsample_add_forward: null pointer");
    PDB_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && x_batch_stride % 4 == 0,
                "upsample_add_forward: bad shape (C and the batch stride must be multiples of 4)");
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lateral) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "upsample_add_forward: buffers must be 16-byte aligned");
    const int C4 = C / 4;
    const int64_t total = (int64_t)B * H * W * C4;
    PDB_REQUIRE((total + 255) / 256 < (1ll << 31), "upsample_add_forward: too large");
    upsample_add_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(lateral), reinterpret_cast<float4*>(out), total, h, w, H,
        W, C4, x_batch_stride / 4, (float)h / (float)H, (float)w / (float)W);
    return launched("upsample_add_forward");
}

// grad_out: (B, H, W, C) contiguous; grad_x: (B, h, w, C) contiguous, overwritten.
extern "C" int pdb_upsample_backward(const float* grad_out, float* grad_x, int B, int h, int w, int H, int W, int C, void* stream) {
    PDB_REQUIRE(grad_out && grad_x, "upsample_backward: null pointer");
    PDB_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "upsample_backward: bad shape");
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0,
                "upsample_backward: buffers must be 16-byte aligned");
    const int C4 = C / 4;
    const int64_t total = (int64_t)B * h * w * C4;
    PDB_REQUIRE((total + 255) / 256 < (1ll << 31), "upsample_backward: too large");
    upsample_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(grad_out), reinterpret_cast<float4*>(grad_x), total, h, w, H, W, C4, (float)h / (float)H,
        (float)w / (float)W, (float)H / (float)h, (float)W / (float)w);
    return launched("upsample_backward");
}

// Zero-padded copy of a channels-last map for the tap-shifted 3x3 convolution GEMM (functional._pad_nhwc): (B, H, W, C) ->
// (B, H + top + bottom, W + left + right, C), borders written as zeros in the same pass (ATen: a fill pass plus a strided copy
// at 1.4 TB/s).
namespace pdb {
__global__ void __launch_bounds__(256)
pad_nhwc_kernel(const float4* __restrict__ x, float4* __restrict__ out, int64_t total, int H, int W, int C4, int Hp, int Wp, int top,
                int left) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C4);
    int64_t p = idx / C4;
    const int xx = (int)(p % Wp) - left;
    p /= Wp;
    const int yy = (int)(p % Hp) - top, b = (int)(p / Hp);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xx >= 0 && xx < W && yy >= 0 && yy < H) v = __ldg(x + (((int64_t)b * H + yy) * W + xx) * C4 + c);
    out[idx] = v;
}
}  // namespace pdb

extern "C" int pdb_pad_nhwc(const float* x, float* out, int B, int H, int W, int C, int top, int bottom, int left, int right,
                            void* stream) {
    PDB_REQUIRE(x && out, "pad_nhwc: null pointer");
    PDB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && top >= 0 && bottom >= 0 && left >= 0 && right >= 0, "pad_nhwc: bad shape");
    PDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "pad_nhwc: buffers must be 16-byte aligned");
    const int Hp = H + top + bottom, Wp = W + left + right, C4 = C / 4;
    const int64_t total = (int64_t)B * Hp * Wp * C4;
    PDB_REQUIRE((total + 255) / 256 < (1ll << 31), "pad_nhwc: too large");
    pdb::pad_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), total, H, W, C4, Hp, Wp, top, left);
    return launched("pad_nhwc");
}
