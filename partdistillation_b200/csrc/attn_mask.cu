// Attention-mask build: bilinear resize (align_corners=False) of the mask logits to the next level's
// size, sigmoid() < 0.5, stored once per (b, q) as uint8 (the reference repeats it over heads and
// batches it as (B*heads, Q, hw): mask2former_transformer_decoder.py:453-457), plus the per-row
// "has any attended key" flag that implements the all-masked-row reset of :405.
//
// Bit-exactness contract: the interpolation follows ATen's upsample_bilinear2d expression
//   h0*(w0*a + w1*b) + h1*(w0*c + w1*d),  src = scale*(dst+0.5)-0.5 clamped at 0,
// and the threshold is the literal fp32  1/(1+expf(-x)) < 0.5  (NOT x < 0: the predicate is false
// for -1.79e-7 < x < 0).  For the power-of-two ratios of every shipped config all lambdas are 0.5,
// the products are exact and the result is independent of FMA contraction.
#include "common.cuh"

namespace pdb {

__global__ void __launch_bounds__(256)
attn_mask_kernel(const float* __restrict__ logits, uint8_t* __restrict__ mask, int32_t* __restrict__ row_any,
                 int H, int W, int h, int w, float rh, float rw) {
    const int r = blockIdx.y;                       // row = b*Q + q
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int hw = h * w;
    bool attended = false;
    if (o < hw) {
        const int oy = o / w, ox = o - oy * w;
        float h1r = fmaxf(rh * ((float)oy + 0.5f) - 0.5f, 0.f);
        float w1r = fmaxf(rw * ((float)ox + 0.5f) - 0.5f, 0.f);
        int h1 = (int)h1r, w1 = (int)w1r;
        int h1p = (h1 < H - 1) ? 1 : 0, w1p = (w1 < W - 1) ? 1 : 0;
        float h1l = h1r - (float)h1, h0l = 1.f - h1l;
        float w1l = w1r - (float)w1, w0l = 1.f - w1l;
        const float* src = logits + (int64_t)r * H * W;
        float a = __ldg(src + (int64_t)h1 * W + w1);
        float b = __ldg(src + (int64_t)h1 * W + w1 + w1p);
        float c = __ldg(src + (int64_t)(h1 + h1p) * W + w1);
        float d = __ldg(src + (int64_t)(h1 + h1p) * W + w1 + w1p);
        float val = h0l * (w0l * a + w1l * b) + h1l * (w0l * c + w1l * d);
        float sig = 1.0f / (1.0f + expf(-val));
        bool masked = sig < 0.5f;
        mask[(int64_t)r * hw + o] = masked ? 1 : 0;
        attended = !masked;
    }
    unsigned any = __ballot_sync(0xffffffffu, attended);
    if ((threadIdx.x & 31) == 0 && any) atomicOr(row_any + r, 1);
}

__global__ void attn_mask_reset_kernel(uint8_t* __restrict__ mask, const int32_t* __restrict__ row_any, int64_t hw) {
    const int r = blockIdx.y;
    if (row_any[r] != 0) return;
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < hw) mask[(int64_t)r * hw + o] = 0;
}

}  // namespace pdb

using namespace pdb;

extern "C" int pdb_attn_mask_build(const float* logits, uint8_t* mask, int32_t* row_any, int B, int Q, int H, int W,
                                   int h, int w, void* stream) {
    PDB_REQUIRE(logits && mask && row_any, "attn_mask_build: null pointer");
    PDB_REQUIRE(B > 0 && Q > 0 && H > 0 && W > 0 && h > 0 && w > 0, "attn_mask_build: non-positive dimension");
    PDB_REQUIRE((int64_t)B * Q <= 65535, "attn_mask_build: B*Q=%lld exceeds grid.y", (long long)B * Q);
    // area_pixel_compute_scale(align_corners=False, no explicit scale): (float)in / out
    float rh = (float)H / (float)h, rw = (float)W / (float)w;
    dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)(B * Q));
    attn_mask_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, mask, row_any, H, W, h, w, rh, rw);
    return launched("attn_mask_build");
}

extern "C" int pdb_attn_mask_reset_rows(uint8_t* mask, const int32_t* row_any, int rows, int64_t hw, void* stream) {
    PDB_REQUIRE(mask && row_any, "attn_mask_reset_rows: null pointer");
    PDB_REQUIRE(rows > 0 && rows <= 65535 && hw > 0, "attn_mask_reset_rows: bad shape");
    dim3 grid((unsigned)((hw + 255) / 256), (unsigned)rows);
    attn_mask_reset_kernel<<<grid, 256, 0, as_stream(stream)>>>(mask, row_any, hw);
    return launched("attn_mask_reset_rows");
}
