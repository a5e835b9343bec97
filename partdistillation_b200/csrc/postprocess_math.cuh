// Per-pixel arithmetic of the inference post-processing kernels (postprocess.cu): bilinear source taps, the 2x2
// logit patch behind one sample of the padded-size map, and the composed two-pass value.  Kept in a header of
// __host__ __device__ functions so that tests/test_postprocess_host_cpu.py can compile the SAME code for the host
// (tests/native/postprocess_host.cpp) and compare it with the oracle where there is no GPU; the product only ever
// calls it from the kernels.
//
// Every bilinear pass evaluates ATen's upsample_bilinear2d expression
//   h0*(w0*a + w1*b) + h1*(w0*c + w1*d),   src = scale*(dst + 0.5) - 0.5 clamped at 0,  scale = (float)in/out,
// with every product and sum rounded separately (no FMA contraction, so the bits do not depend on the compiler).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PDB_HD __host__ __device__ __forceinline__
#else
#define PDB_HD inline
#endif

namespace pdb {

#if defined(__CUDA_ARCH__)
PDB_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
PDB_HD float add_rn(float a, float b) { return __fadd_rn(a, b); }
PDB_HD float load_f32(const float* p) { return __ldg(p); }
#else
PDB_HD float mul_rn(float a, float b) { volatile float r = a * b; return r; }
PDB_HD float add_rn(float a, float b) { volatile float r = a + b; return r; }
PDB_HD float load_f32(const float* p) { return *p; }
#endif

PDB_HD float pdb_fmaxf(float a, float b) { return a > b ? a : b; }

struct Tap1D {          // one bilinear source coordinate: indices i, i + p and weights (1 - l), l
    int i, p;
    float l0, l1;
};

PDB_HD Tap1D make_tap(float scale, int dst, int in_size) {
    float r = pdb_fmaxf(add_rn(mul_rn(scale, add_rn((float)dst, 0.5f)), -0.5f), 0.f);
    Tap1D t;
    t.i = (int)r;
    if (t.i > in_size - 1) t.i = in_size - 1;       // only reachable through rounding when up-sampling by < 1 ulp
    t.p = (t.i < in_size - 1) ? 1 : 0;
    t.l1 = add_rn(r, -(float)t.i);
    t.l0 = add_rn(1.f, -t.l1);
    return t;
}

PDB_HD float bilerp(float a, float b, float c, float d, float h0, float h1, float w0, float w1) {
    float top = add_rn(mul_rn(w0, a), mul_rn(w1, b));
    float bot = add_rn(mul_rn(w0, c), mul_rn(w1, d));
    return add_rn(mul_rn(h0, top), mul_rn(h1, bot));
}

// One stage-1 sample (value of the padded-size map at row Y, column X) needs a 2x2 patch of the logits.
struct Patch {
    int o00, o01, o10, o11;     // offsets into one query's (h, w) logit plane
    float h0, h1, w0, w1;
};

PDB_HD Patch make_patch(const Tap1D& ty, const Tap1D& tx, int w) {
    Patch p;
    p.o00 = ty.i * w + tx.i;
    p.o01 = p.o00 + tx.p;
    p.o10 = (ty.i + ty.p) * w + tx.i;
    p.o11 = p.o10 + tx.p;
    p.h0 = ty.l0; p.h1 = ty.l1; p.w0 = tx.l0; p.w1 = tx.l1;
    return p;
}

PDB_HD float sample_patch(const float* __restrict__ plane, const Patch& p) {
    return bilerp(load_f32(plane + p.o00), load_f32(plane + p.o01), load_f32(plane + p.o10), load_f32(plane + p.o11),
                  p.h0, p.h1, p.w0, p.w1);
}

// Offsets / weights of everything one output pixel needs: 1 patch (second pass = identity) or 4 patches + outer weights.
struct PixelTaps {
    Patch p00, p01, p10, p11;
    float H0, H1, W0, W1;
};

template <bool TWO_STAGE>
PDB_HD PixelTaps make_pixel_taps(int oy, int ox, int h, int w, int Hi, int Wi, float s1h, float s1w, float s2h, float s2w) {
    PixelTaps t;
    if (TWO_STAGE) {
        Tap1D ty = make_tap(s2h, oy, Hi), tx = make_tap(s2w, ox, Wi);
        t.H0 = ty.l0; t.H1 = ty.l1; t.W0 = tx.l0; t.W1 = tx.l1;
        Tap1D y0 = make_tap(s1h, ty.i, h), y1 = make_tap(s1h, ty.i + ty.p, h);
        Tap1D x0 = make_tap(s1w, tx.i, w), x1 = make_tap(s1w, tx.i + tx.p, w);
        t.p00 = make_patch(y0, x0, w); t.p01 = make_patch(y0, x1, w);
        t.p10 = make_patch(y1, x0, w); t.p11 = make_patch(y1, x1, w);
    } else {
        Tap1D y0 = make_tap(s1h, oy, h), x0 = make_tap(s1w, ox, w);
        t.p00 = make_patch(y0, x0, w);
        t.p01 = t.p10 = t.p11 = t.p00;
        t.H0 = 1.f; t.H1 = 0.f; t.W0 = 1.f; t.W1 = 0.f;
    }
    return t;
}

template <bool TWO_STAGE>
PDB_HD float sample_pixel(const float* plane, const PixelTaps& t) {
    if (TWO_STAGE)
        return bilerp(sample_patch(plane, t.p00), sample_patch(plane, t.p01), sample_patch(plane, t.p10),
                      sample_patch(plane, t.p11), t.H0, t.H1, t.W0, t.W1);
    return sample_patch(plane, t.p00);
}

// sem_seg_postprocess(mask.float()).bool() of one output pixel of a zero-padded 0/1 mask with row stride Wp
PDB_HD bool resized_mask_bit(const uint8_t* src, int Wp, int Hi, int Wi, int oy, int ox, float sh, float sw) {
    Tap1D ty = make_tap(sh, oy, Hi), tx = make_tap(sw, ox, Wi);
    float a = src[(int64_t)ty.i * Wp + tx.i] ? 1.f : 0.f;
    float b = src[(int64_t)ty.i * Wp + tx.i + tx.p] ? 1.f : 0.f;
    float c = src[(int64_t)(ty.i + ty.p) * Wp + tx.i] ? 1.f : 0.f;
    float d = src[(int64_t)(ty.i + ty.p) * Wp + tx.i + tx.p] ? 1.f : 0.f;
    return bilerp(a, b, c, d, ty.l0, ty.l1, tx.l0, tx.l1) != 0.f;
}

}  // namespace pdb
