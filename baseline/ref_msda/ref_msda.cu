// Same-box bar: the REFERENCE's own MSDeformAttn CUDA kernels, compiled UNMODIFIED for
// sm_100a from where they lie in the reference checkout (ops/src/cuda/ms_deform_im2col_cuda.cuh — included by path at
// build time, never copied into this repository), behind a C ABI so that they can be timed next to libpdb200's kernels.
// This file is only the launcher the reference's ms_deform_attn_cuda.cu:26-159 would be without ATen: one call per
// im2col_step chunk of the batch; the backward's gradient buffers are zero-filled by the caller (the reference
// allocates them with at::zeros_like, .cu:127-129).  Not product code, not linked into libpdb200.so.
#include <cstdint>
#include <cuda_runtime.h>

#include "cuda/ms_deform_im2col_cuda.cuh"

extern "C" int ref_msda_forward_f32(const float* value, const int64_t* spatial_shapes_dev, const int64_t* level_start_dev,
                                    const float* loc, const float* attn, float* out, int N, int S, int M, int D, int L, int Lq,
                                    int P, int im2col_step, void* stream) {
    const int step = N < im2col_step ? N : im2col_step;
    if (N % step) return 1;
    for (int n = 0; n < N / step; ++n)
        ms_deformable_im2col_cuda<float>((cudaStream_t)stream, value + (size_t)n * step * S * M * D, spatial_shapes_dev, level_start_dev,
                                         loc + (size_t)n * step * Lq * M * L * P * 2, attn + (size_t)n * step * Lq * M * L * P, step, S,
                                         M, D, L, Lq, P, out + (size_t)n * step * Lq * M * D);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" int ref_msda_backward_f32(const float* value, const int64_t* spatial_shapes_dev, const int64_t* level_start_dev,
                                     const float* loc, const float* attn, const float* grad_out, float* grad_value, float* grad_loc,
                                     float* grad_attn, int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                                     void* stream) {
    const int step = N < im2col_step ? N : im2col_step;
    if (N % step) return 1;
    for (int n = 0; n < N / step; ++n)
        ms_deformable_col2im_cuda<float>((cudaStream_t)stream, grad_out + (size_t)n * step * Lq * M * D,
                                         value + (size_t)n * step * S * M * D, spatial_shapes_dev, level_start_dev,
                                         loc + (size_t)n * step * Lq * M * L * P * 2, attn + (size_t)n * step * Lq * M * L * P, step, S,
                                         M, D, L, Lq, P, grad_value + (size_t)n * step * S * M * D,
                                         grad_loc + (size_t)n * step * Lq * M * L * P * 2, grad_attn + (size_t)n * step * Lq * M * L * P);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
