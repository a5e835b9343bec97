// TEST INFRASTRUCTURE ONLY.  The attention-mask kernels of partdistillation_b200/csrc/attn_mask.cu compiled for the host
// through cuda_on_cpu.h.  tests/test_attn_mask_host_cpu.py cuts the `namespace pdb { ... }` block out of attn_mask.cu into
// attn_mask_section.inc (temporary directory on the include path) right before compiling this file; the launch geometry
// below restates pdb_attn_mask_build / pdb_attn_mask_reset_rows.
#include "cuda_on_cpu.h"

#include "attn_mask_section.inc"

extern "C" int host_attn_mask_build(const float* logits, uint8_t* mask, int32_t* row_any, int B, int Q, int H, int W, int h,
                                    int w) {
    const float rh = (float)H / (float)h, rw = (float)W / (float)w;
    cpu_cuda::launch(dim3((unsigned)((h * w + 255) / 256), (unsigned)(B * Q)), dim3(256),
                     [&] { pdb::attn_mask_kernel(logits, mask, row_any, H, W, h, w, rh, rw); });
    return 0;
}

extern "C" int host_attn_mask_reset_rows(uint8_t* mask, const int32_t* row_any, int rows, int64_t hw) {
    cpu_cuda::launch(dim3((unsigned)((hw + 255) / 256), (unsigned)rows), dim3(256),
                     [&] { pdb::attn_mask_reset_kernel(mask, row_any, hw); });
    return 0;
}
