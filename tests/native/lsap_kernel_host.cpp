// TEST INFRASTRUCTURE ONLY.  The batched LSAP kernel of partdistillation_b200/csrc/loss.cu compiled for the host through
// cuda_on_cpu.h.  The kernel text is not duplicated: tests/test_lsap_host_cpu.py cuts the section between the
// "batched rectangular LSAP" and "fused point-sampled BCE + dice" banners out of loss.cu into lsap_section.inc (in a
// temporary directory on the include path) right before compiling this file.
#include "cuda_on_cpu.h"

namespace pdb {
constexpr int kMaxBatch = 255;                  // as at the top of loss.cu
struct Offsets { int v[kMaxBatch + 1]; };
#include "lsap_section.inc"
}  // namespace pdb

// cost: image b's (Q, K_b) row-major block at Q * offsets[b]; pred_idx / tgt_idx: (offsets[B]) int64, as pdb_lsap_batched
extern "C" int host_lsap_batched(const float* cost, const int32_t* offsets, int64_t* pred_idx, int64_t* tgt_idx, int B,
                                 int Q) {
    pdb::Offsets off;
    for (int b = 0; b <= B; ++b) off.v[b] = offsets[b];
    cpu_cuda::launch(dim3((unsigned)B), dim3(32), [&] { pdb::lsap_kernel(cost, off, pred_idx, tgt_idx, Q); });
    return 0;
}
