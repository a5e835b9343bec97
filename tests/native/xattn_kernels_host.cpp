// TEST INFRASTRUCTURE ONLY.  The masked cross-attention kernels of partdistillation_b200/csrc/xattn.cu compiled for the host
// through cuda_on_cpu.h.  tests/test_xattn_kernels_host_cpu.py cuts the `namespace pdb { ... }` block out of xattn.cu into
// xattn_section.inc; the entry points restate the launchers at the end of xattn.cu (C ABI argument order minus the stream).
#include "pdb_common_host.h"
#include "xattn_section.inc"

using namespace pdb;
using cpu_cuda::launch;

extern "C" int64_t host_masked_xattn_workspace_bytes(int B, int heads, int Q, int Lk, int d) {
    if (B <= 0 || heads <= 0 || Q <= 0 || Lk <= 0 || d != XD) return -1;
    int ns = xattn_nsplit(B, heads, Q, Lk);
    return (int64_t)B * heads * ns * Q * (XD + 2) * (int64_t)sizeof(float);
}

extern "C" int host_masked_xattn_forward(const float* q, const float* k, const float* v, const uint8_t* mask,
                                         const int32_t* row_any, float* out, float* lse, void* workspace, int B, int heads,
                                         int Q, int Lk, int d) {
    if (d != XD) return -1;
    int ns = xattn_nsplit(B, heads, Q, Lk);
    int tiles = (Lk + XTK - 1) / XTK;
    int tiles_per = (tiles + ns - 1) / ns;
    int qtiles = (Q + XTHREADS - 1) / XTHREADS;
    float* ws_acc = (float*)workspace;
    float* ws_ml = ws_acc + (int64_t)B * heads * ns * Q * XD;
    launch(dim3((unsigned)(ns * qtiles), (unsigned)heads, (unsigned)B), dim3(XTHREADS),
           [&] { xattn_fwd_partial(q, k, v, mask, row_any, ws_acc, ws_ml, heads, Q, Lk, ns, tiles_per, qtiles); });
    int64_t warps = (int64_t)B * heads * Q;
    launch(dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256),
           [&] { xattn_fwd_combine(ws_acc, ws_ml, out, lse, B, heads, Q, ns); });
    return 0;
}

extern "C" int host_masked_xattn_backward(const float* q, const float* k, const float* v, const uint8_t* mask,
                                          const int32_t* row_any, const float* out, const float* lse, const float* grad_out,
                                          float* grad_q, float* grad_k, float* grad_v, int B, int heads, int Q, int Lk, int d) {
    if (d != XD) return -1;
    int ns = xattn_nsplit(B, heads, Q, Lk);
    int tiles = (Lk + XTK - 1) / XTK;
    int tiles_per = (tiles + ns - 1) / ns;
    int qtiles = (Q + XTHREADS - 1) / XTHREADS;
    std::fill(grad_q, grad_q + (size_t)B * Q * heads * XD, 0.f);                     // cudaMemsetAsync in the launcher
    launch(dim3((unsigned)(ns * qtiles), (unsigned)heads, (unsigned)B), dim3(XTHREADS),
           [&] { xattn_bwd_dq(q, k, v, mask, row_any, out, lse, grad_out, grad_q, heads, Q, Lk, ns, tiles_per); });
    launch(dim3((unsigned)((Lk + XTHREADS - 1) / XTHREADS), (unsigned)heads, (unsigned)B), dim3(XTHREADS),
           [&] { xattn_bwd_dkv(q, k, v, mask, row_any, out, lse, grad_out, grad_k, grad_v, heads, Q, Lk); });
    return 0;
}
