// TEST INFRASTRUCTURE ONLY.  The loss-side kernels of partdistillation_b200/csrc/loss.cu (point sampling, matcher cost,
// batched LSAP, fused point BCE + dice, PartDistillation's classifier rows) compiled for the host through cuda_on_cpu.h.
// tests/test_loss_kernels_host_cpu.py cuts the whole `namespace pdb { ... }` block out of loss.cu into loss_section.inc;
// the entry points below restate the launch geometry of the pdb_* launchers at the end of loss.cu, with the C ABI's argument
// order minus the stream, so that partdistillation_b200/functional.py's wrappers can drive them unmodified.
#include "pdb_common_host.h"

#include "loss_section.inc"

using namespace pdb;
using cpu_cuda::launch;

extern "C" int host_point_sample_forward(const void* src, int src_dtype, const int32_t* map_index, const float* coords,
                                         const int32_t* coord_index, float* out, int R, int P, int H, int W) {
    if (R == 0) return 0;
    const int64_t total = (int64_t)R * P;
    launch(dim3((unsigned)((total + 255) / 256)), dim3(256),
           [&] { point_sample_fwd(src, src_dtype, map_index, coords, coord_index, out, R, P, H, W); });
    return 0;
}

extern "C" int host_point_sample_backward(const float* grad_out, const int32_t* map_index, const float* coords,
                                          const int32_t* coord_index, float* grad_src, int R, int P, int H, int W) {
    if (R == 0) return 0;
    const int64_t total = (int64_t)R * P;
    launch(dim3((unsigned)((total + 255) / 256)), dim3(256),
           [&] { point_sample_bwd(grad_out, map_index, coords, coord_index, grad_src, R, P, H, W); });
    return 0;
}

extern "C" int host_matcher_cost(const float* pred_pts, const float* tgt_pts, const float* cls_prob, const int32_t* tgt_label,
                                 const int32_t* tgt_offset, float* cost, int B, int Q, int Kc, int P, float w_class,
                                 float w_mask, float w_dice) {
    Offsets off;
    PDB_TRY(fill_offsets(tgt_offset, B, off, "matcher_cost"));
    if (off.v[B] == 0) return 0;
    launch(dim3((unsigned)Q, (unsigned)B), dim3(512),
           [&] { matcher_cost_kernel(pred_pts, tgt_pts, cls_prob, tgt_label, off, cost, Q, Kc, P, w_class, w_mask, w_dice); });
    return 0;
}

extern "C" int host_lsap_batched(const float* cost, const int32_t* tgt_offset, int64_t* pred_idx, int64_t* tgt_idx, int B,
                                 int Q) {
    Offsets off;
    PDB_TRY(fill_offsets(tgt_offset, B, off, "lsap_batched"));
    if (off.v[B] == 0) return 0;
    launch(dim3((unsigned)B), dim3(32), [&] { lsap_kernel(cost, off, pred_idx, tgt_idx, Q); });
    return 0;
}

extern "C" int host_point_loss_forward(const float* pred, const int64_t* pred_index, const void* gt, const int64_t* gt_index,
                                       const float* coords, float* sums, float* partial, int splits, int Nm, int P, int H, int W,
                                       int Hg, int Wg, int gt_bits) {
    if (Nm == 0) return 0;
    if (splits <= 1) {
        launch(dim3((unsigned)Nm), dim3(1024),
               [&] { point_loss_fwd(pred, pred_index, gt, gt_index, coords, sums, P, H, W, Hg, Wg, gt_bits, P); });
        return 0;
    }
    const int chunk = (P + splits - 1) / splits;
    launch(dim3((unsigned)Nm, (unsigned)splits), dim3(256),
           [&] { point_loss_fwd(pred, pred_index, gt, gt_index, coords, partial, P, H, W, Hg, Wg, gt_bits, chunk); });
    launch(dim3((unsigned)((Nm * 4 + 127) / 128)), dim3(128), [&] { point_loss_fwd_final(partial, sums, Nm * 4, splits); });
    return 0;
}

extern "C" int host_point_loss_backward(const float* pred, const int64_t* pred_index, const void* gt,
                                        const int64_t* gt_index, const float* coords, const float* sums, const float* g_bce,
                                        const float* g_dice, float* grad_pred, int splits, int Nm, int P, int H, int W, int Hg,
                                        int Wg, int gt_bits) {
    if (Nm == 0) return 0;
    if (splits <= 1)
        launch(dim3((unsigned)Nm), dim3(1024), [&] {
            point_loss_bwd(pred, pred_index, gt, gt_index, coords, sums, g_bce, g_dice, grad_pred, P, H, W, Hg, Wg, gt_bits, P);
        });
    else
        launch(dim3((unsigned)Nm, (unsigned)splits), dim3(256), [&] {
            point_loss_bwd(pred, pred_index, gt, gt_index, coords, sums, g_bce, g_dice, grad_pred, P, H, W, Hg, Wg, gt_bits,
                           (P + splits - 1) / splits);
        });
    return 0;
}

extern "C" int host_class_rows_forward(const float* x, const double* weight, const double* bias, const int32_t* obj,
                                       double* out, int B, int Q, int C, int Pn, int64_t Ncls) {
    const int64_t warps = (int64_t)B * Q * (Pn + 1);
    launch(dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256),
           [&] { class_rows_fwd(x, weight, bias, obj, out, B, Q, C, Pn, Ncls); });
    return 0;
}

extern "C" int host_class_rows_backward(const float* x, const double* weight, const int32_t* obj, const double* grad_out,
                                        float* grad_x, double* grad_weight, double* grad_bias, int B, int Q, int C, int Pn,
                                        int64_t Ncls) {
    const int64_t n1 = (int64_t)B * Q * C;
    launch(dim3((unsigned)((n1 + 255) / 256)), dim3(256),
           [&] { class_rows_bwd_x(weight, obj, grad_out, grad_x, B, Q, C, Pn, Ncls); });
    const int64_t n2 = (int64_t)B * (Pn + 1) * C;
    launch(dim3((unsigned)((n2 + 255) / 256)), dim3(256),
           [&] { class_rows_bwd_w(x, obj, grad_out, grad_weight, grad_bias, B, Q, C, Pn, Ncls); });
    return 0;
}
