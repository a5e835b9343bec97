// TEST INFRASTRUCTURE ONLY.  The MSDeformAttn kernels of partdistillation_b200/csrc/msda.cu (generic f32 / f64, the D = 32
// fast path and the tiled encoder path, forward and backward) compiled for the host through cuda_on_cpu.h.
// tests/test_msda_host_cpu.py cuts the kernel part of msda.cu (everything inside `namespace pdb` ahead of the launch
// helpers) into msda_section.inc, rewriting only the `extern __shared__` declarations to the shim's dynamic shared
// memory pointer.  The dispatch below restates pdb_msda_forward / pdb_msda_backward (msda.cu) line by line.
#include "pdb_common_host.h"

namespace pdb {
#include "msda_section.inc"
}  // namespace pdb

using namespace pdb;
using cpu_cuda::launch;

extern "C" int host_msda_forward(const void* value, const int64_t* shapes_hw, const int64_t* level_start, const void* loc,
                                 const void* attn, void* out, int N, int S, int M, int D, int Lq, int L, int P, int dtype) {
    LevelTable lt;
    if (make_levels(shapes_hw, level_start, L, S, lt) != PDB_OK) return -1;
    const int64_t total = (int64_t)N * Lq * M * D;
    const unsigned gblocks = (unsigned)((total + 255) / 256);
    if (dtype == 1) {
        launch(dim3(gblocks), dim3(256), [&] {
            msda_fwd_generic<double>((const double*)value, lt, (const double*)loc, (const double*)attn, (double*)out, total, S,
                                     M, D, Lq, L, P);
        });
        return 1;
    }
    if (D == 32 && P == 4 && levels_at_least_2x2(lt, L)) {
        const int64_t slots = (int64_t)N * Lq * M;
        PatchTable pt;
        const bool tiled = Lq == S;
        int64_t blocks = make_patches(lt, L, N, M, pt);
        if (!tiled) blocks = (slots + kTileSlots - 1) / kTileSlots;
        const size_t smem = (sizeof(float4) + sizeof(int)) * kTileSlots * (L * P + 1);
        launch(dim3((unsigned)blocks), dim3(kTileSlots * 8), smem, [&] {
            msda_fwd_tiled<4>((const float*)value, lt, pt, (const float*)loc, (const float*)attn, (float*)out, slots, S, M, Lq,
                              L, tiled ? 1 : 0);
        });
        return tiled ? 3 : 4;
    }
    if (D == 32 && P == 4) {
        const int64_t slots = (int64_t)N * Lq * M;
        const int64_t blocks = (slots + kSlotsPerCta - 1) / kSlotsPerCta;
        const size_t smem = sizeof(float) * kSlotsPerCta * ((L * P * 2 + 2) + (L * P + 1));
        launch(dim3((unsigned)blocks), dim3(kFastThreads), smem, [&] {
            msda_fwd_d32<4>((const float*)value, lt, (const float*)loc, (const float*)attn, (float*)out, slots, S, M, Lq, L);
        });
        return 2;
    }
    launch(dim3(gblocks), dim3(256), [&] {
        msda_fwd_generic<float>((const float*)value, lt, (const float*)loc, (const float*)attn, (float*)out, total, S, M, D, Lq,
                                L, P);
    });
    return 1;
}

// grad_value must be ZERO-FILLED by the caller here (pdb_msda_backward does it with cudaMemsetAsync); grad_loc / grad_attn
// are zero-filled below where the generic kernel accumulates into them, as bwd_generic does.
extern "C" int host_msda_backward(const void* value, const int64_t* shapes_hw, const int64_t* level_start, const void* loc,
                                  const void* attn, const void* grad_out, void* grad_value, void* grad_loc, void* grad_attn,
                                  int N, int S, int M, int D, int Lq, int L, int P, int dtype) {
    LevelTable lt;
    if (make_levels(shapes_hw, level_start, L, S, lt) != PDB_OK) return -1;
    const int64_t total = (int64_t)N * Lq * M * D;
    const unsigned gblocks = (unsigned)((total + 255) / 256);
    const size_t taps = (size_t)N * Lq * M * L * P;
    if (dtype == 1) {
        std::fill((double*)grad_loc, (double*)grad_loc + 2 * taps, 0.0);
        std::fill((double*)grad_attn, (double*)grad_attn + taps, 0.0);
        launch(dim3(gblocks), dim3(256), [&] {
            msda_bwd_generic<double>((const double*)value, lt, (const double*)loc, (const double*)attn, (const double*)grad_out,
                                     (double*)grad_value, (double*)grad_loc, (double*)grad_attn, total, S, M, D, Lq, L, P);
        });
        return 1;
    }
    if (D == 32 && P == 4 && L * P <= 16 && levels_at_least_2x2(lt, L)) {
        const int64_t slots = (int64_t)N * Lq * M;
        PatchTable pt;
        const bool tiled = Lq == S;
        int64_t blocks = make_patches(lt, L, N, M, pt);
        if (!tiled) blocks = (slots + kTileSlots - 1) / kTileSlots;
        const size_t smem = sizeof(float4) * kTileSlots * (L * P * 4 + 1);
        launch(dim3((unsigned)blocks), dim3(kTileSlots * 8), smem, [&] {
            msda_bwd_tiled<4>((const float*)value, lt, pt, (const float*)loc, (const float*)attn, (const float*)grad_out,
                              (float*)grad_value, (float*)grad_loc, (float*)grad_attn, slots, S, M, Lq, L, tiled ? 1 : 0);
        });
        return tiled ? 3 : 4;
    }
    if (D == 32 && P == 4) {
        const int64_t slots = (int64_t)N * Lq * M;
        const int64_t blocks = (slots + kSlotsPerCta - 1) / kSlotsPerCta;
        const int LP = L * P;
        const size_t smem = sizeof(float) * kSlotsPerCta * ((LP * 2 + 2) + (LP + 1) + LP * 2 + LP);
        launch(dim3((unsigned)blocks), dim3(kFastThreads), smem, [&] {
            msda_bwd_d32<4>((const float*)value, lt, (const float*)loc, (const float*)attn, (const float*)grad_out,
                            (float*)grad_value, (float*)grad_loc, (float*)grad_attn, slots, S, M, Lq, L);
        });
        return 2;
    }
    std::fill((float*)grad_loc, (float*)grad_loc + 2 * taps, 0.f);
    std::fill((float*)grad_attn, (float*)grad_attn + taps, 0.f);
    launch(dim3(gblocks), dim3(256), [&] {
        msda_bwd_generic<float>((const float*)value, lt, (const float*)loc, (const float*)attn, (const float*)grad_out,
                                (float*)grad_value, (float*)grad_loc, (float*)grad_attn, total, S, M, D, Lq, L, P);
    });
    return 1;
}
