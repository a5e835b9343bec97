// TEST INFRASTRUCTURE ONLY.  msda_kernel_host.cpp behind entry points with the C ABI's return convention (0 = launched),
// including the zero fill of grad_value that pdb_msda_backward performs, for whole-model runs on the CPU tier.
#define host_msda_forward host_msda_forward_path
#define host_msda_backward host_msda_backward_path
#include "msda_kernel_host.cpp"
#undef host_msda_forward
#undef host_msda_backward

extern "C" int host_msda_forward(const void* value, const int64_t* shapes_hw, const int64_t* level_start, const void* loc,
                                 const void* attn, void* out, int N, int S, int M, int D, int Lq, int L, int P, int dtype) {
    return host_msda_forward_path(value, shapes_hw, level_start, loc, attn, out, N, S, M, D, Lq, L, P, dtype) > 0 ? 0 : -1;
}

extern "C" int host_msda_backward(const void* value, const int64_t* shapes_hw, const int64_t* level_start, const void* loc,
                                  const void* attn, const void* grad_out, void* grad_value, void* grad_loc, void* grad_attn,
                                  int N, int S, int M, int D, int Lq, int L, int P, int dtype) {
    const size_t n = (size_t)N * S * M * D;
    if (dtype == 1) std::fill((double*)grad_value, (double*)grad_value + n, 0.0);
    else std::fill((float*)grad_value, (float*)grad_value + n, 0.f);
    return host_msda_backward_path(value, shapes_hw, level_start, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, S, M, D,
                                   Lq, L, P, dtype) > 0 ? 0 : -1;
}
