// TEST INFRASTRUCTURE ONLY.  The Swin window-attention kernels of partdistillation_b200/csrc/window_attn.cu compiled for the
// host through cuda_on_cpu.h (tests/test_window_attn_kernels_host_cpu.py cuts both `namespace pdb` blocks into
// window_attn_section.inc); the entry points restate the two launchers of window_attn.cu.
#include "pdb_common_host.h"

#include "window_attn_section.inc"

using namespace pdb;
using cpu_cuda::launch;

extern "C" int host_window_attention_forward(const float* qkv, const float* bias, const float* mask, float* out, int Bw, int N,
                                             int heads, int d, int nW, float scale) {
    if (d != kWinD || !(Bw > 0 && heads > 0 && N > 0 && N <= 256)) return -1;
    if (mask && !(nW > 0 && Bw % nW == 0)) return -1;
    const int threads = ((N + 31) / 32) * 32;
    const size_t smem = sizeof(float) * 2 * N * kWinD;
    launch(dim3((unsigned)(Bw * heads)), dim3((unsigned)threads), smem,
           [&] { window_attention_kernel(qkv, bias, mask, out, N, heads, mask ? nW : 1, scale); });
    return 0;
}

extern "C" int host_swin_window_attention_forward(const float* qkv, const float* qkv_bias, const float* bias, float* out, int B,
                                                  int H, int W, int heads, int d, int ws, int shift, float scale) {
    if (d != kWinD || !(B > 0 && H > 0 && W > 0 && heads > 0 && ws > 0 && ws * ws <= 256 && shift >= 0 && shift < ws)) return -1;
    const int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
    const int64_t ctas = (int64_t)B * (Hp / ws) * (Wp / ws) * heads;
    const int N = ws * ws;
    const int threads = ((N + 31) / 32) * 32;
    const size_t smem = sizeof(float) * 2 * N * kWinD + sizeof(int) * N;
    launch(dim3((unsigned)ctas), dim3((unsigned)threads), smem,
           [&] { swin_window_attention_kernel(qkv, qkv_bias, bias, out, H, W, heads, ws, shift, Hp, Wp, scale); });
    return 0;
}
