// Mask-head einsum  out[b,q,p] = sum_c embed[b,q,c] * feat[b,p,c]  — C-ABI entry points and the two
// gradient products.  Reference: torch.einsum("bqc,bchw->bqhw") at mask2former_transformer_decoder.py:449
// and part_distillation_transformer_decoder.py:244 (autograd supplies the backward there).
//
// The feature map is pixel-major (B, HW, C) ("channels last"): the 1x1 mask_features convolution that
// produces it is a GEMM over pixels, so this layout is free, and it makes both einsum operands K-major.
// Forward and both gradient products run on the tcgen05 3xTF32 GEMM of gemm_tc.cu (operand layouts are
// described to TMA / UMMA in place, nothing is transposed in memory):
//   out (Q x HW)^T      : A = feat (pixels x C, K-major),  B = embed (Q x C, K-major), transposed store
//   grad_feat (HW x C)  = grad_out^T (HW x Q) x embed (Q x C)     A, B MN-major
//   grad_embed (Q x C)  = grad_out (Q x HW) x feat (HW x C)        B MN-major, split-K over HW, red.add
// The shared-memory tiled fp32 FFMA GEMM below (128x128x16 tiles, 8x8 register micro-tiles) only serves
// shapes whose rows are not 16-byte multiples (C or HW not divisible by 4).
#include <algorithm>
#include "common.cuh"

namespace pdb {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, GEMM_THREADS = 256;

enum { STORE = 0, ACCUM = 1, ATOMIC = 2 };

// C[m][n] (+)= sum_k A(m,k) * B(k,n) for one batch item (blockIdx.z / ksplit) and one K slice.
//   A(m,k) at A[m*lda + k] if A_KC (k contiguous) else A[k*lda + m]
//   B(k,n) at B[k*ldb + n] if B_NC (n contiguous) else B[n*ldb + k]
template <bool A_KC, bool B_NC, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS)
tile_gemm(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, int M, int N, int64_t K,
          int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC, int ksplit,
          int64_t kchunk) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int b = blockIdx.z / ksplit;
    const int ks = blockIdx.z - b * ksplit;
    const int64_t k_begin = (int64_t)ks * kchunk;
    const int64_t k_end = min(K, k_begin + kchunk);
    A += (int64_t)b * strideA;
    Bm += (int64_t)b * strideB;
    C += (int64_t)b * strideC;
    const int m0 = blockIdx.y * BM;
    const int64_t n0 = (int64_t)blockIdx.x * BN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;     // tx -> columns, ty -> rows

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) acc[i][jn] = 0.f;

    for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
        // ---- A tile (BM x BK) -> As[k][m]
        if (A_KC) {
#pragma unroll
            for (int it = 0; it < (BM * BK) / GEMM_THREADS; ++it) {
                int e = it * GEMM_THREADS + tid;
                int kk = e & (BK - 1), mm = e >> 4;
                int gm = m0 + mm;
                int64_t gk = k0 + kk;
                As[kk][mm] = (gm < M && gk < k_end) ? __ldg(A + (int64_t)gm * lda + gk) : 0.f;
            }
        } else {
#pragma unroll
            for (int it = 0; it < (BM * BK) / GEMM_THREADS; ++it) {
                int e = it * GEMM_THREADS + tid;
                int mm = e & (BM - 1), kk = e >> 7;
                int gm = m0 + mm;
                int64_t gk = k0 + kk;
                As[kk][mm] = (gm < M && gk < k_end) ? __ldg(A + gk * lda + gm) : 0.f;
            }
        }
        // ---- B tile (BK x BN) -> Bs[k][n]
        if (B_NC) {
#pragma unroll
            for (int it = 0; it < (BN * BK) / GEMM_THREADS; ++it) {
                int e = it * GEMM_THREADS + tid;
                int nn = e & (BN - 1), kk = e >> 7;
                int64_t gn = n0 + nn, gk = k0 + kk;
                Bs[kk][nn] = (gn < N && gk < k_end) ? __ldg(Bm + gk * ldb + gn) : 0.f;
            }
        } else {
#pragma unroll
            for (int it = 0; it < (BN * BK) / GEMM_THREADS; ++it) {
                int e = it * GEMM_THREADS + tid;
                int kk = e & (BK - 1), nn = e >> 4;
                int64_t gn = n0 + nn, gk = k0 + kk;
                Bs[kk][nn] = (gn < N && gk < k_end) ? __ldg(Bm + gn * ldb + gk) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], bv[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int jn = 0; jn < TN; ++jn) acc[i][jn] = fmaf(a[i], bv[jn], acc[i][jn]);
        }
        __syncthreads();
    }
    // rows: ty*4+{0..3} and 64+ty*4+{0..3}; cols: tx*4+{0..3} and 64+tx*4+{0..3}
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) {
            int64_t gn = n0 + (jn < 4 ? tx * 4 + jn : 64 + tx * 4 + (jn - 4));
            if (gn >= N) continue;
            float* p = C + (int64_t)gm * ldc + gn;
            if (MODE == STORE) *p = acc[i][jn];
            else if (MODE == ACCUM) *p += acc[i][jn];
            else atomicAdd(p, acc[i][jn]);
        }
    }
}

}  // namespace pdb

namespace pdb {
int gemm_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M, int N, int K, int batch,
                int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn, int b_mn,
                int c_trans, int relu, int accumulate, int ksplit, cudaStream_t st);
}

using namespace pdb;

extern "C" int pdb_mask_einsum_forward(const float* embed, const float* embed_lo, const float* feat, float* out, int B,
                                       int Q, int C, int64_t HW, void* stream) {
    PDB_REQUIRE(embed && feat && out, "mask_einsum_forward: null pointer");
    PDB_REQUIRE(B > 0 && Q > 0 && C > 0 && HW > 0, "mask_einsum_forward: non-positive dimension");
    PDB_REQUIRE(C % 4 == 0, "mask_einsum_forward: C=%d must be a multiple of 4 (16-byte TMA rows)", C);
    PDB_REQUIRE((reinterpret_cast<uintptr_t>(embed) & 15) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0,
                "mask_einsum_forward: embed / feat must be 16-byte aligned");
    PDB_REQUIRE(HW < (1ll << 31), "mask_einsum_forward: too many pixels");
    // out[b][q][p] = sum_c feat[b][p][c] * embed[b][q][c]: M = pixels (TMEM lanes), N = queries, both operands
    // K-major, transposed store (32 consecutive pixels per warp store)
    return gemm_tf32x3(feat, embed, embed_lo, out, nullptr, (int)HW, Q, C, B, C, C, HW, (int64_t)HW * C, (int64_t)Q * C,
                       (int64_t)Q * HW, 0, 0, 1, 0, 0, 1, as_stream(stream));
}

extern "C" int pdb_mask_einsum_backward(const float* embed, const float* feat, const float* grad_out,
                                        float* grad_embed, float* grad_feat, int accumulate, int B, int Q, int C,
                                        int64_t HW, void* stream) {
    PDB_REQUIRE(embed && feat && grad_out, "mask_einsum_backward: null pointer");
    PDB_REQUIRE(B > 0 && Q > 0 && C > 0 && HW > 0, "mask_einsum_backward: non-positive dimension");
    cudaStream_t st = as_stream(stream);
    // tensor-core path (tcgen05 3xTF32, operands described in place as MN-major): needs 16-byte rows
    const bool tc_ok = C % 4 == 0 && HW % 4 == 0 && HW < (1ll << 31) &&
                       ((reinterpret_cast<uintptr_t>(embed) | reinterpret_cast<uintptr_t>(feat) |
                         reinterpret_cast<uintptr_t>(grad_out)) & 15) == 0;
    if (tc_ok) {
        if (grad_feat) {
            // grad_feat[b][p][c] (+)= sum_q grad_out[b][q][p] * embed[b][q][c]: A(m=p,k=q) and B(n=c,k=q) are MN-major
            PDB_TRY(gemm_tf32x3(grad_out, embed, nullptr, grad_feat, nullptr, (int)HW, C, Q, B, HW, C, C, (int64_t)Q * HW,
                                (int64_t)Q * C, (int64_t)C * HW, 1, 1, 0, 0, accumulate ? 1 : 0, 1, st));
        }
        if (grad_embed) {
            // grad_embed[b][q][c] = sum_p grad_out[b][q][p] * feat[b][p][c]: A K-major, B(n=c,k=p) MN-major, split-K over HW
            cudaMemsetAsync(grad_embed, 0, sizeof(float) * (size_t)B * Q * C, st);
            int tiles = B * ((Q + 127) / 128) * ((C + 127) / 128);
            int ksplit = (int)std::max<int64_t>(1, std::min<int64_t>((HW + 1023) / 1024, (2 * kNumSMs + tiles - 1) / tiles));
            PDB_TRY(gemm_tf32x3(grad_out, feat, nullptr, grad_embed, nullptr, Q, C, (int)HW, B, HW, C, C, (int64_t)Q * HW,
                                (int64_t)C * HW, (int64_t)Q * C, 0, 1, 0, 0, 1, ksplit, st));
        }
        return PDB_OK;
    }
    if (grad_feat) {
        // grad_feat[b] (HW x C) (+)= grad_out[b]^T (HW x Q) * embed[b] (Q x C)
        dim3 grid((unsigned)((C + BN - 1) / BN), (unsigned)((HW + BM - 1) / BM), (unsigned)B);
        if (accumulate)
            tile_gemm<false, true, ACCUM><<<grid, GEMM_THREADS, 0, st>>>(
                grad_out, embed, grad_feat, (int)HW, C, Q, HW, C, C, (int64_t)Q * HW, (int64_t)Q * C,
                (int64_t)C * HW, 1, Q);
        else
            tile_gemm<false, true, STORE><<<grid, GEMM_THREADS, 0, st>>>(
                grad_out, embed, grad_feat, (int)HW, C, Q, HW, C, C, (int64_t)Q * HW, (int64_t)Q * C,
                (int64_t)C * HW, 1, Q);
        PDB_TRY(launched("mask_einsum_backward(grad_feat)"));
    }
    if (grad_embed) {
        // grad_embed[b] (Q x C) = grad_out[b] (Q x HW) * feat[b] (HW x C); split-K over HW
        cudaMemsetAsync(grad_embed, 0, sizeof(float) * (size_t)B * Q * C, st);
        int64_t kchunk = 1024;
        int ksplit = (int)((HW + kchunk - 1) / kchunk);
        dim3 grid((unsigned)((C + BN - 1) / BN), (unsigned)((Q + BM - 1) / BM), (unsigned)(B * ksplit));
        tile_gemm<true, true, ATOMIC><<<grid, GEMM_THREADS, 0, st>>>(
            grad_out, feat, grad_embed, Q, C, HW, HW, C, C, (int64_t)Q * HW, (int64_t)C * HW, (int64_t)Q * C, ksplit,
            kchunk);
        PDB_TRY(launched("mask_einsum_backward(grad_embed)"));
    }
    return PDB_OK;
}
