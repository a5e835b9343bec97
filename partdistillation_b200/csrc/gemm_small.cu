// Short-and-wide fp32 contractions of the transformer decoder — y = x W^T + b and dx = dy W with only B * Q = 200 rows
// (mask2former_transformer_decoder.py:148-208: self- / cross-attention projections, FFN, mask MLP; ~220 launches per step).
// The persistent tcgen05 kernel of gemm_tc.cu pays ~10 us of fixed cost per launch (TMEM allocation, 18-warp role set-up, TMA
// descriptor fetch, a 128-row tile that is 36 % empty) for 13-100 MFLOP; here the same 3xTF32 arithmetic (hi / lo operand split,
// fp32-accurate) runs as warp-level mma.sync.m16n8k8 on 16 x 64 / 32 x 64 tiles, with a 4-deep cp.async ring of 32-wide k-chunks and an
// optional split over K (red.add into a zero-filled C) so that ~150 CTAs are in flight whatever the shape.
//     C[m][n] (+)= sum_k A[m*lda + k] * B(n, k)  (+ bias[n]) (ReLU)
//     b_mn = 0: B(n, k) = B[n*ldb + k]   (nn.Linear forward: B = W)
//     b_mn = 1: B(n, k) = B[k*ldb + n]   (input gradient: B = W read in place)
#include "common.cuh"

namespace pdb {

constexpr int S_BN = 64, S_BK = 32;
constexpr int S_AS = 36;           // A tile row stride (floats): fragment reads [g][8ks + t] hit banks 4g + t
constexpr int S_BS_K = 36;         // K-major B tile [n][k]
constexpr int S_BS_MN = 72;        // MN-major B tile [k][n]: fragment reads [8ks + t][8j + g] hit banks 8t + g

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;           // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void s_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float s_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// BM = 16 (4 warps) or 32 (8 warps) rows per CTA; every warp owns a 16 x 16 block of the 64-column tile.  The k-chunks run through a
// 4-deep (BM = 32: 3-deep) cp.async ring (three / two chunks in flight: a chunk's 24 MMAs per warp are shorter than one L2 round trip).

template <bool B_MN, int BM>
__global__ void __launch_bounds__(BM * 8)
gemm_small_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, const float* __restrict__ bias,
                  int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int relu, int ksplit, int chunks_per_split) {
    constexpr int NT = BM * 8;
    constexpr int S_STAGES = BM == 16 ? 4 : 3;      // static shared memory: 45 KB / 40.5 KB
    constexpr int B_TILE = B_MN ? S_BK * S_BS_MN : S_BN * S_BS_K;
    __shared__ __align__(16) float sA[S_STAGES][BM * S_AS];
    __shared__ __align__(16) float sB[S_STAGES][B_TILE];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * S_BN, split = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 16, wn = (warp & 3) * 16;
    const int total_chunks = (K + S_BK - 1) / S_BK;
    const int c_begin = split * chunks_per_split, c_end = min(total_chunks, c_begin + chunks_per_split);

    auto load = [&](int buf, int chunk) {
        const int k0 = chunk * S_BK;
        // A: BM rows x 8 float4
#pragma unroll
        for (int e = threadIdx.x; e < BM * 8; e += NT) {
            const int r = e >> 3, c = (e & 7) * 4;
            const bool ok = m0 + r < M && k0 + c < K;
            cp_async16(&sA[buf][r * S_AS + c], ok ? A + (int64_t)(m0 + r) * lda + k0 + c : A, ok);
        }
        if (!B_MN) {        // [n][k]: 64 rows x 8 float4
#pragma unroll
            for (int i = 0; i < 512 / NT; ++i) {
                const int e = threadIdx.x + i * NT, r = e >> 3, c = (e & 7) * 4;
                const bool ok = n0 + r < N && k0 + c < K;
                cp_async16(&sB[buf][r * S_BS_K + c], ok ? B + (int64_t)(n0 + r) * ldb + k0 + c : B, ok);
            }
        } else {            // [k][n]: 32 rows x 16 float4
#pragma unroll
            for (int i = 0; i < 512 / NT; ++i) {
                const int e = threadIdx.x + i * NT, r = e >> 4, c = (e & 15) * 4;
                const bool ok = k0 + r < K && n0 + c < N;
                cp_async16(&sB[buf][r * S_BS_MN + c], ok ? B + (int64_t)(k0 + r) * ldb + n0 + c : B, ok);
            }
        }
    };

    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int s = 0; s < S_STAGES - 1; ++s) {
        if (c_begin + s < c_end) load(s, c_begin + s);
        cp_async_commit();                       // one group per ring slot, empty or not: the wait below counts groups
    }
    for (int c = c_begin; c < c_end; ++c) {
        const int buf = (c - c_begin) % S_STAGES;
        cp_async_wait<S_STAGES - 2>();           // chunk c has landed (this thread's part) ...
        __syncthreads();                         // ... everybody's part, and everybody is done with chunk c - 1's buffer
        if (c + S_STAGES - 1 < c_end) load((buf + S_STAGES - 1) % S_STAGES, c + S_STAGES - 1);
        cp_async_commit();
        const float* a = sA[buf] + wm * S_AS;
        const float* b = sB[buf];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float af[4] = {a[g * S_AS + ks * 8 + t], a[(g + 8) * S_AS + ks * 8 + t], a[g * S_AS + ks * 8 + t + 4],
                                 a[(g + 8) * S_AS + ks * 8 + t + 4]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ah[i] = __float_as_uint(af[i]);
                al[i] = __float_as_uint(s_lo(af[i]));
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float b0, b1;
                if (!B_MN) {
                    b0 = b[(wn + j * 8 + g) * S_BS_K + ks * 8 + t];
                    b1 = b[(wn + j * 8 + g) * S_BS_K + ks * 8 + t + 4];
                } else {
                    b0 = b[(ks * 8 + t) * S_BS_MN + wn + j * 8 + g];
                    b1 = b[(ks * 8 + t + 4) * S_BS_MN + wn + j * 8 + g];
                }
                s_mma(acc[j], ah, __float_as_uint(b0), __float_as_uint(b1));
                s_mma(acc[j], al, __float_as_uint(b0), __float_as_uint(b1));
                s_mma(acc[j], ah, __float_as_uint(s_lo(b0)), __float_as_uint(s_lo(b1)));
            }
        }
    }
    // epilogue: thread holds (rows m0 + wm + g | + 8, columns n0 + wn + 8j + 2t, + 1)
    const bool first = split == 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int col = n0 + wn + j * 8 + 2 * t;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = m0 + wm + g + hh * 8;
            if (row >= M) continue;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                if (col + cc >= N) continue;
                float v = acc[j][hh * 2 + cc];
                if (bias && first) v += __ldg(bias + col + cc);
                float* o = C + (int64_t)row * ldc + col + cc;
                if (ksplit > 1) {
                    atomicAdd(o, v);
                } else {
                    *o = relu ? fmaxf(v, 0.f) : v;
                }
            }
        }
    }
}

template <bool B_MN, int BM>
static void launch_small(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                         int64_t ldc, int relu, int ksplit, int per, cudaStream_t st) {
    dim3 grid((unsigned)((N + S_BN - 1) / S_BN), (unsigned)((M + BM - 1) / BM), (unsigned)ksplit);
    gemm_small_kernel<B_MN, BM><<<grid, BM * 8, 0, st>>>(A, B, C, bias, M, N, K, lda, ldb, ldc, relu, ksplit, per);
}

}  // namespace pdb

using namespace pdb;

// ksplit > 1 requires C zero-filled by the caller and relu == 0 (the activation needs the complete sum).
extern "C" int pdb_gemm_small_tf32x3(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int64_t lda,
                                     int64_t ldb, int64_t ldc, int b_mn, int relu, int ksplit, void* stream) {
    PDB_REQUIRE(A && B && C, "gemm_small: null pointer");
    PDB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, "gemm_small: K, lda, ldb must be multiples of 4");
    PDB_REQUIRE(!b_mn || N % 4 == 0, "gemm_small: N must be a multiple of 4 for an MN-major B");
    PDB_REQUIRE((((uintptr_t)A | (uintptr_t)B) & 15) == 0, "gemm_small: A and B must be 16-byte aligned");
    const int chunks = (K + S_BK - 1) / S_BK;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > chunks) ksplit = chunks;
    PDB_REQUIRE(ksplit == 1 || !relu, "gemm_small: split-K cannot apply the activation");
    const int per = (chunks + ksplit - 1) / ksplit;
    ksplit = (chunks + per - 1) / per;
    PDB_REQUIRE((M + 15) / 16 <= 65535, "gemm_small: too many rows (%d)", M);
    // 16-row tiles while 32-row tiles would leave SMs idle (the 200-row decoder products: 28 -> 52 CTAs of half the work each)
    const bool narrow = (int64_t)((N + S_BN - 1) / S_BN) * ((M + 31) / 32) * ksplit < 148;
    cudaStream_t st = as_stream(stream);
    if (b_mn) {
        if (narrow) launch_small<true, 16>(A, B, C, bias, M, N, K, lda, ldb, ldc, relu, ksplit, per, st);
        else launch_small<true, 32>(A, B, C, bias, M, N, K, lda, ldb, ldc, relu, ksplit, per, st);
    } else {
        if (narrow) launch_small<false, 16>(A, B, C, bias, M, N, K, lda, ldb, ldc, relu, ksplit, per, st);
        else launch_small<false, 32>(A, B, C, bias, M, N, K, lda, ldb, ldc, relu, ksplit, per, st);
    }
    return launched("gemm_small");
}
