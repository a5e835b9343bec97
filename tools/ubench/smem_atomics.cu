// Micro-benchmark (not product code): shared-memory update primitives on sm_100a, cycles per warp instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
// mode 0: float atomicAdd (CAS loop), conflict-free within the warp (lane -> distinct bank), random row per warp
// mode 1: int atomicAdd (native ATOMS.ADD, result unused), same addressing
// mode 2: int atomicAdd with the result used
// mode 3: plain LDS + FADD + STS (no atomicity), same addressing
// mode 4: LDS.128 only (reference)
template <int MODE>
__global__ void k(float* out, long long* cycles, int rows) {
    extern __shared__ float s[];
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) s[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned r = warp * 2654435761u + blockIdx.x * 40503u + 12345u;
    float acc = 0.f;
    int iacc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        r = r * 1664525u + 1013904223u;
        const int row = (r >> 8) % rows;
        float* p = s + row * 32 + lane;
        if (MODE == 0) atomicAdd(p, 1.0f);
        if (MODE == 1) atomicAdd(reinterpret_cast<int*>(p), 1);
        if (MODE == 2) iacc += atomicAdd(reinterpret_cast<int*>(p), 1);
        if (MODE == 3) *p = *p + 1.0f;
        if (MODE == 4) { float4 v = *reinterpret_cast<float4*>(s + ((row * 32 + lane * 4) % (rows * 32 - 4) & ~3)); acc += v.x + v.w; }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + iacc + s[threadIdx.x];
}
template <int MODE>
void run(const char* name, int threads, int ctas_per_sm) {
    int rows = 256;
    float* out; long long* cyc;
    int blocks = 148 * ctas_per_sm;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads, rows * 32 * 4>>>(out, cyc, rows);
    cudaDeviceSynchronize();
    k<MODE><<<blocks, threads, rows * 32 * 4>>>(out, cyc, rows);
    cudaDeviceSynchronize();
    long long h[4096];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    int warps_per_sm = threads / 32 * ctas_per_sm;
    printf("%-28s threads=%4d ctas/sm=%d: %.1f cyc per iteration per warp; %.2f cyc per warp-instruction per SM\n", name, threads,
           ctas_per_sm, avg / ITERS, avg / ITERS / warps_per_sm);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int t : {256, 1024}) {
        run<0>("float atomicAdd (CAS loop)", t, 1);
        run<1>("int atomicAdd (RED form)", t, 1);
        run<2>("int atomicAdd (result used)", t, 1);
        run<3>("LDS+FADD+STS", t, 1);
        run<4>("LDS.128", t, 1);
    }
    return 0;
}
