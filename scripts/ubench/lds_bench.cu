// Micro-benchmark: shared-memory load / shuffle throughput per SM for the access patterns of wavenet6.cu (8 warps per CTA).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bench lds_bench.cu && ./lds_bench
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int MODE>
__global__ void k(float* out, long long* cyc, int stride) {
    __shared__ __align__(16) float sm[8192];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = (float)i * 1e-3f;
    __syncthreads();
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const float4* s4 = reinterpret_cast<const float4*>(sm);
    const float2* s2 = reinterpret_cast<const float2*>(sm);
    int idx;
    if (MODE == 0) idx = (lane & 15);            // LDS.128, 16 distinct float4, halves duplicated (x pattern)
    if (MODE == 1) idx = tid;                     // LDS.128, all distinct (W pattern)
    if (MODE == 2) idx = 0;                       // LDS.128, one address
    if (MODE == 3) idx = (lane & 15);            // LDS.64, 16 distinct
    if (MODE == 4) idx = (lane & 15);            // LDS.32, 16 distinct
    if (MODE == 5) idx = lane;                    // LDS.32, 32 distinct
    if (MODE == 6) idx = (lane & 7);             // LDS.128, 8 distinct float4 in every quarter-warp
    if (MODE == 8) idx = lane;                    // LDS.128, 32 distinct but the same for every warp
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        const int o = ((i * stride) & 15) * 64;   // the stride is a run-time value: nothing can be hoisted
        if (MODE == 0 || MODE == 2 || MODE == 6 || MODE == 8) { float4 v = s4[o + idx]; a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
        if (MODE == 1) { float4 v = s4[((o + idx) & 2047)]; a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
        if (MODE == 3) { float2 v = s2[o + idx]; a0 += v.x; a1 += v.y; }
        if (MODE == 4 || MODE == 5) { a0 += sm[o + idx]; }
        if (MODE == 7) { a0 += __shfl_xor_sync(0xffffffffu, a1, 8); a1 += __shfl_xor_sync(0xffffffffu, a2, 4); a2 += __shfl_xor_sync(0xffffffffu, a3, 2); a3 += __shfl_xor_sync(0xffffffffu, a0, 1); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = a0 + a1 + a2 + a3;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int nthreads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 64);
    k<MODE><<<1, nthreads>>>(out, cyc, 1); k<MODE><<<1, nthreads>>>(out, cyc, 1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-52s warps %2d: %.2f cycles per warp-instruction (SM-wide)\n", name, nthreads / 32, (double)h / ITERS / (nthreads / 32));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int nt : {256, 512}) {
        run<0>("LDS.128 16 distinct float4, halves duplicated", nt);
        run<6>("LDS.128 8 distinct float4 per quarter-warp", nt);
        run<1>("LDS.128 all lanes distinct", nt);
        run<8>("LDS.128 32 distinct, same for every warp", nt);
        run<2>("LDS.128 one address", nt);
        run<3>("LDS.64 16 distinct", nt);
        run<4>("LDS.32 16 distinct", nt);
        run<5>("LDS.32 32 distinct", nt);
        run<7>("4 x SHFL.BFLY per iteration (value / 4 = per shuffle)", nt);
    }
    return 0;
}
