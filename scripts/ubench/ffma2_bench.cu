// Micro-benchmark: issue rate of packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a.
// One CTA of NW warps per SM; every thread runs ITER iterations of 16 independent accumulator chains.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    unsigned long long dd, aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
}
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float2 acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 a = make_float2(1.0001f + threadIdx.x * 1e-7f, 0.9999f), b = make_float2(0.5f, 0.25f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(a.x, acc[i].x, b.x); acc[i].y = fmaf(a.y, acc[i].y, b.y); }
            else { float2 d = b; unsigned long long dd, aa, cc;
                   asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
                   asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(acc[i].x), "f"(acc[i].y));
                   asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
                   asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(cc) : "l"(aa), "l"(cc), "l"(dd));
                   asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i].x), "=f"(acc[i].y) : "l"(cc)); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    for (int nw : {4, 8, 16}) for (int mode = 0; mode < 2; ++mode) {
        for (int r = 0; r < 2; ++r) { if (mode == 0) k<0><<<148, nw * 32>>>(out, iters, cyc); else k<1><<<148, nw * 32>>>(out, iters, cyc); cudaDeviceSynchronize(); }
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double fl = (double)iters * 32 * nw * 32;   // scalar FMAs per CTA
        printf("%s warps %2d: %lld cycles, %.1f fp32 FMA / clk / SM\n", mode ? "FFMA2" : "FFMA ", nw, h[0], fl / h[0]);
    }
    return 0;
}
