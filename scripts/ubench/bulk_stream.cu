// Micro-benchmark: how fast can every SM pull the SAME L2-resident rows with cp.async.bulk (the SampleRNN frame-tier pattern)?
// 128 CTAs x 32 threads; each streams `total` bytes in `chunk`-byte copies with `depth` copies in flight.
//   mode 0: all CTAs read one shared buffer in the same order     mode 1: same buffer, start rotated per CTA
//   mode 2: a private buffer per CTA (still L2 resident)          mode 3: shared buffer, multicast over a cluster of CS
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(unsigned dst, const void* src, unsigned bytes, unsigned bar, unsigned short mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) { unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_arrive_remote(unsigned rbar) { asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory"); }

__global__ void stream_kernel(const char* buf, size_t total, int chunk, int depth, int mode, int CS, int reps, long long* cycles, int NW) {
    extern __shared__ __align__(128) char smem[];
    const int w = threadIdx.x >> 5;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem) + w * 2 * depth;     // full[depth], empty[depth] per issuing warp
    char* stage = smem + 1024 + (size_t)w * depth * chunk;
    const int c = blockIdx.x, nchunk = (int)(total / chunk);
    const unsigned rank = CS > 1 ? cluster_ctarank() : 0;
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < depth; ++i) { mbar_init(smem_u32(bars + i), 1); mbar_init(smem_u32(bars + depth + i), CS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CS > 1) cluster_sync_all();
    const char* src = mode == 2 ? buf + (size_t)c * total : buf;
    const int rot = mode == 1 ? (int)(((long long)c * nchunk) / gridDim.x) : mode == 3 ? (int)(((long long)(c / CS) * nchunk) / (gridDim.x / CS)) : 0;
    long long t0 = clock64();
    if ((threadIdx.x & 31) == 0) {
        unsigned g = 0;
        auto issue = [&](unsigned gi) {
            const unsigned s = gi % depth, u = gi / depth;
            const int ch = (int)((gi * NW + w) % nchunk), x = ch + rot >= nchunk ? ch + rot - nchunk : ch + rot;
            if (mode == 3) {
                if (u > 0) while (!mbar_try_wait(smem_u32(bars + depth + s), (u - 1) & 1)) {}
                mbar_expect_tx(smem_u32(bars + s), chunk);
                const unsigned slice = chunk / CS;
                bulk_g2s_mc(smem_u32(stage + (size_t)s * chunk) + slice * rank, src + (size_t)x * chunk + slice * rank, slice, smem_u32(bars + s), (unsigned short)((1u << CS) - 1));
            } else {
                mbar_expect_tx(smem_u32(bars + s), chunk);
                bulk_g2s(smem_u32(stage + (size_t)s * chunk), src + (size_t)x * chunk, chunk, smem_u32(bars + s));
            }
        };
        const unsigned n = (unsigned)nchunk * reps / NW;
        for (unsigned i = 0; i < (unsigned)depth - 1 && i < n; ++i) issue(i);
        for (g = 0; g < n; ++g) {
            if (g + depth - 1 < n) issue(g + depth - 1);
            const unsigned s = g % depth, u = g / depth;
            while (!mbar_try_wait(smem_u32(bars + s), u & 1)) {}
            if (mode == 3) for (int r = 0; r < CS; ++r) mbar_arrive_remote(mapa(smem_u32(bars + depth + s), r));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[c] = clock64() - t0;
    if (CS > 1) cluster_sync_all();
}

int main() {
    const size_t total = 512 << 10;
    const int NC = 128;
    char* buf; cudaMalloc(&buf, (size_t)NC * total); cudaMemset(buf, 1, (size_t)NC * total);
    long long* cyc; cudaMalloc(&cyc, NC * 8);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 20;
    for (int mode : {0, 1})
        for (int chunk : {2048, 8192, 16384, 32768, 65536})
            for (int depth : {2, 4})
                for (int NW : {1, 2, 4, 8}) {
                    const int CS = 1;
                    if ((size_t)chunk * depth * NW > 190 * 1024) continue;
                    cudaLaunchConfig_t cfg{};
                    cfg.gridDim = dim3(CS == 8 ? 120 : NC); cfg.blockDim = dim3(32 * NW); cfg.dynamicSmemBytes = 1024 + (size_t)chunk * depth * NW;
                    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = CS > 1 ? 1 : 0;
                    for (int it = 0; it < 2; ++it) {
                        cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel, (const char*)buf, total, chunk, depth, mode, CS, reps, cyc, NW);
                        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); break; }
                        e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("run failed: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    long long h[NC]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                    long long mx = 0; for (int i = 0; i < (int)cfg.gridDim.x; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("mode %d chunk %5d depth %2d issuers %d: %.1f B/clk/SM (max over CTAs), %.0f cycles per 512 KB\n", mode, chunk, depth, NW,
                           (double)total * reps / mx, (double)mx / reps);
                }
    return 0;
}
