// Small-batch fp32 contraction over shared memory, shared by the persistent generation kernels (sm_100a).
//
// out[col][p] = sum_k W[k][col] * x[k][p] for `ncolp` (multiple of 4) output columns and PB prompts, W and x both
// resident in shared memory, k-major.  The (PB/2) x (ncolp/4) register tiles of 2 prompts x 4 columns are
// replicated over `nslice` K-slices (slice s takes k = s, s + nslice, ...); partial sums meet in shared memory
// and are added in slice order, so the summation order is fixed (deterministic run to run).
#pragma once
#include <cuda_runtime.h>

namespace mmk {

template <int PB, int NT>
struct TileGemm {
    int slice, nslice, pp, cq, lout;
    bool active;

    __device__ __forceinline__ explicit TileGemm(int ncolp) {
        lout = (PB / 2) * (ncolp / 4);
        nslice = NT / lout;
        slice = threadIdx.x / lout;
        const int tile = threadIdx.x % lout;
        pp = tile % (PB / 2);
        cq = tile / (PB / 2);
        active = slice < nslice;
    }

    __device__ __forceinline__ void accum(float (&acc)[8], const float* __restrict__ Ws, int ldw,
                                          const float* __restrict__ xs, int K) const {
        if (!active) return;
        const float* wp = Ws + cq * 4;
        const float* xp = xs + pp * 2;
#pragma unroll 4
        for (int k = slice; k < K; k += nslice) {
            const float4 w = *reinterpret_cast<const float4*>(wp + k * ldw);
            const float2 x = *reinterpret_cast<const float2*>(xp + k * PB);
            acc[0] = fmaf(x.x, w.x, acc[0]); acc[1] = fmaf(x.x, w.y, acc[1]);
            acc[2] = fmaf(x.x, w.z, acc[2]); acc[3] = fmaf(x.x, w.w, acc[3]);
            acc[4] = fmaf(x.y, w.x, acc[4]); acc[5] = fmaf(x.y, w.y, acc[5]);
            acc[6] = fmaf(x.y, w.z, acc[6]); acc[7] = fmaf(x.y, w.w, acc[7]);
        }
    }

    __device__ __forceinline__ void store(const float (&acc)[8], float* part) const {
        if (active) {
            float4* d = reinterpret_cast<float4*>(part + (size_t)threadIdx.x * 8);
            d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    }

    // sum of the partials of output (col, p) — call after __syncthreads()
    __device__ __forceinline__ static float reduce(const float* part, int ncolp, int col, int p) {
        const int lout = (PB / 2) * (ncolp / 4), nslice = NT / lout;
        const int tile = (col >> 2) * (PB / 2) + (p >> 1);
        const float* q = part + tile * 8 + (p & 1) * 4 + (col & 3);
        float s = 0.0f;
        for (int sl = 0; sl < nslice; ++sl) s += q[(size_t)sl * lout * 8];
        return s;
    }
};

}  // namespace mmk
