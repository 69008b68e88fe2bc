// Output head tail shared by the generation kernels: learned temperature (networks/mlp.py:58-63), then
// CategoricalSampler (modules/targets.py:40-52) — argmax, or the noise-driven inverse-CDF draw that replaces
// torch.multinomial (contract in oracle/restate.py: sample_inverse_cdf).
#pragma once
#include "common.cuh"

namespace mmk {

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float mish_acc(float x) {
    float sp = x > 20.0f ? x : log1pf(expf(x));  // F.softplus, threshold 20
    return x * tanhf(sp);
}

// One warp decides one prompt from its SCALED logits z[0..Q) = raw / max(sigmoid(raw[Q]), min_temp) (shared memory, overwritten when
// sampling).  sample: draw with temperature T and uniform u, else argmax (lowest index wins ties, as torch.argmax).  All 32 lanes
// must call; every lane returns the decision.
__device__ __forceinline__ int decide_scaled_warp(float* z, int Q, bool sample, float T, float u) {
    const int lane = threadIdx.x & 31;
    float best = -INFINITY;
    int besti = 0x7fffffff;
    for (int c = lane; c < Q; c += 32) {
        const float v = z[c];
        if (v > best) { best = v; besti = c; }   // strided visit keeps the lowest index per lane
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (!sample) return besti;                                   // targets.py:42-43
    // inverse-CDF draw, blocked-scan order: lane i owns classes [i*n, (i+1)*n)
    float mm = __fdiv_rn(best, T);                               // max_k (z_k / T) = (max_k z_k) / T for T > 0
    if (!(T > 0.0f)) {                                           // non-positive T: take the true maximum
        mm = -INFINITY;
        for (int c = lane; c < Q; c += 32) mm = fmaxf(mm, __fdiv_rn(z[c], T));
        for (int o = 16; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    }
    __syncwarp();
    const int n = (Q + 31) / 32;
    float run = 0.0f;
    for (int i = 0; i < n; ++i) {
        const int c = lane * n + i;
        if (c < Q) {
            const float e = p_expf(__fsub_rn(__fdiv_rn(z[c], T), mm));
            run = __fadd_rn(run, e);
            z[c] = run;                                          // in-lane inclusive prefix
        }
    }
    float incl = run;
    for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = __fadd_rn(incl, up);
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.0f;
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    const float thr = __fmul_rn(u, total);
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const int c = lane * n + i;
        if (c < Q && __fadd_rn(excl, z[c]) <= thr) ++cnt;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return min(Q - 1, cnt);
}

// One warp decides one prompt.  z: shared-memory row of the Q+1 raw head outputs (overwritten).  logits_out:
// nullable global row of Q floats.  All 32 lanes must call; every lane returns the decision.
__device__ __forceinline__ int decide_warp(float* z, int Q, float min_temp, float* logits_out, bool sample, float T,
                                           float u) {
    const int lane = threadIdx.x & 31;
    const float temp = fmaxf(sigmoid_acc(z[Q]), min_temp);        // mlp.py:60-62
    __syncwarp();
    for (int c = lane; c < Q; c += 32) {
        const float v = z[c] / temp;
        z[c] = v;
        if (logits_out) __stcs(logits_out + c, v);
    }
    __syncwarp();
    return decide_scaled_warp(z, Q, sample, T, u);
}

}  // namespace mmk
