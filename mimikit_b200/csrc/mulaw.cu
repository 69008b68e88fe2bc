// mu-law compand/quantise and expand: vectorised HBM-bandwidth kernels (sm_100a).
//
// Reference arithmetic: mimikit/features/functionals.py:330-338 (MuLawCompress.torch_func) and :361-369
// (MuLawExpand.torch_func), evaluated by torch in fp32.  The compress kernel reproduces the CPU reference bit
// for bit: same op order, IEEE division, and Sleef's log1pf (mmk::p_log1pf) — see oracle/c/oracle_feat.c.
//
// Roofline: HBM.  Algorithmic bytes per sample: compress 4 (fp32 in) + 8 (int64 out) = 12 B; the u8 variant
// 4 + 1 = 5 B; expand 8 + 4 = 12 B.  Each thread handles 4 consecutive samples per iteration (one 16-byte load,
// two 16-byte stores), grid-stride over a grid of 148 SMs x 8 resident CTAs.
#include "common.cuh"
#include "../../include/mmk_b200.h"

namespace mmk {

__device__ __forceinline__ float mulaw_level(float v, float mu, float C, float denom) {
    float a = __fmul_rn(__fmul_rn(mu, fabsf(v)), C);
    float xm = __fdiv_rn(__fmul_rn(p_sign(v), p_log1pf(a)), denom);
    float r = __fdiv_rn(__fadd_rn(xm, 1.0f), 2.0f);
    r = __fadd_rn(__fmul_rn(r, mu), 0.5f);
    return r;  // caller truncates toward zero, as .to(torch.int64) does
}

template <typename OutT>
__global__ void __launch_bounds__(256) mulaw_compress_kernel(const float* __restrict__ x, OutT* __restrict__ q,
                                                             size_t n, float mu, float C) {
    const float denom = p_log1pf(__fmul_rn(mu, C));
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
        long long r0 = (long long)mulaw_level(v.x, mu, C, denom), r1 = (long long)mulaw_level(v.y, mu, C, denom);
        long long r2 = (long long)mulaw_level(v.z, mu, C, denom), r3 = (long long)mulaw_level(v.w, mu, C, denom);
        if constexpr (sizeof(OutT) == 8) {
            longlong2* o = reinterpret_cast<longlong2*>(q) + 2 * i;
            __stcs(o, make_longlong2(r0, r1));
            __stcs(o + 1, make_longlong2(r2, r3));
        } else {
            uchar4 o = make_uchar4((unsigned char)r0, (unsigned char)r1, (unsigned char)r2, (unsigned char)r3);
            reinterpret_cast<uchar4*>(q)[i] = o;
        }
    }
    // ragged tail (n % 4 samples)
    size_t t = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) q[t] = (OutT)(long long)mulaw_level(x[t], mu, C, denom);
}

__global__ void __launch_bounds__(256) mulaw_expand_kernel(const long long* __restrict__ q, float* __restrict__ x,
                                                           size_t n, float mu, float C) {
    const float l1p = p_log1pf(__fmul_rn(mu, C));
    const float muC = __fmul_rn(mu, C);
    auto one = [&](long long idx) {
        float v = (float)idx;
        float xx = __fsub_rn(__fmul_rn(__fdiv_rn(v, mu), 2.0f), 1.0f);
        float e = p_expf(__fmul_rn(fabsf(xx), l1p));
        return __fdiv_rn(__fmul_rn(p_sign(xx), __fsub_rn(e, 1.0f)), muC);
    };
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const longlong2* p = reinterpret_cast<const longlong2*>(q) + 2 * i;
        longlong2 a = __ldcs(p), b = __ldcs(p + 1);
        __stcs(reinterpret_cast<float4*>(x) + i, make_float4(one(a.x), one(a.y), one(b.x), one(b.y)));
    }
    size_t t = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t] = one(q[t]);
}

static int feature_grid(size_t work_items) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t want = (work_items + 255) / 256;
    size_t cap = (size_t)sms * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace mmk

using namespace mmk;

extern "C" int mmk_mulaw_compress(const float* d_x, int64_t* d_q, size_t n, int q_levels, float compression,
                                  void* stream) {
    MMK_CHECK(q_levels >= 2, "mmk_mulaw_compress: q_levels must be >= 2");
    if (n == 0) return 0;
    MMK_CHECK(d_x && d_q, "mmk_mulaw_compress: null pointer");
    MMK_CHECK(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_q % 16) == 0, "mmk_mulaw_compress: buffers must be 16-byte aligned");
    mulaw_compress_kernel<long long><<<feature_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(
        d_x, reinterpret_cast<long long*>(d_q), n, (float)q_levels - 1.0f, compression);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmk_mulaw_compress_u8(const float* d_x, uint8_t* d_q, size_t n, int q_levels, float compression,
                                     void* stream) {
    MMK_CHECK(q_levels >= 2 && q_levels <= 256, "mmk_mulaw_compress_u8: q_levels must be in [2, 256]");
    if (n == 0) return 0;
    MMK_CHECK(d_x && d_q, "mmk_mulaw_compress_u8: null pointer");
    MMK_CHECK(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_q % 4) == 0, "mmk_mulaw_compress_u8: misaligned buffers");
    mulaw_compress_kernel<unsigned char><<<feature_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(
        d_x, d_q, n, (float)q_levels - 1.0f, compression);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmk_mulaw_expand(const int64_t* d_q, float* d_x, size_t n, int q_levels, float compression,
                                void* stream) {
    MMK_CHECK(q_levels >= 2, "mmk_mulaw_expand: q_levels must be >= 2");
    if (n == 0) return 0;
    MMK_CHECK(d_x && d_q, "mmk_mulaw_expand: null pointer");
    MMK_CHECK(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_q % 16) == 0, "mmk_mulaw_expand: buffers must be 16-byte aligned");
    mulaw_expand_kernel<<<feature_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long*>(d_q), d_x, n, (float)q_levels - 1.0f, compression);
    MMK_CUDA(cudaGetLastError());
    return 0;
}
