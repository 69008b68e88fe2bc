// mu-law compand/quantise and expand: vectorised HBM-bandwidth kernels (sm_100a).
//
// Reference arithmetic: mimikit/features/functionals.py:330-338 (MuLawCompress.torch_func) and :361-369
// (MuLawExpand.torch_func), evaluated by torch in fp32.
//
// Two implementations of the same function, bit for bit:
//   * exact kernels: the reference's op order, IEEE division and Sleef's log1pf/expf (mmk::p_log1pf, p_expf; see
//     oracle/c/oracle_feat.c).  ~90 instructions per sample: issue-bound at 59 % of the HBM roof.
//   * table kernels (the default): the quantiser restricted to |x| <= 1 is a monotone step function of x with
//     q_levels - 1 steps, so it is fully described by the smallest float that reaches each level.  Per
//     (device, q_levels, compression) those thresholds are found ONCE by bisection with the exact function, and the
//     table quantiser (MUFU.LG2 estimate of the level, then a +-1 correction against the two neighbouring thresholds)
//     is then PROVEN equal to the exact function on every one of the 2 130 706 434 floats in [-1, 1] by
//     mulaw_verify_kernel before it is ever used; samples outside [-1, 1] (and NaN) take the exact path inline.  If
//     the proof fails the exact kernel stays in use.  Expansion is a plain q_levels-entry table of the exact values.
//     ~14 instructions per sample: HBM-bound.
//
// Roofline: HBM.  Algorithmic bytes per sample: compress 4 (fp32 in) + 8 (int64 out) = 12 B; the u8 variant
// 4 + 1 = 5 B; expand 8 + 4 = 12 B.  Each thread handles 4 consecutive samples per iteration (one 16-byte load,
// two 16-byte stores), grid-stride over a grid of 148 SMs x 8 resident CTAs.
#include "common.cuh"
#include "../../include/mmk_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

namespace mmk {

constexpr int MULAW_WIDE_DEFAULT = 1;     // int64 compressor: 32-byte stores (measured 1.69 ms for 10 h; 16-byte 1.91,
                                          // 32-byte loads as well 1.87)
constexpr int MULAW_TABLE_MAX_Q = 2048;   // 16 KB of shared memory for the (lower, upper) threshold pairs

__device__ __forceinline__ float mulaw_level(float v, float mu, float C, float denom) {
    float a = __fmul_rn(__fmul_rn(mu, fabsf(v)), C);
    float xm = __fdiv_rn(__fmul_rn(p_sign(v), p_log1pf(a)), denom);
    float r = __fdiv_rn(__fadd_rn(xm, 1.0f), 2.0f);
    r = __fadd_rn(__fmul_rn(r, mu), 0.5f);
    return r;  // caller truncates toward zero, as .to(torch.int64) does
}

__device__ __noinline__ long long mulaw_exact_slow(float v, float mu, float C) {
    return (long long)mulaw_level(v, mu, C, p_log1pf(__fmul_rn(mu, C)));
}

// constants of the level estimate r ~ sign(x) * log2(1 + mu C |x|) * scale + bias
struct MuLawFast { float muC, scale, bias; int Q; };

// The table quantiser for |v| <= 1.  tab[k] = (smallest float with level >= k, smallest float with level >= k+1),
// tab[0].x = -inf, tab[Q-1].y = +inf.
__device__ __forceinline__ int mulaw_table_level(float v, const MuLawFast f, const float2* __restrict__ tab) {
    float y = __log2f(__fmaf_rn(f.muC, fabsf(v), 1.0f));
    float r = __fmaf_rn(copysignf(y, v), f.scale, f.bias);
    int k = min(max(__float2int_rz(r), 0), f.Q - 1);
    float2 t = tab[k];
    return k + (int)(v >= t.y) - (int)(v < t.x);
}

// ---- exact kernels --------------------------------------------------------------------------------------------
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

__device__ __forceinline__ float mulaw_expand_one(long long idx, float mu, float muC, float l1p) {
    float v = (float)idx;
    float xx = __fsub_rn(__fmul_rn(__fdiv_rn(v, mu), 2.0f), 1.0f);
    float e = p_expf(__fmul_rn(fabsf(xx), l1p));
    return __fdiv_rn(__fmul_rn(p_sign(xx), __fsub_rn(e, 1.0f)), muC);
}

__device__ __noinline__ float mulaw_expand_slow(long long idx, float mu, float C) {
    return mulaw_expand_one(idx, mu, __fmul_rn(mu, C), p_log1pf(__fmul_rn(mu, C)));
}

__global__ void __launch_bounds__(256) mulaw_expand_kernel(const long long* __restrict__ q, float* __restrict__ x,
                                                           size_t n, float mu, float C) {
    const float l1p = p_log1pf(__fmul_rn(mu, C));
    const float muC = __fmul_rn(mu, C);
    auto one = [&](long long idx) { return mulaw_expand_one(idx, mu, muC, l1p); };
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

// ---- table construction and proof --------------------------------------------------------------------------------
// floats <-> keys whose unsigned order is the numeric order (-0 sorts just below +0; both quantise alike)
__device__ __forceinline__ uint32_t f2key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// One thread per level k in [0, Q]: smallest float in [-1, 1] whose exact level is >= k (bisection on keys).
// thr[0] = -inf, thr[Q] = +inf.  Also fills the expansion table.  status[0] |= 1 if the end points are off.
__global__ void mulaw_build_kernel(float* __restrict__ thr, float* __restrict__ expand_tab, int Q, float mu, float C,
                                   unsigned* __restrict__ status) {
    const float denom = p_log1pf(__fmul_rn(mu, C));
    const float muC = __fmul_rn(mu, C);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= Q; k += gridDim.x * blockDim.x) {
        if (k < Q) expand_tab[k] = mulaw_expand_one(k, mu, muC, denom);
        if (k == 0) { thr[0] = __int_as_float(0xff800000); continue; }
        if (k == Q) { thr[Q] = __int_as_float(0x7f800000); continue; }
        uint32_t lo = f2key(-1.0f), hi = f2key(1.0f);   // invariant: level(lo) < k <= level(hi)
        long long l_lo = (long long)mulaw_level(-1.0f, mu, C, denom), l_hi = (long long)mulaw_level(1.0f, mu, C, denom);
        if (!(l_lo < k && k <= l_hi)) { atomicOr(status, 1u); thr[k] = __int_as_float(0x7fc00000); continue; }
        while (hi - lo > 1) {
            uint32_t mid = lo + (hi - lo) / 2;
            if ((long long)mulaw_level(key2f(mid), mu, C, denom) >= k) hi = mid; else lo = mid;
        }
        thr[k] = key2f(hi);
    }
}

// The proof: every float in [-1, 1] (by key) through the exact function and through the table quantiser.
__global__ void __launch_bounds__(256) mulaw_verify_kernel(const float* __restrict__ thr, MuLawFast f, float mu, float C,
                                                           unsigned long long* __restrict__ mismatches) {
    extern __shared__ float2 s_tab[];
    for (int k = threadIdx.x; k < f.Q; k += blockDim.x) s_tab[k] = make_float2(thr[k], thr[k + 1]);
    __syncthreads();
    const float denom = p_log1pf(__fmul_rn(mu, C));
    const uint32_t k0 = f2key(-1.0f), k1 = f2key(1.0f);
    const unsigned long long total = (unsigned long long)(k1 - k0) + 1ull;
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        float v = key2f(k0 + (uint32_t)i);
        long long e = (long long)mulaw_level(v, mu, C, denom);
        long long t = mulaw_table_level(v, f, s_tab);
        bad += (e != t);
    }
    bad = __reduce_add_sync(0xffffffffu, (unsigned)bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

// ---- table kernels ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st256(long long* p, long long a, long long b, long long c, long long d) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void ld256(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p) : "memory");
}

// WIDE: 0 = 16-byte accesses; 1 = one 32-byte store per 4 int64 results (STG.256, sm_100); 2 = 32-byte loads too
template <typename OutT, int WIDE>
__global__ void __launch_bounds__(256, 8) mulaw_compress_table_kernel(const float* __restrict__ x, OutT* __restrict__ q,
                                                                      size_t n, float mu, float C, MuLawFast f,
                                                                      const float* __restrict__ thr) {
    extern __shared__ float2 s_tab[];
    for (int k = threadIdx.x; k < f.Q; k += blockDim.x) s_tab[k] = make_float2(thr[k], thr[k + 1]);
    __syncthreads();
    auto one = [&](float v) -> long long {
        if (fabsf(v) <= 1.0f) return mulaw_table_level(v, f, s_tab);
        return mulaw_exact_slow(v, mu, C);   // out of range / NaN: the reference's arithmetic, whatever it yields
    };
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    auto store4 = [&](size_t i, const float4& v) {
        long long r0 = one(v.x), r1 = one(v.y), r2 = one(v.z), r3 = one(v.w);
        if constexpr (sizeof(OutT) == 8) {
            if constexpr (WIDE >= 1) {
                st256(reinterpret_cast<long long*>(q) + 4 * i, r0, r1, r2, r3);
            } else {
                longlong2* o = reinterpret_cast<longlong2*>(q) + 2 * i;
                __stcs(o, make_longlong2(r0, r1));
                __stcs(o + 1, make_longlong2(r2, r3));
            }
        } else {
            reinterpret_cast<uchar4*>(q)[i] =
                make_uchar4((unsigned char)r0, (unsigned char)r1, (unsigned char)r2, (unsigned char)r3);
        }
    };
    if constexpr (WIDE == 2) {
        const size_t n8 = n / 8;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
            float4 a, b;
            ld256(x + 8 * i, a, b);
            store4(2 * i, a);
            store4(2 * i + 1, b);
        }
        if (n4 > 2 * n8 && blockIdx.x == 0 && threadIdx.x == 0) store4(2 * n8, __ldcs(reinterpret_cast<const float4*>(x) + 2 * n8));
    } else {
        // (two separate 16-byte loads in flight per thread measured 7 % slower than one: 2.01 vs 1.88 ms for 10 h)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
            store4(i, __ldcs(reinterpret_cast<const float4*>(x) + i));
    }
    size_t t = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) q[t] = (OutT)one(x[t]);
}

__global__ void __launch_bounds__(256, 8) mulaw_expand_table_kernel(const long long* __restrict__ q, float* __restrict__ x,
                                                                    size_t n, float mu, float C, int Q,
                                                                    const float* __restrict__ expand_tab) {
    extern __shared__ float s_exp[];
    for (int k = threadIdx.x; k < Q; k += blockDim.x) s_exp[k] = expand_tab[k];
    __syncthreads();
    auto one = [&](long long idx) -> float {
        if ((unsigned long long)idx < (unsigned long long)Q) return s_exp[(int)idx];
        return mulaw_expand_slow(idx, mu, C);
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

// ---- Normalize(p = inf, dim = -1): F.normalize(x, p=inf, dim=-1) = x / max(max|x|, eps), eps = 1e-12 ----------------------
// (mimikit/features/functionals.py:236-253).  Pass 1: row maxima of |x| (NaN-propagating: the bit pattern of a NaN is
// above +inf's, so an unsigned max of the |x| bit patterns keeps it).  Pass 2: one IEEE division per sample — alone, or
// fused with the mu-law quantiser (Compose(Normalize(), MuLawCompress()) in 16 B per sample instead of 24).
constexpr int NORM_CHUNK4 = 256 * 8;   // float4 per CTA

__global__ void __launch_bounds__(256) rowmax_kernel(const float* __restrict__ x, long long row_len, long long row_stride,
                                                     unsigned* __restrict__ rowmax_bits) {
    const float* xr = x + (size_t)blockIdx.y * row_stride;
    const bool vec = ((reinterpret_cast<uintptr_t>(xr) & 15u) == 0);
    unsigned m = 0u;
    const long long c0 = (long long)blockIdx.x * NORM_CHUNK4 * 4, c1 = min(row_len, c0 + (long long)NORM_CHUNK4 * 4);
    if (vec) {
        const long long n4 = (c1 - c0) / 4;
        const float4* x4 = reinterpret_cast<const float4*>(xr + c0);
        for (long long i = threadIdx.x; i < n4; i += 256) {
            const float4 v = __ldg(x4 + i);
            m = max(max(m, __float_as_uint(fabsf(v.x))), __float_as_uint(fabsf(v.y)));
            m = max(max(m, __float_as_uint(fabsf(v.z))), __float_as_uint(fabsf(v.w)));
        }
        for (long long i = c0 + n4 * 4 + threadIdx.x; i < c1; i += 256) m = max(m, __float_as_uint(fabsf(xr[i])));
    } else {
        for (long long i = c0 + threadIdx.x; i < c1; i += 256) m = max(m, __float_as_uint(fabsf(xr[i])));
    }
    m = __reduce_max_sync(0xffffffffu, m);
    __shared__ unsigned s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = __reduce_max_sync(0xffu, s_m[threadIdx.x]);
        if (threadIdx.x == 0) atomicMax(rowmax_bits + blockIdx.y, m);
    }
}

__device__ __forceinline__ float norm_denominator(unsigned bits) {
    const float n = __uint_as_float(bits);
    return (n != n) ? n : fmaxf(n, 1e-12f);     // clamp_min(norm, eps); a NaN norm stays NaN
}

// MODE 0: fp32 out = x / d;  MODE 1: int64 mu-law level of x / d through the proven table (thr != nullptr) or the exact path
template <int MODE>
__global__ void __launch_bounds__(256) normalize_kernel(const float* __restrict__ x, void* __restrict__ out, long long row_len,
                                                        long long row_stride, const unsigned* __restrict__ rowmax_bits,
                                                        float mu, float C, MuLawFast f, const float* __restrict__ thr) {
    extern __shared__ float2 s_tab[];
    if (MODE == 1 && thr) {
        for (int k = threadIdx.x; k < f.Q; k += blockDim.x) s_tab[k] = make_float2(thr[k], thr[k + 1]);
        __syncthreads();
    }
    const float d = norm_denominator(rowmax_bits[blockIdx.y]);
    const float denom = MODE == 1 ? p_log1pf(__fmul_rn(mu, C)) : 0.0f;
    auto level = [&](float v) -> long long {
        if (thr && fabsf(v) <= 1.0f) return mulaw_table_level(v, f, s_tab);
        return (long long)mulaw_level(v, mu, C, denom);
    };
    const float* xr = x + (size_t)blockIdx.y * row_stride;
    float* of = reinterpret_cast<float*>(out) + (size_t)blockIdx.y * row_len;
    long long* oq = reinterpret_cast<long long*>(out) + (size_t)blockIdx.y * row_len;
    const long long c0 = (long long)blockIdx.x * NORM_CHUNK4 * 4, c1 = min(row_len, c0 + (long long)NORM_CHUNK4 * 4);
    const bool vec = ((reinterpret_cast<uintptr_t>(xr) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(MODE == 0 ? (void*)of : (void*)oq) & (MODE == 0 ? 15u : 31u)) == 0);
    long long done = c0;
    if (vec) {
        const long long n4 = (c1 - c0) / 4;
        const float4* x4 = reinterpret_cast<const float4*>(xr + c0);
        for (long long i = threadIdx.x; i < n4; i += 256) {
            const float4 v = __ldcs(x4 + i);
            const float a = __fdiv_rn(v.x, d), b = __fdiv_rn(v.y, d), c = __fdiv_rn(v.z, d), e = __fdiv_rn(v.w, d);
            if (MODE == 0) {
                __stcs(reinterpret_cast<float4*>(of + c0) + i, make_float4(a, b, c, e));
            } else {
                st256(oq + c0 + 4 * i, level(a), level(b), level(c), level(e));      // one 32-byte store (STG.256)
            }
        }
        done = c0 + n4 * 4;
    }
    for (long long i = done + threadIdx.x; i < c1; i += 256) {
        const float a = __fdiv_rn(xr[i], d);
        if (MODE == 0) of[i] = a; else oq[i] = level(a);
    }
}

static int feature_grid(size_t work_items) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t want = (work_items + 255) / 256;
    size_t cap = (size_t)sms * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// ---- per-(device, q_levels, compression) table cache ---------------------------------------------------------------
struct MuLawTable {
    float* thr = nullptr;      // Q + 1 thresholds
    float* expand = nullptr;   // Q exact expansion values
    MuLawFast fast{};
    int state = 0;             // 1 = proven, use the table kernels; -1 = not usable, exact kernels
    unsigned long long mismatches = 0;
};
static std::mutex g_mulaw_mu;
static std::map<std::tuple<int, int, uint32_t>, MuLawTable> g_mulaw_tables;

static bool mulaw_force_exact() {
    const char* e = getenv("MMK_MULAW_EXACT");
    return e && atoi(e) != 0;
}

// Returns the cache entry (building and proving it on first use: two small launches + one ~20 ms proof launch and a
// stream synchronisation, once per process and parameter pair).  *out = nullptr when the exact kernels must be used.
static int mulaw_table(int q_levels, float compression, cudaStream_t st, const MuLawTable** out) {
    *out = nullptr;
    if (mulaw_force_exact() || q_levels > MULAW_TABLE_MAX_Q || !(compression > 0.0f)) return 0;
    int dev = 0;
    MMK_CUDA(cudaGetDevice(&dev));
    uint32_t cbits;
    memcpy(&cbits, &compression, 4);
    std::lock_guard<std::mutex> lock(g_mulaw_mu);
    MuLawTable& t = g_mulaw_tables[std::make_tuple(dev, q_levels, cbits)];
    if (t.state == 0) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap != cudaStreamCaptureStatusNone) return 0;   // not provable inside a capture: exact kernel this time
        const int Q = q_levels;
        const float mu = (float)Q - 1.0f;
        unsigned* d_status = nullptr;
        unsigned long long* d_bad = nullptr;
        MMK_CUDA(cudaMalloc(&t.thr, (Q + 1) * sizeof(float)));
        MMK_CUDA(cudaMalloc(&t.expand, Q * sizeof(float)));
        MMK_CUDA(cudaMalloc(&d_status, 16));
        d_bad = reinterpret_cast<unsigned long long*>(d_status) + 1;
        MMK_CUDA(cudaMemsetAsync(d_status, 0, 16, st));
        mulaw_build_kernel<<<(Q + 1 + 127) / 128, 128, 0, st>>>(t.thr, t.expand, Q, mu, compression, d_status);
        double muC = (double)mu * (double)compression;
        t.fast.muC = (float)muC;
        t.fast.scale = (float)(0.5 * (double)mu / log2(1.0 + muC));
        t.fast.bias = (float)(0.5 * (double)mu + 0.5);
        t.fast.Q = Q;
        mulaw_verify_kernel<<<feature_grid((size_t)1 << 26), 256, Q * sizeof(float2), st>>>(t.thr, t.fast, mu, compression, d_bad);
        MMK_CUDA(cudaGetLastError());
        unsigned h[4] = {0, 0, 0, 0};
        MMK_CUDA(cudaMemcpyAsync(h, d_status, 16, cudaMemcpyDeviceToHost, st));
        MMK_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_status);
        memcpy(&t.mismatches, &h[2], 8);
        t.state = (h[0] == 0 && t.mismatches == 0) ? 1 : -1;
    }
    if (t.state == 1) *out = &t;
    return 0;
}

}  // namespace mmk

using namespace mmk;

extern "C" int mmk_mulaw_prepare(int q_levels, float compression, int* table_in_use, uint64_t* mismatches, void* stream) {
    MMK_CHECK(q_levels >= 2, "mmk_mulaw_prepare: q_levels must be >= 2");
    const MuLawTable* t = nullptr;
    if (int rc = mulaw_table(q_levels, compression, (cudaStream_t)stream, &t)) return rc;
    if (table_in_use) *table_in_use = t != nullptr;
    if (mismatches) {
        *mismatches = 0;
        if (!t && !mulaw_force_exact() && q_levels <= MULAW_TABLE_MAX_Q) {
            int dev = 0;
            cudaGetDevice(&dev);
            uint32_t cbits;
            memcpy(&cbits, &compression, 4);
            std::lock_guard<std::mutex> lock(g_mulaw_mu);
            auto it = g_mulaw_tables.find(std::make_tuple(dev, q_levels, cbits));
            if (it != g_mulaw_tables.end()) *mismatches = it->second.mismatches;
        }
    }
    return 0;
}

extern "C" int mmk_mulaw_compress(const float* d_x, int64_t* d_q, size_t n, int q_levels, float compression,
                                  void* stream) {
    MMK_CHECK(q_levels >= 2, "mmk_mulaw_compress: q_levels must be >= 2");
    if (n == 0) return 0;
    MMK_CHECK(d_x && d_q, "mmk_mulaw_compress: null pointer");
    MMK_CHECK(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_q % 16) == 0, "mmk_mulaw_compress: buffers must be 16-byte aligned");
    const MuLawTable* t = nullptr;
    if (int rc = mulaw_table(q_levels, compression, (cudaStream_t)stream, &t)) return rc;
    const float mu = (float)q_levels - 1.0f;
    if (t) {
        int wide = (((uintptr_t)d_q % 32) == 0) ? MULAW_WIDE_DEFAULT : 0;
        if (wide == 2 && ((uintptr_t)d_x % 32) != 0) wide = 1;
        const int grid = feature_grid(n / 4 + 1);
        const size_t sm = q_levels * sizeof(float2);
        long long* q = reinterpret_cast<long long*>(d_q);
        if (wide == 2) mulaw_compress_table_kernel<long long, 2><<<grid, 256, sm, (cudaStream_t)stream>>>(d_x, q, n, mu, compression, t->fast, t->thr);
        else if (wide == 1) mulaw_compress_table_kernel<long long, 1><<<grid, 256, sm, (cudaStream_t)stream>>>(d_x, q, n, mu, compression, t->fast, t->thr);
        else mulaw_compress_table_kernel<long long, 0><<<grid, 256, sm, (cudaStream_t)stream>>>(d_x, q, n, mu, compression, t->fast, t->thr);
    } else
        mulaw_compress_kernel<long long><<<feature_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(
            d_x, reinterpret_cast<long long*>(d_q), n, mu, compression);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmk_mulaw_compress_u8(const float* d_x, uint8_t* d_q, size_t n, int q_levels, float compression,
                                     void* stream) {
    MMK_CHECK(q_levels >= 2 && q_levels <= 256, "mmk_mulaw_compress_u8: q_levels must be in [2, 256]");
    if (n == 0) return 0;
    MMK_CHECK(d_x && d_q, "mmk_mulaw_compress_u8: null pointer");
    MMK_CHECK(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_q % 4) == 0, "mmk_mulaw_compress_u8: misaligned buffers");
    const MuLawTable* t = nullptr;
    if (int rc = mulaw_table(q_levels, compression, (cudaStream_t)stream, &t)) return rc;
    const float mu = (float)q_levels - 1.0f;
    if (t)
        mulaw_compress_table_kernel<unsigned char, 0><<<feature_grid(n / 4 + 1), 256, q_levels * sizeof(float2), (cudaStream_t)stream>>>(
            d_x, d_q, n, mu, compression, t->fast, t->thr);
    else
        mulaw_compress_kernel<unsigned char><<<feature_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(
            d_x, d_q, n, mu, compression);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmk_mulaw_expand(const int64_t* d_q, float* d_x, size_t n, int q_levels, float compression,
                                void* stream) {
    MMK_CHECK(q_levels >= 2, "mmk_mulaw_expand: q_levels must be >= 2");
    if (n == 0) return 0;
    MMK_CHECK(d_x && d_q, "mmk_mulaw_expand: null pointer");
    MMK_CHECK(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_q % 16) == 0, "mmk_mulaw_expand: buffers must be 16-byte aligned");
    const MuLawTable* t = nullptr;
    if (int rc = mulaw_table(q_levels, compression, (cudaStream_t)stream, &t)) return rc;
    const float mu = (float)q_levels - 1.0f;
    if (t)
        mulaw_expand_table_kernel<<<feature_grid(n / 4 + 1), 256, q_levels * sizeof(float), (cudaStream_t)stream>>>(
            reinterpret_cast<const long long*>(d_q), d_x, n, mu, compression, q_levels, t->expand);
    else
        mulaw_expand_kernel<<<feature_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const long long*>(d_q), d_x, n, mu, compression);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

static int normalize_common(const float* d_x, void* d_out, float* d_norms, int64_t n_rows, int64_t row_len, int64_t row_stride,
                            int mode, int q_levels, float compression, cudaStream_t st) {
    MMK_CHECK(n_rows >= 0 && row_len >= 0 && row_stride >= row_len, "normalize: bad geometry");
    if (n_rows == 0 || row_len == 0) return 0;
    MMK_CHECK(d_x && d_out && d_norms, "normalize: null pointer");
    MMK_CHECK(n_rows <= 65535, "normalize: at most 65535 rows per call");
    const unsigned chunks = (unsigned)((row_len + (long long)NORM_CHUNK4 * 4 - 1) / ((long long)NORM_CHUNK4 * 4));
    const dim3 grid(chunks, (unsigned)n_rows);
    MMK_CUDA(cudaMemsetAsync(d_norms, 0, (size_t)n_rows * sizeof(float), st));
    rowmax_kernel<<<grid, 256, 0, st>>>(d_x, row_len, row_stride, reinterpret_cast<unsigned*>(d_norms));
    MMK_CUDA(cudaGetLastError());
    if (mode == 0) {
        normalize_kernel<0><<<grid, 256, 0, st>>>(d_x, d_out, row_len, row_stride, reinterpret_cast<const unsigned*>(d_norms),
                                                  0.0f, 0.0f, MuLawFast{}, nullptr);
    } else {
        const MuLawTable* t = nullptr;
        if (int rc = mulaw_table(q_levels, compression, st, &t)) return rc;
        normalize_kernel<1><<<grid, 256, t ? q_levels * sizeof(float2) : 0, st>>>(
            d_x, d_out, row_len, row_stride, reinterpret_cast<const unsigned*>(d_norms), (float)q_levels - 1.0f, compression,
            t ? t->fast : MuLawFast{}, t ? t->thr : nullptr);
    }
    MMK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmk_normalize_inf(const float* d_x, float* d_out, float* d_norms, int64_t n_rows, int64_t row_len,
                                 int64_t row_stride, void* stream) {
    return normalize_common(d_x, d_out, d_norms, n_rows, row_len, row_stride, 0, 0, 0.0f, (cudaStream_t)stream);
}

extern "C" int mmk_normalize_mulaw_compress(const float* d_x, int64_t* d_q, float* d_norms, int64_t n_rows, int64_t row_len,
                                            int64_t row_stride, int q_levels, float compression, void* stream) {
    MMK_CHECK(q_levels >= 2, "mmk_normalize_mulaw_compress: q_levels must be >= 2");
    return normalize_common(d_x, d_q, d_norms, n_rows, row_len, row_stride, 1, q_levels, compression, (cudaStream_t)stream);
}
