// Shared host/device helpers for the mmk_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace mmk {

void set_error(const std::string& msg);
int fail(const char* file, int line, const std::string& msg);

#define MMK_FAIL(msg) return ::mmk::fail(__FILE__, __LINE__, (msg))
#define MMK_CHECK(cond, msg) do { if (!(cond)) MMK_FAIL(msg); } while (0)
#define MMK_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) \
    MMK_FAIL(std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

// ---------------------------------------------------------------------------------------------
// Portable fp32 math: every operation is a correctly-rounded IEEE op written with explicit
// intrinsics (no FMA contraction, no fast-math), so that the same sequence evaluated by the C
// oracle (oracle/c/oracle_feat.c) gives the same bits.  log1pf/expf follow the u10 algorithms of
// the Sleef library that torch's CPU kernels call (see the oracle for provenance and pinning).
// ---------------------------------------------------------------------------------------------
struct f2 { float x, y; };

__device__ __forceinline__ float fmapn(float a, float b, float c) { return __fmaf_rn(a, b, -c); }
__device__ __forceinline__ float fmanp(float a, float b, float c) { return __fmaf_rn(-a, b, c); }

__device__ __forceinline__ float p_log1pf(float d) {
    float dp1 = __fadd_rn(d, 1.0f);
    float q = __fmul_rn(dp1, 1.0f / 0.75f);
    int e = (int)((__float_as_uint(q) >> 23) & 0xff) - 127;
    float t = __uint_as_float((uint32_t)(127 - e) << 23);
    float m = __fmaf_rn(d, t, __fsub_rn(t, 1.0f));
    // (ln2_hi, ln2_lo) * e
    const float lx = 0.69314718246459960938f, ly = -1.904654323148236017e-09f;
    float ef = (float)e;
    f2 s; s.x = __fmul_rn(lx, ef); s.y = __fmaf_rn(ly, ef, fmapn(lx, ef, s.x));
    // x = (m, 0) / (2 + m)
    f2 dn; dn.x = __fadd_rn(2.0f, m); dn.y = __fadd_rn(__fsub_rn(2.0f, dn.x), m);
    float rt = __frcp_rn(dn.x);  // 1.0f / d.x, correctly rounded
    f2 x; x.x = __fmul_rn(m, rt);
    float u = fmapn(rt, m, x.x);
    float v = fmanp(dn.y, rt, fmanp(dn.x, rt, 1.0f));
    x.y = __fmaf_rn(x.x, v, __fmaf_rn(0.0f, rt, u));
    float x2 = __fmul_rn(x.x, x.x);
    float p = +0.3027294874e+0f;
    p = __fmaf_rn(p, x2, +0.3996108174e+0f);
    p = __fmaf_rn(p, x2, +0.6666694880e+0f);
    // s += 2x
    f2 xs; xs.x = __fmul_rn(x.x, 2.0f); xs.y = __fmul_rn(x.y, 2.0f);
    f2 r; r.x = __fadd_rn(s.x, xs.x);
    r.y = __fadd_rn(__fadd_rn(__fadd_rn(__fsub_rn(s.x, r.x), xs.x), s.y), xs.y);
    // s += x2 * x * p
    float y = __fmul_rn(__fmul_rn(x2, x.x), p);
    f2 o; o.x = __fadd_rn(r.x, y);
    o.y = __fadd_rn(__fadd_rn(__fsub_rn(r.x, o.x), y), r.y);
    float res = __fadd_rn(o.x, o.y);
    if (d == 0.0f) res = d;
    return res;
}

__device__ __forceinline__ float p_expf(float d) {
    float qf = rintf(__fmul_rn(d, 1.442695040888963407359924681001892137426645954152985934135449406931f));
    int q = (int)qf;
    float s = __fmaf_rn(qf, -0.693145751953125f, d);
    s = __fmaf_rn(qf, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = __fmaf_rn(u, s, 0.00139304355252534151077271f);
    u = __fmaf_rn(u, s, 0.00833336077630519866943359f);
    u = __fmaf_rn(u, s, 0.0416664853692054748535156f);
    u = __fmaf_rn(u, s, 0.166666671633720397949219f);
    u = __fmaf_rn(u, s, 0.5f);
    u = __fadd_rn(1.0f, __fmaf_rn(__fmul_rn(s, s), u, s));
    int q1 = q >> 1, q2 = q - q1;
    u = __fmul_rn(__fmul_rn(u, __uint_as_float((uint32_t)(q1 + 127) << 23)), __uint_as_float((uint32_t)(q2 + 127) << 23));
    if (d < -104.0f) u = 0.0f;
    if (d > 100.0f) u = __int_as_float(0x7f800000);
    return u;
}

__device__ __forceinline__ float p_sign(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

}  // namespace mmk
