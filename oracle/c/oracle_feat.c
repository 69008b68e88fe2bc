/* TEST INFRASTRUCTURE ONLY (oracle) — plain-C restatement of the reference's mu-law arithmetic.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product path (mimikit_b200/) never does.
 *
 * Reference: mimikit/features/functionals.py:330-338 (MuLawCompress.torch_func) and :361-369
 * (MuLawExpand.torch_func), evaluated by torch on CPU in fp32.  torch's CPU `log1p`/`exp` for fp32 are
 * NOT correctly rounded: they dispatch to the Sleef vector math library bundled with torch
 * (Sleef_log1pf*_u10 / Sleef_expf*_u10; third-party dependency, not vendored in /root/reference; torch
 * pins it as a submodule, sleef 3.6.x).  Bit-exact mu-law indices therefore require Sleef's published
 * algorithm, restated here with explicit fmaf() (the AVX2/AVX-512 builds use hardware FMA).  Pinned:
 * tests/test_oracle_golden.py checks it against fixtures generated from the live reference, and
 * oracle/validate_mulaw.py checks log1pf against torch.log1p on every fp32 value in [0, 255] (0
 * mismatches in this container, torch 2.11.0 CPU, AVX-512).
 *
 * Compile with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y; } f2;

static inline float fmapn(float a, float b, float c) { return fmaf(a, b, -c); }  /* a*b - c */
static inline float fmanp(float a, float b, float c) { return fmaf(-a, b, c); }  /* -a*b + c */
static inline f2 dfadd_f_f(float x, float y) { float s = x + y; f2 r = {s, (x - s) + y}; return r; }
static inline f2 dfmul_f2_f(f2 x, float y) {
    float s = x.x * y; f2 r = {s, fmaf(x.y, y, fmapn(x.x, y, s))}; return r;
}
static inline f2 dfadd_f2_f2(f2 x, f2 y) {
    float s = x.x + y.x; f2 r = {s, (((x.x - s) + y.x) + x.y) + y.y}; return r;
}
static inline f2 dfadd_f2_f(f2 x, float y) {
    float s = x.x + y; f2 r = {s, ((x.x - s) + y) + x.y}; return r;
}
static inline f2 dfdiv(f2 n, f2 d) {
    float t = 1.0f / d.x;
    float s = n.x * t;
    float u = fmapn(t, n.x, s);
    float v = fmanp(d.y, t, fmanp(d.x, t, 1.0f));
    f2 r = {s, fmaf(s, v, fmaf(n.y, t, u))};
    return r;
}

/* Sleef xlog1pf (u10), valid for d > -1 and d+1 >= FLT_MIN (the mu-law domain is d in [0, mu*C]). */
float orc_log1pf(float d) {
    float dp1 = d + 1.0f;
    float q = dp1 * (1.0f / 0.75f);
    uint32_t b; memcpy(&b, &q, 4);
    int e = (int)((b >> 23) & 0xff) - 127;
    uint32_t tb = (uint32_t)(127 - e) << 23;
    float t; memcpy(&t, &tb, 4);
    float m = fmaf(d, t, t - 1.0f);
    f2 ln2 = {0.69314718246459960938f, -1.904654323148236017e-09f};
    f2 s = dfmul_f2_f(ln2, (float)e);
    f2 mm = {m, 0.0f};
    f2 x = dfdiv(mm, dfadd_f_f(2.0f, m));
    float x2 = x.x * x.x;
    float p = +0.3027294874e+0f;
    p = fmaf(p, x2, +0.3996108174e+0f);
    p = fmaf(p, x2, +0.6666694880e+0f);
    f2 xs = {x.x * 2.0f, x.y * 2.0f};
    s = dfadd_f2_f2(s, xs);
    s = dfadd_f2_f(s, x2 * x.x * p);
    float r = s.x + s.y;
    if (d == 0.0f) r = d;  /* keeps -0.0 */
    return r;
}

/* Sleef xexpf (u10). */
float orc_expf(float d) {
    float qf = rintf(d * 1.442695040888963407359924681001892137426645954152985934135449406931f);
    int q = (int)qf;
    float s = fmaf(qf, -0.693145751953125f, d);
    s = fmaf(qf, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = fmaf(u, s, 0.00139304355252534151077271f);
    u = fmaf(u, s, 0.00833336077630519866943359f);
    u = fmaf(u, s, 0.0416664853692054748535156f);
    u = fmaf(u, s, 0.166666671633720397949219f);
    u = fmaf(u, s, 0.5f);
    u = 1.0f + fmaf(s * s, u, s);
    /* vldexp2: u * 2^(q>>1) * 2^(q - (q>>1)) */
    int q1 = q >> 1, q2 = q - q1;
    uint32_t b1 = (uint32_t)(q1 + 127) << 23, b2 = (uint32_t)(q2 + 127) << 23;
    float p1, p2; memcpy(&p1, &b1, 4); memcpy(&p2, &b2, 4);
    u = u * p1 * p2;
    if (d < -104.0f) u = 0.0f;
    if (d > 100.0f) u = INFINITY;
    return u;
}

static inline float signf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

/* functionals.py:330-338, op for op, left to right. */
void orc_mulaw_compress(const float* x, int64_t* out, int64_t n, int q_levels, float compression) {
    const float mu = (float)q_levels - 1.0f;
    const float C = compression;
    const float denom = orc_log1pf(mu * C);
    for (int64_t i = 0; i < n; ++i) {
        float v = x[i];
        float a = mu * fabsf(v);
        a = a * C;
        float l = orc_log1pf(a);
        float xm = signf(v) * l;
        xm = xm / denom;
        float r = (xm + 1.0f) / 2.0f;
        r = r * mu;
        r = r + 0.5f;
        out[i] = (int64_t)r;
    }
}

/* functionals.py:361-369. */
void orc_mulaw_expand(const int64_t* idx, float* out, int64_t n, int q_levels, float compression) {
    const float mu = (float)q_levels - 1.0f;
    const float C = compression;
    const float l1p = orc_log1pf(mu * C);
    const float muC = mu * C;
    for (int64_t i = 0; i < n; ++i) {
        float v = (float)idx[i];
        float x = (v / mu) * 2.0f - 1.0f;
        float e = orc_expf(fabsf(x) * l1p);
        float y = signf(x) * (e - 1.0f);
        out[i] = y / muC;
    }
}

void orc_log1pf_arr(const float* in, float* out, int64_t n) { for (int64_t i = 0; i < n; ++i) out[i] = orc_log1pf(in[i]); }
void orc_expf_arr(const float* in, float* out, int64_t n) { for (int64_t i = 0; i < n; ++i) out[i] = orc_expf(in[i]); }

/* RemoveDC.np_func — mimikit/features/functionals.py:216-233: scipy.signal.lfilter([1, -1], [1, -0.99], x, axis=-1) on a
 * float32 signal: scipy promotes to float64 and runs its direct-form-II-transposed loop (scipy/signal/_lfilter.c.in,
 * @NAME@_filt: y = Z[0] + b[0] x ; Z[last] = x b[last] - y a[last]), zero initial state; the caller casts back to float32. */
void orc_remove_dc(const float* x, float* y, int64_t n_rows, int64_t row_len) {
    const double b1 = -1.0, a1 = -0.99;
    for (int64_t r = 0; r < n_rows; ++r) {
        double z = 0.0;
        for (int64_t n = 0; n < row_len; ++n) {
            const double xn = (double)x[r * row_len + n];
            const double yn = z + 1.0 * xn;
            z = xn * b1 - yn * a1;
            y[r * row_len + n] = (float)yn;
        }
    }
}
