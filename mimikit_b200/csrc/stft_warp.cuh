// Warp-per-frame STFT(2048) -> |.| -> mel kernel (sm_100a): the fast path of mmk_stft_mag_mel for n_fft = 2048, the
// size BASELINE.json quotes.  Same arithmetic contract as the general kernel in stft_mel.cu (reference:
// mimikit/features/functionals.py:468-524, 649-668): periodic hann window, centre zero padding, real FFT, magnitude,
// dense filterbank @ magnitudes.
//
// A real FFT of 2048 points is a complex FFT of 1024 points on z[n] = x[2n] + i x[2n+1] plus a split post-pass.
// 1024 = 32 x 32: ONE warp owns a frame and holds 32 complex points per lane in registers.
//   step A  lane n2 loads z[n2 + 32 n1] (coalesced 256-B rows, window applied on load) and runs a 32-point FFT over n1
//           entirely in registers (radix-2 DIF, compile-time twiddles, trivial twiddles folded);
//   twiddle Y[k1] *= exp(-2 pi i n2 k1 / 1024) from a conflict-free [k1][n2] table;
//   one exchange through a padded per-warp shared-memory tile (__syncwarp only: no CTA barrier in the frame loop);
//   step B  lane k1 runs the second 32-point FFT over n2 -> Z[k1 + 32 k2];
//   post    Z goes back to shared memory in natural order; each lane takes 16 (k, 1024-k) pairs and produces both
//           magnitudes of a pair from one twiddle; magnitudes overwrite the tile (1025 floats);
//   mel     lane m owns filters m, m+32, ...: sequential fp32 dot over the filter's non-zero bin range.
// Per frame: 4 shared-memory passes instead of the general kernel's 12 + 8 CTA barriers, and no idle lanes.
#pragma once
#include "common.cuh"

namespace mmk {

constexpr int FW_WARPS = 16;
constexpr int FW_THREADS = FW_WARPS * 32;
constexpr int FW_MEL_CAP = 384;    // rows of the packed filter table (x 128 B) + one flag word
constexpr int FW_TILE = 32 * 33;   // float2 per warp tile (row stride 33: conflict-free column reads)

struct StftParams {
    const float* x;
    float* mag_out;
    const float* mel_fb;
    float* mel_out;
    long long clip_stride, start, kept_len, n_frames, total_frames;
    int n_fft, hop, pad, n_mels, log2_half;
    int mel_cap;           // warp kernel: rows of the lane-interleaved filter table that fit in shared memory (0 = none)
};

// [lo, hi) of the non-zero support of each dense filter row, computed by the whole CTA into shared memory at kernel
// start (one warp per filter; 525 KB of L2 reads per CTA for 128 x 1025 — no side allocation, no extra launch).
__device__ __forceinline__ void mel_ranges_to_smem(const float* __restrict__ fb, int n_mels, int nb, int* s_rng) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int m = warp; m < n_mels; m += n_warps) {
        int lo = nb, hi = 0;
        for (int k = lane; k < nb; k += 32)
            if (__ldg(fb + (long long)m * nb + k) != 0.0f) { lo = min(lo, k); hi = max(hi, k + 1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { s_rng[2 * m] = (lo < hi) ? lo : 0; s_rng[2 * m + 1] = (lo < hi) ? hi : 0; }
    }
}

__device__ __forceinline__ float2 w32(int j) {   // exp(-2 pi i j / 32)
    switch (j) {
        case 1: return make_float2(0.98078528040323043f, -0.19509032201612825f);
        case 2: return make_float2(0.92387953251128674f, -0.38268343236508978f);
        case 3: return make_float2(0.83146961230254524f, -0.55557023301960218f);
        case 5: return make_float2(0.55557023301960229f, -0.83146961230254524f);
        case 6: return make_float2(0.38268343236508984f, -0.92387953251128674f);
        case 7: return make_float2(0.19509032201612833f, -0.98078528040323043f);
        case 9: return make_float2(-0.19509032201612819f, -0.98078528040323043f);
        case 10: return make_float2(-0.38268343236508973f, -0.92387953251128674f);
        case 11: return make_float2(-0.55557023301960196f, -0.83146961230254546f);
        case 13: return make_float2(-0.83146961230254535f, -0.55557023301960218f);
        case 14: return make_float2(-0.92387953251128674f, -0.38268343236508989f);
        case 15: return make_float2(-0.98078528040323043f, -0.19509032201612861f);
        default: return make_float2(1.0f, 0.0f);
    }
}

// d * exp(-2 pi i idx / 32); idx is a compile-time constant after unrolling, so the chain folds
__device__ __forceinline__ float2 tw32_mul(float2 d, int idx) {
    const float r = 0.70710678118654752f;
    if (idx == 0) return d;
    if (idx == 8) return make_float2(d.y, -d.x);
    if (idx == 4) return make_float2((d.x + d.y) * r, (d.y - d.x) * r);
    if (idx == 12) return make_float2((d.y - d.x) * r, -(d.x + d.y) * r);
    const float2 w = w32(idx);
    return make_float2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x);
}

__host__ __device__ __forceinline__ constexpr int brev5(int r) {
    return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4);
}

// 32-point complex FFT in registers, radix-2 decimation in frequency: v[r] <- X[brev5(r)]
__device__ __forceinline__ void fft32_dif(float2 (&v)[32]) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
        for (int b = 0; b < 32; b += 2 * s) {
#pragma unroll
            for (int j = 0; j < s; ++j) {
                const float2 a = v[b + j], c = v[b + j + s];
                v[b + j] = make_float2(a.x + c.x, a.y + c.y);
                v[b + j + s] = tw32_mul(make_float2(a.x - c.x, a.y - c.y), j * (16 / s));
            }
        }
    }
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(FW_THREADS, 1) stft2048_warp_kernel(const StftParams p) {
    constexpr int N = 2048, H = 1024;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    float2* tw = reinterpret_cast<float2*>(fw_smem);   // [k1][n2] = exp(-2 pi i n2 k1 / 1024)
    float2* twN = tw + H;                              // [k] = exp(-2 pi i k / 2048), k < 512
    float* win = reinterpret_cast<float*>(twN + 512);  // [2048]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* tile = reinterpret_cast<float2*>(win + N) + (size_t)warp * FW_TILE;
    float* magb = reinterpret_cast<float*>(tile);      // magnitudes overwrite the tile once Z has been consumed
    // lane-interleaved filter table: lane L owns filters L, L+32, ...; its non-zero weights, filter after filter, sit in
    // column L of wsm[row][32] (conflict-free reads).  Built once per CTA from the dense filterbank.
    float* wsm = reinterpret_cast<float*>(reinterpret_cast<float2*>(win + N) + (size_t)FW_WARPS * FW_TILE);
    int* s_rng = reinterpret_cast<int*>(wsm + (size_t)FW_MEL_CAP * 32 + 4);   // [n_mels][2], after the flag word
    const int n_grp = (p.n_mels + 31) / 32;
    bool mel_packed = false;
    if (p.mel_out) {
        mel_ranges_to_smem(p.mel_fb, p.n_mels, H + 1, s_rng);
        __syncthreads();
    }
    if (p.mel_out && p.mel_cap > 0 && tid < 32) {
        int rows = 0;
        for (int g = 0; g < n_grp; ++g) {
            const int m = tid + 32 * g;
            if (m < p.n_mels) rows += s_rng[2 * m + 1] - s_rng[2 * m];
        }
        int mx = rows;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (mx <= p.mel_cap) {
            int row = 0;
            for (int g = 0; g < n_grp; ++g) {
                const int m = tid + 32 * g;
                if (m >= p.n_mels) break;
                const int lo = s_rng[2 * m], hi = s_rng[2 * m + 1];
                for (int k = lo; k < hi; ++k) wsm[(row++) * 32 + tid] = __ldg(p.mel_fb + (long long)m * (H + 1) + k);
            }
        }
        if (tid == 0) *reinterpret_cast<int*>(wsm + (size_t)p.mel_cap * 32) = mx <= p.mel_cap ? 1 : 0;
    }

    for (int i = tid; i < H; i += FW_THREADS) {
        const int k1 = i >> 5, n2 = i & 31;
        float s, c;
        sincospif(-2.0f * (float)((k1 * n2) & (H - 1)) / (float)H, &s, &c);
        tw[i] = make_float2(c, s);
    }
    for (int k = tid; k < 512; k += FW_THREADS) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)N, &s, &c);
        twN[k] = make_float2(c, s);
    }
    for (int n = tid; n < N; n += FW_THREADS) win[n] = 0.5f - 0.5f * cospif(2.0f * (float)n / (float)N);
    __syncthreads();
    if (p.mel_out && p.mel_cap > 0) mel_packed = *reinterpret_cast<const int*>(wsm + (size_t)p.mel_cap * 32) != 0;

    const int nb = H + 1;
    // warps of a CTA take adjacent frames (their 75 % input overlap is served by L1)
    for (long long f0 = (long long)blockIdx.x * FW_WARPS; f0 < p.total_frames; f0 += (long long)gridDim.x * FW_WARPS) {
        const long long f = f0 + warp;
        if (f >= p.total_frames) break;
        const long long clip = f / p.n_frames, jf = f - clip * p.n_frames;
        const float* xc = p.x + clip * p.clip_stride + p.start;
        const long long base = jf * p.hop - p.pad;
        float2 v[32];
        // ---- windowed load: lane n2, register n1 <- z[n2 + 32 n1]
        if (base >= 0 && base + N <= p.kept_len) {
            const float* src = xc + base + 2 * lane;
            if ((reinterpret_cast<uintptr_t>(src) & 7u) == 0) {
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) v[n1] = __ldg(reinterpret_cast<const float2*>(src + 64 * n1));
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) v[n1] = make_float2(__ldg(src + 64 * n1), __ldg(src + 64 * n1 + 1));
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const long long p0 = base + 2 * lane + 64 * n1, p1 = p0 + 1;
                v[n1].x = (p0 >= 0 && p0 < p.kept_len) ? __ldg(xc + p0) : 0.0f;
                v[n1].y = (p1 >= 0 && p1 < p.kept_len) ? __ldg(xc + p1) : 0.0f;
            }
        }
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const float2 w = *reinterpret_cast<const float2*>(win + 2 * lane + 64 * n1);
            v[n1].x *= w.x; v[n1].y *= w.y;
        }
        // ---- step A + twiddle + exchange
        fft32_dif(v);
        __syncwarp();   // the previous frame's mel reads of the tile are done
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int k1 = brev5(r);
            float2 y = v[r];
            if (k1 != 0) {
                const float2 w = tw[k1 * 32 + lane];
                y = make_float2(y.x * w.x - y.y * w.y, y.x * w.y + y.y * w.x);
            }
            tile[k1 * 33 + lane] = y;
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2] = tile[lane * 33 + n2];
        __syncwarp();
        // ---- step B: v[r] <- Z[lane + 32 brev5(r)], then natural order in shared memory
        fft32_dif(v);
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[lane + 32 * brev5(r)] = v[r];
        __syncwarp();
        // ---- split post-pass on pairs (k, 1024 - k), k = lane + 32 j < 512:
        //   2 X[k] = e - i w o,  2 X[1024-k] = conj(e + i w o),  e = Z[k] + conj Z[1024-k], o = Z[k] - conj Z[1024-k]
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = lane + 32 * j;
            v[j] = tile[k];
            v[16 + j] = tile[(H - k) & (H - 1)];
        }
        const float2 zmid = tile[512];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = lane + 32 * j;
            const float2 a = v[j], b = v[16 + j], w = twN[k];
            const float2 e = make_float2(a.x + b.x, a.y - b.y), o = make_float2(a.x - b.x, a.y + b.y);
            const float2 wo = make_float2(w.x * o.x - w.y * o.y, w.x * o.y + w.y * o.x);
            const float re0 = e.x + wo.y, im0 = e.y - wo.x, re1 = e.x - wo.y, im1 = e.y + wo.x;
            magb[k] = 0.5f * sqrt_approx(re0 * re0 + im0 * im0);
            magb[H - k] = 0.5f * sqrt_approx(re1 * re1 + im1 * im1);
        }
        if (lane == 0) magb[512] = sqrt_approx(zmid.x * zmid.x + zmid.y * zmid.y);
        __syncwarp();
        if (p.mag_out) {
            float* mrow = p.mag_out + f * (long long)nb;
#pragma unroll 4
            for (int k = lane; k < nb; k += 32) __stcs(mrow + k, magb[k]);
        }
        if (p.mel_out) {
            float* orow = p.mel_out + f * (long long)p.n_mels;
            if (mel_packed) {
                const float* wl = wsm + lane;
                for (int m = lane; m < p.n_mels; m += 32) {
                    const int2 rg = *(reinterpret_cast<const int2*>(s_rng) + m);
                    float acc = 0.0f;
#pragma unroll 4
                    for (int k = rg.x; k < rg.y; ++k, wl += 32) acc = fmaf(*wl, magb[k], acc);
                    __stcs(orow + m, acc);
                }
            } else {
                for (int m = lane; m < p.n_mels; m += 32) {
                    const int lo = s_rng[2 * m], hi = s_rng[2 * m + 1];
                    const float* fb = p.mel_fb + (long long)m * nb;
                    float acc = 0.0f;
                    for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(fb + k), magb[k], acc);
                    __stcs(orow + m, acc);
                }
            }
        }
    }
}

constexpr size_t FW_SMEM_BASE = sizeof(float2) * (1024 + 512) + sizeof(float) * 2048 + sizeof(float2) * FW_TILE * FW_WARPS;
constexpr size_t FW_SMEM_BYTES = FW_SMEM_BASE + (size_t)FW_MEL_CAP * 128 + 16;   // + 8 bytes per mel filter (ranges)
constexpr int FW_MAX_MELS = 2048;

}  // namespace mmk
