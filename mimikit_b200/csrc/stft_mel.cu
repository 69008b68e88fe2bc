// Framed STFT magnitude with the mel projection fused into the epilogue (sm_100a).
//
// Reference arithmetic: MagSpec.torch_func -> STFT(coordinate="mag").torch_func, mimikit/features/functionals.py
// :468-524 (length fix "end"/"start", centre zero padding, periodic hann — :513 always uses torch.hann_window —
// torch.stft, transpose to (frames, bins), abs) and MelSpec.np_func :665-668 (librosa mel basis @ magnitudes).
//
// One CTA processes frames in a grid-stride loop.  A real FFT of n_fft points is done as a complex FFT of
// n_fft/2 points (even samples -> re, odd -> im) with a radix-4 Stockham autosort in shared memory (ping-pong
// buffers, one butterfly per thread and stage, a radix-2 stage when log2 is odd), then the split post-pass that
// yields bins 0..n_fft/2, |.|, a coalesced store of the magnitudes, and — magnitudes still in shared memory — the
// sparse triangular mel reduce (each filter only touches its [lo, hi) bin range).  Twiddles and window are built
// once per CTA with sincospif (exact argument reduction) and reused for every frame the CTA handles.
//
// Roofline: HBM.  Algorithmic bytes per frame: hop*4 in (each sample is read from DRAM once; the 4x frame
// overlap is served by L2) + n_mels*4 out (+ (n_fft/2+1)*4 when the magnitudes are materialised).
#include "common.cuh"
#include "stft_warp.cuh"
#include "../../include/mmk_b200.h"

#include <math.h>
#include <algorithm>
#include <vector>

namespace mmk {

constexpr int STFT_THREADS = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(STFT_THREADS) stft_mag_mel_kernel(StftParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.n_fft, H = N / 2, nb = H + 1;
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);  // H complex
    float2* buf1 = buf0 + H;                             // H complex
    float2* twH = buf1 + H;                              // exp(-2 pi i k / H), k < H
    float2* twN = twH + H;                               // exp(-2 pi i k / N), k < H
    float* win = reinterpret_cast<float*>(twN + H);      // N
    float* mag = win + N;                                // nb (+pad)
    int* s_rng = reinterpret_cast<int*>(mag + H + 8);    // [n_mels][2] non-zero bin range of each filter
    const int tid = threadIdx.x;
    if (p.mel_out) mel_ranges_to_smem(p.mel_fb, p.n_mels, nb, s_rng);

    for (int k = tid; k < H; k += STFT_THREADS) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)H, &s, &c);
        twH[k] = make_float2(c, s);
        sincospif(-2.0f * (float)k / (float)N, &s, &c);
        twN[k] = make_float2(c, s);
    }
    for (int n = tid; n < N; n += STFT_THREADS) win[n] = 0.5f - 0.5f * cospif(2.0f * (float)n / (float)N);
    __syncthreads();

    for (long long f = blockIdx.x; f < p.total_frames; f += gridDim.x) {
        const long long clip = f / p.n_frames, j = f % p.n_frames;
        const float* xc = p.x + clip * p.clip_stride + p.start;
        const long long base = j * p.hop - p.pad;
        // windowed load, packed as z[n] = x[2n] + i x[2n+1]
        for (int n = tid; n < H; n += STFT_THREADS) {
            long long p0 = base + 2 * n, p1 = p0 + 1;
            float a = (p0 >= 0 && p0 < p.kept_len) ? __ldg(xc + p0) : 0.0f;
            float b = (p1 >= 0 && p1 < p.kept_len) ? __ldg(xc + p1) : 0.0f;
            buf0[n] = make_float2(a * win[2 * n], b * win[2 * n + 1]);
        }
        __syncthreads();
        float2* in = buf0;
        float2* out = buf1;
        int Ns = 1, lg = p.log2_half;
        if (lg & 1) {  // one radix-2 stage first
            for (int jj = tid; jj < H / 2; jj += STFT_THREADS) {
                float2 a = in[jj], b = in[jj + H / 2];
                out[2 * jj] = make_float2(a.x + b.x, a.y + b.y);
                out[2 * jj + 1] = make_float2(a.x - b.x, a.y - b.y);
            }
            __syncthreads();
            float2* t = in; in = out; out = t;
            Ns = 2;
        }
        for (; Ns < H; Ns *= 4) {
            const int Q = H / 4;
            const int tstep = H / (Ns * 4);
            for (int jj = tid; jj < Q; jj += STFT_THREADS) {
                const int k = jj & (Ns - 1);
                float2 v0 = in[jj], v1 = in[jj + Q], v2 = in[jj + 2 * Q], v3 = in[jj + 3 * Q];
                const int ti = k * tstep;
                v1 = cmul(v1, twH[ti]);
                v2 = cmul(v2, twH[2 * ti]);
                v3 = cmul(v3, twH[3 * ti]);
                float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
                float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
                const int j0 = ((jj - k) << 2) + k;  // (jj / Ns) * Ns * 4 + k
                out[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
                out[j0 + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 - i d13
                out[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
                out[j0 + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i d13
            }
            __syncthreads();
            float2* t = in; in = out; out = t;
        }
        // split post-pass: X[k] = (Z[k] + conj Z[H-k]) / 2 - i/2 e^{-2 pi i k / N} (Z[k] - conj Z[H-k])
        float* mrow = p.mag_out ? p.mag_out + f * (long long)nb : nullptr;
        for (int k = tid; k <= H; k += STFT_THREADS) {
            float2 a = in[k & (H - 1)], b = in[(H - k) & (H - 1)];
            float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
            float2 o = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
            float2 w = (k < H) ? twN[k] : make_float2(-1.0f, 0.0f);
            // -i * w * o
            float2 wo = cmul(w, o);
            float re = e.x + wo.y, im = e.y - wo.x;
            float m = sqrtf(re * re + im * im);
            mag[k] = m;
            if (mrow) __stcs(mrow + k, m);
        }
        __syncthreads();
        if (p.mel_out) {
            const int warp = tid >> 5, lane = tid & 31;
            float* orow = p.mel_out + f * (long long)p.n_mels;
            for (int m = warp; m < p.n_mels; m += STFT_THREADS / 32) {
                const int lo = s_rng[2 * m], hi = s_rng[2 * m + 1];
                const float* fb = p.mel_fb + (long long)m * nb;
                float acc = 0.0f;
                for (int k = lo + lane; k < hi; k += 32) acc = fmaf(__ldg(fb + k), mag[k], acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) orow[m] = acc;
            }
        }
        __syncthreads();
    }
}

static long long target_length(long long L, int n_fft, int hop, int center) {
    // STFT._fix_length through item_spec.convert (functionals.py:468-486, item_spec.py:58-98)
    long long extra = center ? 0 : (n_fft - hop);
    long long n = (L - extra >= 0 ? (L - extra) / hop : -((extra - L + hop - 1) / hop)) + (center ? 1 : 0);
    n -= center ? 1 : 0;
    return n * hop + extra;
}

}  // namespace mmk

using namespace mmk;

extern "C" int mmk_stft_n_frames(int64_t clip_len, int n_fft, int hop, int center, int align, int64_t* n_frames,
                                 int64_t* kept_len) {
    MMK_CHECK(n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "n_fft must be a power of two in [64, 4096]");
    MMK_CHECK(hop > 0 && clip_len >= 0, "hop must be > 0 and clip_len >= 0");
    long long kept = clip_len;
    if (align != 0) {
        kept = target_length(clip_len, n_fft, hop, center);
        if (kept == 0 && align == 1) kept = clip_len;  // python's x[-0:] keeps everything
        if (kept < 0) kept = 0;
        if (kept > clip_len) kept = clip_len;
    }
    long long padded = kept + (center ? n_fft : 0);
    long long nf = padded >= n_fft ? (padded - n_fft) / hop + 1 : 0;
    if (n_frames) *n_frames = nf;
    if (kept_len) *kept_len = kept;
    return 0;
}

extern "C" int mmk_stft_mag_mel(const float* d_x, int n_clips, int64_t clip_len, int64_t clip_stride, int n_fft,
                                int hop, int center, int align, float* d_mag_out, const float* d_mel_fb, int n_mels,
                                float* d_mel_out, void* stream) {
    int64_t n_frames = 0, kept = 0;
    if (int rc = mmk_stft_n_frames(clip_len, n_fft, hop, center, align, &n_frames, &kept)) return rc;
    MMK_CHECK(align >= 0 && align <= 2, "align must be 0 (none), 1 (end) or 2 (start)");
    MMK_CHECK(n_clips >= 0 && clip_stride >= clip_len, "bad clip geometry");
    MMK_CHECK(d_mel_out == nullptr || (d_mel_fb != nullptr && n_mels > 0), "mel output needs a filterbank");
    // torch.stft raises when the (padded) signal is shorter than n_fft; mirror that instead of returning nothing
    MMK_CHECK(n_frames > 0, "input too short for n_fft (the reference's torch.stft raises here as well)");
    if (n_clips == 0 || (d_mag_out == nullptr && d_mel_out == nullptr)) return 0;
    MMK_CHECK(d_x != nullptr, "null input");
    cudaStream_t st = (cudaStream_t)stream;
    StftParams p{};
    p.x = d_x; p.mag_out = d_mag_out; p.mel_fb = d_mel_fb; p.mel_out = d_mel_out;
    p.clip_stride = clip_stride;
    p.start = (align == 1) ? (clip_len - kept) : 0;
    p.kept_len = kept; p.n_frames = n_frames; p.total_frames = n_frames * (long long)n_clips;
    p.n_fft = n_fft; p.hop = hop; p.pad = center ? n_fft / 2 : 0; p.n_mels = d_mel_out ? n_mels : 0;
    int lg = 0; while ((1 << lg) < n_fft / 2) ++lg;
    p.log2_half = lg;
    MMK_CHECK(n_mels <= 8192, "n_mels must be <= 8192");
    if (n_fft == 2048 && n_mels <= FW_MAX_MELS) {   // warp-per-frame register FFT (stft_warp.cuh)
        const size_t fw_smem = FW_SMEM_BYTES + (d_mel_out ? sizeof(int) * 2 * (size_t)n_mels : 0);
        MMK_CUDA(cudaFuncSetAttribute(stft2048_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fw_smem));
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        long long grid = std::min<long long>(sms, (p.total_frames + FW_WARPS - 1) / FW_WARPS);
        p.mel_cap = d_mel_out ? FW_MEL_CAP : 0;
        stft2048_warp_kernel<<<(int)grid, FW_THREADS, fw_smem, st>>>(p);
        MMK_CUDA(cudaGetLastError());
        return 0;
    }
    const int H = n_fft / 2;
    size_t smem = sizeof(float2) * 4 * H + sizeof(float) * (n_fft + H + 8) + (d_mel_out ? sizeof(int) * 2 * (size_t)n_mels : 0);
    MMK_CUDA(cudaFuncSetAttribute(stft_mag_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    MMK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_mag_mel_kernel, STFT_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    if (grid > p.total_frames) grid = p.total_frames;
    stft_mag_mel_kernel<<<(int)grid, STFT_THREADS, smem, st>>>(p);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmk_mel_filterbank(int n_fft, int n_mels, float fmin, float fmax, int htk, float* h_out) {
    // librosa.filters.mel(sr=22050, ..., norm="slaney"), computed in fp64 and rounded to fp32 as librosa does.
    MMK_CHECK(h_out != nullptr && n_fft >= 2 && n_mels >= 1, "mmk_mel_filterbank: bad arguments");
    const double sr = 22050.0;
    const double fmx = fmax > 0.0f ? (double)fmax : sr / 2.0;
    const int nb = 1 + n_fft / 2;
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    auto hz2mel = [&](double f) {
        if (htk) return 2595.0 * log10(1.0 + f / 700.0);
        return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
    };
    auto mel2hz = [&](double m) {
        if (htk) return 700.0 * (pow(10.0, m / 2595.0) - 1.0);
        return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
    };
    std::vector<double> pts(n_mels + 2);
    const double m0 = hz2mel((double)fmin), m1 = hz2mel(fmx);
    for (int i = 0; i < n_mels + 2; ++i) pts[i] = mel2hz(m0 + (m1 - m0) * (double)i / (double)(n_mels + 1));
    for (int m = 0; m < n_mels; ++m) {
        const double enorm = 2.0 / (pts[m + 2] - pts[m]);
        for (int k = 0; k < nb; ++k) {
            const double fr = (sr / 2.0) * (double)k / (double)(nb - 1);
            const double lower = (fr - pts[m]) / (pts[m + 1] - pts[m]);
            const double upper = (pts[m + 2] - fr) / (pts[m + 2] - pts[m + 1]);
            double w = lower < upper ? lower : upper;
            if (w < 0.0) w = 0.0;
            h_out[(size_t)m * nb + k] = (float)(w * enorm);
        }
    }
    return 0;
}
