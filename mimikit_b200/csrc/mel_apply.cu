// MelSpec.np_func on precomputed magnitudes — mimikit/features/functionals.py:665-668:
// librosa.feature.melspectrogram(S=inputs.T, n_mels, fmin, fmax, htk).T = inputs @ mel_basis^T.
// (The fused STFT -> |.| -> mel kernels of stft_warp.cuh / stft_mel.cu serve MagSpec.mel; this is the stand-alone form.)
//
// HBM-bound by design: 4 (n_bins + n_mels) bytes per frame.  One warp per frame: the magnitude row is staged in shared
// memory with coalesced 16-byte loads, every lane then owns mel filters lane, lane + 32, ...; a triangular filter touches
// only its support [lo, hi), found once per CTA from the dense filterbank (so the arithmetic is ~2 n_bins FMAs per frame
// instead of n_mels x n_bins) and summed in ascending bin order, like the oracle's dense row dot product skips nothing
// but zeros.
#include "common.cuh"
#include <algorithm>
#include "../../include/mmk_b200.h"

namespace mmk_mel {

constexpr int NT = 256, NW = NT / 32;

__global__ void __launch_bounds__(NT) mel_apply_kernel(const float* __restrict__ mag, long long n_frames, int n_bins,
                                                       long long mag_stride, const float* __restrict__ fb, int n_mels,
                                                       float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    int2* sup = reinterpret_cast<int2*>(smem);                 // [n_mels] support [lo, hi) of every filter
    const int row_pad = (n_bins + 3) / 4 * 4;
    float* rows = smem + 2 * ((n_mels + 1) / 2 * 2);           // [NW][row_pad]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int m = tid; m < n_mels; m += NT) {
        const float* f = fb + (size_t)m * n_bins;
        int lo = n_bins, hi = 0;
        for (int k = 0; k < n_bins; ++k)
            if (__ldg(f + k) != 0.0f) { lo = min(lo, k); hi = k + 1; }
        sup[m] = make_int2(lo, hi > lo ? hi : lo);
    }
    __syncthreads();
    float* row = rows + warp * row_pad;
    const bool vec = (n_bins % 4 == 0) && (mag_stride % 4 == 0) && ((reinterpret_cast<size_t>(mag) & 15) == 0);
    for (long long f = (long long)blockIdx.x * NW + warp; f < n_frames; f += (long long)gridDim.x * NW) {
        const float* src = mag + f * mag_stride;
        if (vec) {
            for (int k = lane; k < n_bins / 4; k += 32)
                reinterpret_cast<float4*>(row)[k] = __ldcs(reinterpret_cast<const float4*>(src) + k);
        } else {
            for (int k = lane; k < n_bins; k += 32) row[k] = __ldcs(src + k);
        }
        __syncwarp();
        for (int m = lane; m < n_mels; m += 32) {
            const int2 s = sup[m];
            const float* w = fb + (size_t)m * n_bins;
            float acc = 0.0f;
            for (int k = s.x; k < s.y; ++k) acc = fmaf(__ldg(w + k), row[k], acc);
            __stcs(out + f * n_mels + m, acc);
        }
        __syncwarp();
    }
}

}  // namespace mmk_mel

extern "C" int mmk_mel_apply(const float* d_mag, int64_t n_frames, int n_bins, int64_t mag_stride, const float* d_mel_fb,
                             int n_mels, float* d_mel_out, void* stream) {
    if (n_frames == 0) return 0;
    MMK_CHECK(d_mag && d_mel_fb && d_mel_out, "null pointer");
    MMK_CHECK(n_frames >= 0 && n_bins >= 1 && n_bins <= 8193 && n_mels >= 1 && n_mels <= 1024, "bad mel geometry");
    MMK_CHECK(mag_stride >= n_bins, "row stride smaller than the row");
    if (n_frames == 0) return 0;
    int dev = 0, sms = 0;
    MMK_CUDA(cudaGetDevice(&dev));
    MMK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int row_pad = (n_bins + 3) / 4 * 4;
    const size_t smem = (size_t)(2 * ((n_mels + 1) / 2 * 2) + mmk_mel::NW * row_pad) * sizeof(float);
    MMK_CUDA(cudaFuncSetAttribute(mmk_mel::mel_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // a multiple of the SM count, 2 CTAs per SM, never more CTAs than there are frame octets
    long long want = (n_frames + mmk_mel::NW - 1) / mmk_mel::NW;
    int grid = (int)std::min<long long>(want, 2LL * sms);
    mmk_mel::mel_apply_kernel<<<grid, mmk_mel::NT, smem, (cudaStream_t)stream>>>(d_mag, n_frames, n_bins, mag_stride, d_mel_fb,
                                                                              n_mels, d_mel_out);
    MMK_CUDA(cudaGetLastError());
    return 0;
}
