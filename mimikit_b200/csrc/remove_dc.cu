// RemoveDC — mimikit/features/functionals.py:216-233 (np_func, the path the extraction pipeline uses: Compose(FileToSignal,
// Normalize(), RemoveDC()), io_spec.py:227-231): scipy.signal.lfilter([1, -1], [1, -0.99], x, axis=-1) on a float32 signal,
// i.e. scipy's float64 direct-form-II-transposed loop  y[n] = z + x[n];  z = x[n] * -1 - y[n] * -0.99  (zero initial
// state), cast back to float32.  (The reference's torch_func passes lfilter's arguments in the wrong order and cannot run.)
//
// The recurrence is sequential in time, so bit-exactness with the reference means evaluating exactly this chain: one LANE
// per clip walks its samples with IEEE fp64 adds / multiplies in scipy's order; the parallelism is across clips.  A warp
// owns 32 clips and moves them through a 32 x 32 shared-memory tile so that every global access is a coalesced 128-byte
// row segment; the next tile's loads are in flight while the current one is filtered.
// Roofline: HBM, 8 B per sample (4 in + 4 out); with few clips the kernel is bound by the dependent fp64 chain instead
// (3 operations per sample per clip): 10 h as 3 600 clips = 113 warps — one per SM.
#include "common.cuh"
#include "../../include/mmk_b200.h"

namespace mmk {

__global__ void __launch_bounds__(32) remove_dc_kernel(const float* __restrict__ x, float* __restrict__ out, long long n_rows,
                                                       long long row_len, long long row_stride) {
    __shared__ float tile[32][33];
    const int lane = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * 32;
    const int rows = (int)min(32LL, n_rows - row0);
    float nxt[32];
    auto fetch = [&](long long t0) {
#pragma unroll
        for (int r = 0; r < 32; ++r)
            nxt[r] = (r < rows && t0 + lane < row_len) ? __ldcs(x + (row0 + r) * row_stride + t0 + lane) : 0.0f;
    };
    fetch(0);
    double z = 0.0;
    for (long long t0 = 0; t0 < row_len; t0 += 32) {
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[r][lane] = nxt[r];
        __syncwarp();
        if (t0 + 32 < row_len) fetch(t0 + 32);
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {       // this lane's clip: scipy's loop, operation for operation
            const double xn = (double)tile[lane][i];
            const double yn = __dadd_rn(z, __dmul_rn(1.0, xn));
            z = __dsub_rn(__dmul_rn(xn, -1.0), __dmul_rn(yn, -0.99));
            tile[lane][i] = (float)yn;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r)
            if (r < rows && t0 + lane < row_len) __stcs(out + (row0 + r) * row_len + t0 + lane, tile[r][lane]);
        __syncwarp();
    }
}

}  // namespace mmk

using namespace mmk;

extern "C" int mmk_remove_dc(const float* d_x, float* d_out, int64_t n_rows, int64_t row_len, int64_t row_stride, void* stream) {
    MMK_CHECK(n_rows >= 0 && row_len >= 0 && row_stride >= row_len, "mmk_remove_dc: bad geometry");
    if (n_rows == 0 || row_len == 0) return 0;
    MMK_CHECK(d_x && d_out, "mmk_remove_dc: null pointer");
    MMK_CHECK(d_x != d_out, "mmk_remove_dc: in-place operation is not supported");
    const long long grid = (n_rows + 31) / 32;
    MMK_CHECK(grid <= 0x7fffffffLL, "mmk_remove_dc: too many rows");
    remove_dc_kernel<<<(unsigned)grid, 32, 0, (cudaStream_t)stream>>>(d_x, d_out, n_rows, row_len, row_stride);
    MMK_CUDA(cudaGetLastError());
    return 0;
}
