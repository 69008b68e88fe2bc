// RemoveDC — mimikit/features/functionals.py:216-233 (np_func, the path the extraction pipeline uses: Compose(FileToSignal,
// Normalize(), RemoveDC()), io_spec.py:227-231): scipy.signal.lfilter([1, -1], [1, -0.99], x, axis=-1) on a float32 signal,
// i.e. scipy's float64 direct-form-II-transposed loop  y[n] = z + x[n];  z = x[n] * -1 - y[n] * -0.99  (zero initial
// state), cast back to float32.  (The reference's torch_func passes lfilter's arguments in the wrong order and cannot run.)
//
// Bit-exactness with the reference means evaluating exactly this chain of IEEE fp64 operations, which is sequential in
// time (3 dependent operations, ~165 cycles per sample).  Two kernels share one body:
//   * sequential: one LANE per clip walks the whole clip.  Exact by construction; parallel across clips only (10 h as
//     3 600 clips = 113 warps: 18.5 ms, bound by the dependent chain).
//   * speculative split in time (used when the caller provides scratch): a clip is cut into chunks of CH samples and every
//     (clip, chunk) gets a lane, which starts W samples early from a ZERO state.  The filter forgets (0.99^W), and once a
//     speculative trajectory hits the exact one bit for bit it stays on it, so after the warm-up the lane is — almost
//     surely — computing the exact sequence.  "Almost" is then removed by a check: chunk c is exact if chunk c-1 is and
//     the state it had reached at its first output sample EQUALS, bit for bit, the state chunk c-1 ended in (chunk 0 starts
//     from the true zero state).  A clip with any failed seam is recomputed by the sequential kernel on the same stream (a
//     no-op launch when nothing failed).  The result is therefore always the exact chain; the split only changes how long
//     it takes: 10 h in ~1.5 ms instead of 18.5 (default geometry: CH = W = 8 192 samples).
// In both, a warp owns 32 units and moves them through a 32 x 32 shared-memory tile so that every global access is a
// coalesced 128-byte row segment; the next tile's loads are in flight while the current one is filtered.
// Roofline: HBM, 8 B per sample (4 in + 4 out; the warm-up re-reads W / CH of the input, mostly from L2).
#include "common.cuh"
#include "../../include/mmk_b200.h"

#include <cstdlib>

namespace mmk {

constexpr int DC_WARPS = 4;

struct DcParams {
    const float* x; float* out;
    long long n_rows, row_len, row_stride;
    long long n_ch, CH, W;           // chunks per row, chunk length, warm-up (multiples of 32); n_ch == 0: sequential mode
    double* z_in; double* z_out;     // [n_rows * n_ch] state at a chunk's first output sample / after its last one
    int* row_flag;                   // [n_rows] set when a seam of the row failed (sequential mode: rows to recompute, or null = all)
};

template <bool spec>
__global__ void __launch_bounds__(32 * DC_WARPS) remove_dc_kernel(const DcParams p) {
    __shared__ float tile_s[DC_WARPS][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*tile)[33] = tile_s[warp];
    const long long n_units = spec ? p.n_rows * p.n_ch : p.n_rows;
    const long long u0 = ((long long)blockIdx.x * DC_WARPS + warp) * 32;
    if (u0 >= n_units) return;
    const int units = (int)min(32LL, n_units - u0);
    // this lane's unit: row, first time index it visits (may be negative: zeros), first / one-past-last output index
    long long my_row = 0, my_t0 = 0, my_o0 = 0, my_o1 = 0;
    bool active = lane < units;
    if (active) {
        const long long u = u0 + lane;
        if (spec) {
            my_row = u / p.n_ch;
            const long long c = u - my_row * p.n_ch;
            my_o0 = c * p.CH; my_o1 = min(p.row_len, my_o0 + p.CH); my_t0 = my_o0 - p.W;
        } else {
            my_row = u; my_t0 = 0; my_o0 = 0; my_o1 = p.row_len;
            if (p.row_flag && p.row_flag[my_row] == 0) active = false;
        }
    }
    const unsigned amask = __ballot_sync(0xffffffffu, active);
    if (amask == 0u) return;                             // sequential mode after a clean speculative pass: nothing to redo
    const long long span = spec ? p.W + p.CH : p.row_len;
    // per-unit geometry of the other lanes, for the coalesced tile traffic (unit r's values live in lane r)
    float nxt[32];
    auto fetch = [&](long long k) {                      // tile k: every unit's samples [t0 + 32 k, + 32)
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const bool act = (amask >> r) & 1u;
            if (spec) {
                const long long row = __shfl_sync(0xffffffffu, my_row, r), t = __shfl_sync(0xffffffffu, my_t0, r) + 32 * k + lane;
                nxt[r] = (act && t >= 0 && t < p.row_len) ? __ldg(p.x + row * p.row_stride + t) : 0.0f;
            } else {                                     // unit r is row u0 + r, from time 0: closed-form addresses
                const long long t = 32 * k + lane;
                nxt[r] = (act && t < p.row_len) ? __ldcs(p.x + (u0 + r) * p.row_stride + t) : 0.0f;
            }
        }
    };
    fetch(0);
    double z = 0.0;
    for (long long k = 0; 32 * k < span; ++k) {
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[r][lane] = nxt[r];
        __syncwarp();
        if (32 * (k + 1) < span) fetch(k + 1);
        if (spec && active && my_t0 + 32 * k == my_o0) p.z_in[u0 + lane] = z;     // state on entering the chunk proper
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {       // this lane's unit: scipy's loop, operation for operation
            const double xn = (double)tile[lane][i];
            const double yn = __dadd_rn(z, __dmul_rn(1.0, xn));
            z = __dsub_rn(__dmul_rn(xn, -1.0), __dmul_rn(yn, -0.99));
            tile[lane][i] = (float)yn;
            if (spec && my_t0 + 32 * k + i == my_o1 - 1 && active) p.z_out[u0 + lane] = z;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const bool act = (amask >> r) & 1u;
            if (spec) {
                const long long row = __shfl_sync(0xffffffffu, my_row, r), t = __shfl_sync(0xffffffffu, my_t0, r) + 32 * k + lane;
                const long long o0 = __shfl_sync(0xffffffffu, my_o0, r), o1 = __shfl_sync(0xffffffffu, my_o1, r);
                if (act && t >= o0 && t < o1) __stcs(p.out + row * p.row_len + t, tile[r][lane]);
            } else {
                const long long t = 32 * k + lane;
                if (act && t < p.row_len) __stcs(p.out + (u0 + r) * p.row_len + t, tile[r][lane]);
            }
        }
        __syncwarp();
    }
}

// seam c of a row holds if the speculative chunk entered with exactly the state its predecessor left
__global__ void remove_dc_check_kernel(const double* __restrict__ z_in, const double* __restrict__ z_out, long long n_rows,
                                       long long n_ch, int* __restrict__ row_flag) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_rows * n_ch) return;
    const long long row = u / n_ch, c = u - row * n_ch;
    if (c > 0 && __double_as_longlong(z_in[u]) != __double_as_longlong(z_out[u - 1])) row_flag[row] = 1;
}

// Chunk / warm-up lengths.  0.99^k needs ~3 700 samples to bring a unit difference below one fp64 ulp; from there two
// trajectories merge with probability ~1 % per sample (the rounding of 0.99 y maps neighbours to one value), so a warm-up
// of 8 192 leaves ~4 500 samples for the merge: a failed seam is then a ~e^-45 event (4 096 was measured to fail a few
// seams in 10 h of audio).  Rows shorter than 4 (W + CH) are not worth splitting.  Returns false for "do not split".
static bool dc_geometry(int64_t row_len, long long* CH, long long* W, long long* n_ch) {
    long long ch = 8192, w = 8192;
    bool forced = false;
    if (const char* e = getenv("MMK_DC_CHUNK")) { ch = std::max(32LL, atoll(e) / 32 * 32); forced = true; }
    if (const char* e = getenv("MMK_DC_WARMUP")) { w = std::max(0LL, atoll(e) / 32 * 32); forced = true; }
    *CH = ch; *W = w; *n_ch = (row_len + ch - 1) / ch;
    return forced ? *n_ch >= 2 : row_len >= 4 * (w + ch);
}

}  // namespace mmk

using namespace mmk;

extern "C" size_t mmk_remove_dc_scratch_bytes(int64_t n_rows, int64_t row_len) {
    if (n_rows <= 0 || row_len <= 0) return 0;
    long long CH, W, n_ch;
    if (!dc_geometry(row_len, &CH, &W, &n_ch)) return 0;
    return (size_t)n_rows * (size_t)n_ch * 2 * sizeof(double) + (size_t)n_rows * sizeof(int);
}

extern "C" int mmk_remove_dc(const float* d_x, float* d_out, int64_t n_rows, int64_t row_len, int64_t row_stride,
                             void* d_scratch, size_t scratch_bytes, void* stream) {
    MMK_CHECK(n_rows >= 0 && row_len >= 0 && row_stride >= row_len, "mmk_remove_dc: bad geometry");
    if (n_rows == 0 || row_len == 0) return 0;
    MMK_CHECK(d_x && d_out, "mmk_remove_dc: null pointer");
    MMK_CHECK(d_x != d_out, "mmk_remove_dc: in-place operation is not supported");
    cudaStream_t st = (cudaStream_t)stream;
    DcParams p{};
    p.x = d_x; p.out = d_out; p.n_rows = n_rows; p.row_len = row_len; p.row_stride = row_stride;
    auto launch = [&](long long n_units) -> int {
        const long long grid = (n_units + 32 * DC_WARPS - 1) / (32 * DC_WARPS);
        MMK_CHECK(grid <= 0x7fffffffLL, "mmk_remove_dc: too many rows");
        if (p.n_ch > 0) remove_dc_kernel<true><<<(unsigned)grid, 32 * DC_WARPS, 0, st>>>(p);
        else remove_dc_kernel<false><<<(unsigned)grid, 32 * DC_WARPS, 0, st>>>(p);
        MMK_CUDA(cudaGetLastError());
        return 0;
    };
    const size_t need = mmk_remove_dc_scratch_bytes(n_rows, row_len);
    if (need == 0 || d_scratch == nullptr || scratch_bytes < need) return launch(n_rows);   // sequential: exact by construction
    long long CH, W, n_ch;
    dc_geometry(row_len, &CH, &W, &n_ch);
    MMK_CHECK(((uintptr_t)d_scratch & 7u) == 0, "mmk_remove_dc: scratch must be 8-byte aligned");
    p.n_ch = n_ch; p.CH = CH; p.W = W;
    p.z_in = reinterpret_cast<double*>(d_scratch);
    p.z_out = p.z_in + (size_t)n_rows * n_ch;
    p.row_flag = reinterpret_cast<int*>(p.z_out + (size_t)n_rows * n_ch);
    MMK_CUDA(cudaMemsetAsync(p.row_flag, 0, (size_t)n_rows * sizeof(int), st));
    if (int rc = launch(n_rows * n_ch)) return rc;
    const long long n_units = n_rows * n_ch;
    remove_dc_check_kernel<<<(unsigned)((n_units + 255) / 256), 256, 0, st>>>(p.z_in, p.z_out, n_rows, n_ch, p.row_flag);
    MMK_CUDA(cudaGetLastError());
    p.n_ch = 0;                              // rows with a failed seam: the exact sequential chain (returns at once otherwise)
    return launch(n_rows);
}

// Diagnostic: how many rows the last speculative call on this scratch had to recompute sequentially (synchronises).
extern "C" int mmk_remove_dc_recomputed_rows(const void* d_scratch, int64_t n_rows, int64_t row_len, int64_t* h_count, void* stream) {
    MMK_CHECK(h_count, "null argument");
    *h_count = 0;
    const size_t need = mmk_remove_dc_scratch_bytes(n_rows, row_len);
    if (need == 0 || !d_scratch) return 0;
    long long CH, W, n_ch;
    dc_geometry(row_len, &CH, &W, &n_ch);
    int* flags = new int[(size_t)n_rows];
    const int* d_flags = reinterpret_cast<const int*>(reinterpret_cast<const double*>(d_scratch) + (size_t)n_rows * n_ch * 2);
    cudaError_t e = cudaMemcpyAsync(flags, d_flags, (size_t)n_rows * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    long long cnt = 0;
    for (int64_t i = 0; i < n_rows; ++i) cnt += flags[i] != 0;
    delete[] flags;
    MMK_CHECK(e == cudaSuccess, std::string("mmk_remove_dc_recomputed_rows: ") + cudaGetErrorString(e));
    *h_count = cnt;
    return 0;
}
