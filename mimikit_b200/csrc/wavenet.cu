// Persistent, cluster-pipelined WaveNet generation kernel (sm_100a).
//
// Replaces the reference's per-sample Python loop: GenerateLoopV2.run (mimikit/loops/generate.py:184-229) calling
// WaveNet.generate_step == WaveNet.forward on the last rf samples (mimikit/networks/wavenet_v2.py:447-452,
// 276-293; WNLayer.forward 131-176), the MLP head (networks/mlp.py:44-63) and CategoricalSampler
// (modules/targets.py:40-52).  The reference recomputes the whole receptive field for every sample; here each
// layer keeps a ring of its last d_l inputs, which is the same function of the samples.
//
// Design:
//   * ONE launch covers all time steps.  The grid is NST thread-block clusters of CS CTAs.  Cluster ("stage") s
//     owns a contiguous range of layers whose weights stay resident in its shared memory for the whole launch;
//     inside a stage every 1x1 / 2-tap contraction is split by OUTPUT channel over the CS CTAs, the gated output
//     and the next layer input are all-gathered through distributed shared memory (st to peers + barrier.cluster).
//   * The batch is cut into groups of GB prompts that flow through the stages as a pipeline; stage s works on
//     group g while stage s+1 works on group g-1.  Hand-off between stages goes through an L2-resident mailbox
//     with release/acquire counters; the last stage samples and publishes the new sample index, which stage 0
//     picks up for the next time step.  No grid-wide barrier and no collective anywhere in the step loop.
//   * fp32 FFMA throughout (bit-exact sequences against the fp32 reference need fp32 products; at GB = 8 rows the
//     contraction is too small for a tcgen05 tile to pay for its TMEM round trip).
//   * Gate, residual add, skip accumulation, head (Linear-Mish-Linear, learned temperature), softmax and
//     inverse-CDF sampling from externally supplied uniform noise are fused; nothing but the sampled index (and,
//     on request, the logits) leaves the chip.
#include "common.cuh"
#include "sampler.cuh"
#include "wavenet_impl.h"
#include "../../include/mmk_b200.h"

#include <cooperative_groups.h>
#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;

namespace mmk {

constexpr int WN_NT = 256;   // threads per CTA
constexpr int WN_GB = 8;     // prompts per pipeline group
constexpr int WN_MAX_LAYERS = 96;
constexpr int WN_MAX_STAGES = 32;
constexpr unsigned WN_SPIN_LIMIT = 1u << 24;

constexpr int WN_MAX_K = 4;  // conv kernel sizes 2..4

struct WnLayer {
    int dilation;
    int nres;             // residual output columns per CTA (nf or 0)
    int ksize;            // taps of the dilated conv (tap j reads the layer input (ksize - 1 - j) dilations back)
    long long ring_off;   // float offset of this layer's ring in the ring buffer: (ksize - 1) * dilation slots
};

struct WnParams {
    // network geometry
    int L, C, S, Hh, Q, Kh;
    int CS, NST;
    int nf, ns, nh, nz;        // per-CTA output columns: gate channels, skip, head hidden, head logits
    int NA, NB, NH, NZ;        // the same, padded to multiples of 4 (NA covers f and g: 2*nf)
    int layer_block;           // floats per (layer, rank) weight block
    int head_block;            // floats per rank head block
    int kmax;                  // largest conv kernel size of the network
    int layerwise;             // layerwise_inputs (wavenet_v2.py:283-284): the embedded input is added to every layer's output
    int n_hh;                  // hidden layers of the MLP head (one shared Linear, mlp.py:47-50)
    int act_f, act_g;          // MMK_ACT_* of the gated unit (resolved: never MMK_ACT_DEFAULT)
    int affine;                // with_affine_residuals: every layer input goes through aff_res first (wavenet_v2.py:148-149)
    int NC;                    // 3 * nf padded to a multiple of 4 (x_hat | a | b columns of aff_res), 0 without it
    int G;                     // groups the rings are laid out for
    float min_temp;
    WnLayer layers[WN_MAX_LAYERS];
    int stage_lo[WN_MAX_STAGES + 1];
    const float* wpack;
    const float* hpack;
    const float* E;
    float* rings;
    float* mail_h;             // [NST][G][C][GB]
    float* mail_s;             // [NST][G][S][GB]
    unsigned* ready;           // [NST][G]
    unsigned* ack;             // [NST][G]
    long long* avail;          // [G]
    unsigned* abort_flag;
    // this run
    long long* seq;
    long long seq_stride, t_begin, t_head, t_end;
    int B, n_groups, teacher_forced;
    const float* temperature;
    int n_temperature;
    const float* noise;
    long long noise_stride, noise_t0;
    float* logits_out;
    long long* decisions;
    unsigned long long* step_ts;
    // shared-memory carve-up (float offsets)
    int off_w, off_head, off_x1, off_x0, off_y, off_sacc, off_hin, off_hid, off_z, off_part, off_slice, off_e, smem_floats;
    int zrow;                  // padded row length of the logits buffer
};

// ------------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_acquire_s64(const long long* p) {
    long long v;
    asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_s64(long long* p, long long v) {
    asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int CS>
__device__ __forceinline__ void cluster_sync_all() {
    if constexpr (CS == 1) {
        __syncthreads();
    } else {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

// Thread 0 spins until *flag >= target (or the launch is aborted), then the CTA syncs.
__device__ __forceinline__ void wait_u32(const unsigned* flag, unsigned target, unsigned* abort_flag) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_u32(flag) < target) {
            if (++spins > WN_SPIN_LIMIT || (((spins & 63u) == 1u) && ld_acquire_u32(abort_flag) != 0u)) {
                atomicExch(abort_flag, 1u);
                break;
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void wait_s64(const long long* flag, long long target, unsigned* abort_flag) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_s64(flag) < target) {
            if (++spins > WN_SPIN_LIMIT || (((spins & 63u) == 1u) && ld_acquire_u32(abort_flag) != 0u)) {
                atomicExch(abort_flag, 1u);
                break;
            }
        }
    }
    __syncthreads();
}


// act_f / act_g of the gated unit: the point-wise members of ActivationEnum (modules/activations.py:26-40, 70-83)
__device__ __forceinline__ float apply_act(int code, float x) {
    switch (code) {
        case MMK_ACT_TANH: return tanhf(x);
        case MMK_ACT_SIGMOID: return sigmoid_acc(x);
        case MMK_ACT_MISH: return mish_acc(x);
        case MMK_ACT_RELU: return fmaxf(x, 0.0f);
        case MMK_ACT_SOFTPLUS: return x > 20.0f ? x : log1pf(expf(x));
        case MMK_ACT_IDENTITY: return x;
        case MMK_ACT_ABS: return fabsf(x);
        case MMK_ACT_SIN: return sinf(x);
        case MMK_ACT_COS: return cosf(x);
    }
    return x;
}

// ------------------------------------------------------------------------------------------------------------
// Slice contraction.  out[col][p] = sum_k W[k][col] * x[k][p] for the CTA's ncolp (multiple of 4) columns and GB
// prompts.  The (GB/2) x (ncolp/4) register tiles of 2 prompts x 4 columns are replicated over `nslice` K-slices
// (slice s takes k = s, s + nslice, ...); partial sums meet in shared memory and are added in slice order, so the
// summation order is fixed (deterministic run to run).
// ------------------------------------------------------------------------------------------------------------
struct Tile {
    int slice, nslice, pp, cq, lout;
    bool active;
};
__device__ __forceinline__ Tile make_tile(int ncolp) {
    Tile t;
    t.lout = (WN_GB / 2) * (ncolp / 4);
    t.nslice = WN_NT / t.lout;
    t.slice = threadIdx.x / t.lout;
    int tile = threadIdx.x % t.lout;
    t.pp = tile % (WN_GB / 2);
    t.cq = tile / (WN_GB / 2);
    t.active = t.slice < t.nslice;
    return t;
}
__device__ __forceinline__ void gemm_accum(float (&acc)[8], const Tile& t, const float* __restrict__ Ws, int ldw,
                                           const float* __restrict__ xs, int K) {
    if (!t.active) return;
    const float* wp = Ws + t.cq * 4;
    const float* xp = xs + t.pp * 2;
#pragma unroll 4
    for (int k = t.slice; k < K; k += t.nslice) {
        const float4 w = *reinterpret_cast<const float4*>(wp + k * ldw);
        const float2 x = *reinterpret_cast<const float2*>(xp + k * WN_GB);
        acc[0] = fmaf(x.x, w.x, acc[0]); acc[1] = fmaf(x.x, w.y, acc[1]);
        acc[2] = fmaf(x.x, w.z, acc[2]); acc[3] = fmaf(x.x, w.w, acc[3]);
        acc[4] = fmaf(x.y, w.x, acc[4]); acc[5] = fmaf(x.y, w.y, acc[5]);
        acc[6] = fmaf(x.y, w.z, acc[6]); acc[7] = fmaf(x.y, w.w, acc[7]);
    }
}
__device__ __forceinline__ void gemm_store_partials(const float (&acc)[8], const Tile& t, float* part) {
    if (t.active) {
        float4* d = reinterpret_cast<float4*>(part + (size_t)threadIdx.x * 8);
        d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}
// sum of the partials of output (col, p) — call after __syncthreads()
__device__ __forceinline__ float gemm_reduce(const float* part, int ncolp, int col, int p) {
    const int lout = (WN_GB / 2) * (ncolp / 4), nslice = WN_NT / lout;
    const int tile = (col >> 2) * (WN_GB / 2) + (p >> 1);
    const float* q = part + tile * 8 + (p & 1) * 4 + (col & 3);
    float s = 0.0f;
    for (int sl = 0; sl < nslice; ++sl) s += q[(size_t)sl * lout * 8];
    return s;
}

// Copies the CTA's contiguous slice (nfl4 float4s) to the same offset of `buf` in every CTA of the cluster.
template <int CS>
__device__ __forceinline__ void scatter_slice(cg::cluster_group& cluster, float* buf_slice_local, const float* src,
                                              int nfl4) {
    if constexpr (CS == 1) {
        for (int i = threadIdx.x; i < nfl4; i += WN_NT)
            reinterpret_cast<float4*>(buf_slice_local)[i] = reinterpret_cast<const float4*>(src)[i];
    } else {
        for (int i = threadIdx.x; i < nfl4 * CS; i += WN_NT) {
            const int peer = i / nfl4, q = i - peer * nfl4;
            float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(buf_slice_local, peer));
            dst[q] = reinterpret_cast<const float4*>(src)[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------------------
template <int CS>
__global__ void __launch_bounds__(WN_NT, 1) wavenet_pipe_kernel(const __grid_constant__ WnParams P) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (CS == 1) ? 0 : (int)cluster.block_rank();
    const int stage = blockIdx.x / CS;
    const int l_lo = P.stage_lo[stage], l_hi = P.stage_lo[stage + 1];
    const bool first_stage = stage == 0, last_stage = stage == P.NST - 1;
    constexpr int GB = WN_GB;
    const int C = P.C, S = P.S, nf = P.nf, ns = P.ns;

    float* w_s = smem + P.off_w;
    float* head_s = smem + P.off_head;
    float* x1 = smem + P.off_x1;      // [C][GB]   current layer input h_l(t)
    float* x0 = smem + P.off_x0;      // [2][kmax-1][C][GB] ring reads h_l(t - a d), a = k-1 .. 1, double buffered
    float* eown = smem + P.off_e;     // [nf][GB] this CTA's channels of the embedded network input (layerwise_inputs)
    float* ybuf = smem + P.off_y;     // [2][C][GB] gated outputs, double buffered
    float* sacc = smem + P.off_sacc;  // [ns][GB]   this CTA's slice of the running skip sum
    float* hin = smem + P.off_hin;    // [Kh][GB]   head input
    float* hid = smem + P.off_hid;    // [Hh][GB]
    float* zbuf = smem + P.off_z;     // [GB][zrow] raw head outputs (rank 0 only)
    float* part = smem + P.off_part;  // [NT*8]
    float* slice = smem + P.off_slice;  // staging of this CTA's freshly computed slice

    // resident weights: straight copy of this CTA's packed blocks
    {
        const int n_own = l_hi - l_lo;
        const float4* src = reinterpret_cast<const float4*>(P.wpack + ((size_t)l_lo * CS) * P.layer_block);
        for (int l = 0; l < n_own; ++l) {
            const float4* s4 = src + ((size_t)(l * CS + rank) * P.layer_block) / 4;
            float4* d4 = reinterpret_cast<float4*>(w_s + (size_t)l * P.layer_block);
            for (int i = tid; i < P.layer_block / 4; i += WN_NT) d4[i] = __ldg(s4 + i);
        }
        if (last_stage) {
            const float4* s4 = reinterpret_cast<const float4*>(P.hpack + (size_t)rank * P.head_block);
            float4* d4 = reinterpret_cast<float4*>(head_s);
            for (int i = tid; i < P.head_block / 4; i += WN_NT) d4[i] = __ldg(s4 + i);
        }
    }
    __syncthreads();
    cluster_sync_all<CS>();

    const size_t blk = (size_t)C * GB;   // floats of one (group) activation block
    const size_t x0set = (size_t)(P.kmax - 1) * blk;   // one parity of the tap buffers
    unsigned par = 0;                    // parity of the double-buffered x0 / ybuf, advances once per layer
    // the older taps of layer ly at time t: tap j (j < k - 1) is the input (k - 1 - j) dilations back; the ring has
    // (k - 1) d slots, the slot of time t' is t' mod that
    auto prefetch_taps = [&](const WnLayer& ly, long long t, int g, float* dst) {
        const long long R = (long long)(ly.ksize - 1) * ly.dilation;
        for (int j = 0; j < ly.ksize - 1; ++j) {
            long long sl = (t - (long long)(ly.ksize - 1 - j) * ly.dilation) % R;
            if (sl < 0) sl += R;
            const float* src = P.rings + ly.ring_off + ((size_t)sl * P.G + g) * blk;
            for (int i = tid; i < (int)(blk / 4); i += WN_NT) cp_async16(dst + (size_t)j * blk + i * 4, src + i * 4);
        }
    };

    for (long long t = P.t_begin; t < P.t_end; ++t) {
        const unsigned delivery = (unsigned)(t - P.t_begin);
        const bool head_on = t >= P.t_head;
        for (int g = 0; g < P.n_groups; ++g) {
            // ---------------- stage input ----------------
            prefetch_taps(P.layers[l_lo], t, g, x0 + par * x0set);   // ring reads of the first owned layer
            cp_async_commit();
            if (first_stage) {
                wait_s64(P.avail + g, t + 1, P.abort_flag);
                // embedding gather: x1[k][p] = E[q_{b,t}][k]   (EmbeddingIO, modules/io.py:148-154)
                if (warp < GB) {
                    const int b = g * GB + warp;
                    long long q = 0;
                    if (b < P.B) q = __ldcg(P.seq + (size_t)b * P.seq_stride + t);
                    q = q < 0 ? 0 : (q >= P.Q ? P.Q - 1 : q);
                    const float* row = P.E + (size_t)q * C;
                    for (int k = lane; k < C; k += 32) x1[k * GB + warp] = (b < P.B) ? __ldg(row + k) : 0.0f;
                }
                for (int i = tid; i < ns * GB; i += WN_NT) sacc[i] = 0.0f;
            } else {
                wait_u32(P.ready + stage * P.G + g, (delivery + 1) * CS, P.abort_flag);
                const float4* mh = reinterpret_cast<const float4*>(P.mail_h + ((size_t)stage * P.G + g) * blk);
                for (int i = tid; i < (int)(blk / 4); i += WN_NT) reinterpret_cast<float4*>(x1)[i] = __ldcg(mh + i);
                if (ns > 0) {
                    const float* ms = P.mail_s + ((size_t)stage * P.G + g) * (size_t)S * GB + (size_t)rank * ns * GB;
                    for (int i = tid; i < ns * GB; i += WN_NT) sacc[i] = __ldcg(ms + i);
                }
                __syncthreads();
                if (tid == 0) red_release_add(P.ack + stage * P.G + g, 1u);
            }
            if (P.layerwise) {   // this CTA's channels of E[q_{b,t}] (the sample is in seq: it was published before the group
                                 // was handed to this stage)
                for (int o = tid; o < nf * GB; o += WN_NT) {
                    const int i = o / GB, pp = o - i * GB, b = g * GB + pp;
                    long long q = 0;
                    if (b < P.B) q = __ldcg(P.seq + (size_t)b * P.seq_stride + t);
                    q = q < 0 ? 0 : (q >= P.Q ? P.Q - 1 : q);
                    eown[o] = (b < P.B) ? __ldg(P.E + (size_t)q * C + rank * nf + i) : 0.0f;
                }
            }
            __syncthreads();

            // ---------------- owned layers ----------------
            for (int l = l_lo; l < l_hi; ++l) {
                const WnLayer& ly = P.layers[l];
                const float* W1 = w_s + (size_t)(l - l_lo) * P.layer_block;   // [kmax C][NA], tap j at rows [j C, (j+1) C)
                const float* b1 = W1 + (size_t)P.kmax * C * P.NA;             // [NA]
                const float* W2 = b1 + P.NA;                                  // [C][NB]
                const float* b2 = W2 + (size_t)C * P.NB;                      // [NB]
                float* x0c = x0 + par * x0set;
                float* yc = ybuf + par * blk;
                const bool last_owned = (l == l_hi - 1);
                const bool last_layer = (l == P.L - 1);

                // (0) with_affine_residuals: the layer input becomes z = x_hat * a + b, (x_hat | a | b) = aff_res.params(h_l)
                //     (parametrized.py:44-47: mul, then add — two roundings).  Everything below (taps, ring, residual) reads z.
                if (P.affine) {
                    const float* W3 = b2 + P.NB;                                  // [C][NC]
                    const float* b3 = W3 + (size_t)C * P.NC;                      // [NC]
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const Tile tl = make_tile(P.NC);
                    gemm_accum(acc, tl, W3, P.NC, x1, C);
                    gemm_store_partials(acc, tl, part);
                    __syncthreads();
                    for (int o = tid; o < nf * GB; o += WN_NT) {
                        const int i = o / GB, p = o - i * GB;
                        const float xh = gemm_reduce(part, P.NC, i, p) + b3[i];
                        const float aa = gemm_reduce(part, P.NC, nf + i, p) + b3[nf + i];
                        const float bb = gemm_reduce(part, P.NC, 2 * nf + i, p) + b3[2 * nf + i];
                        slice[o] = __fadd_rn(__fmul_rn(xh, aa), bb);
                    }
                    cluster_sync_all<CS>();   // every CTA of the cluster is done reading h_l from its x1
                    scatter_slice<CS>(cluster, x1 + (size_t)rank * nf * GB, slice, nf * GB / 4);
                    cluster_sync_all<CS>();   // z complete everywhere
                }

                // (a) ring read landed everywhere in this CTA
                cp_async_wait_all();
                __syncthreads();

                // (b) gated unit: a = Wd[:, :, 0] h(t-d) + Wd[:, :, 1] h(t) + b ; y = tanh(a_f) * sigmoid(a_g)
                {
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const Tile tl = make_tile(P.NA);
                    for (int j = 0; j < ly.ksize - 1; ++j)
                        gemm_accum(acc, tl, W1 + (size_t)j * C * P.NA, P.NA, x0c + (size_t)j * blk, C);
                    gemm_accum(acc, tl, W1 + (size_t)(ly.ksize - 1) * C * P.NA, P.NA, x1, C);
                    gemm_store_partials(acc, tl, part);
                    __syncthreads();
                    for (int o = tid; o < nf * GB; o += WN_NT) {
                        const int i = o / GB, p = o - i * GB;
                        const float f = gemm_reduce(part, P.NA, i, p) + b1[i];
                        const float gg = gemm_reduce(part, P.NA, nf + i, p) + b1[nf + i];
                        slice[o] = apply_act(P.act_f, f) * apply_act(P.act_g, gg);
                    }
                    __syncthreads();
                    scatter_slice<CS>(cluster, yc + (size_t)rank * nf * GB, slice, nf * GB / 4);
                }
                cluster_sync_all<CS>();   // #1: y complete everywhere; every CTA is past its ring read of layer l

                // (d) ring write of this layer's input (own channel slice), then prefetch the next ring read
                {
                    const long long Rl = (long long)(ly.ksize - 1) * ly.dilation;
                    float* dst = P.rings + ly.ring_off + ((size_t)(t % Rl) * P.G + g) * blk + (size_t)rank * nf * GB;
                    const float* src = x1 + (size_t)rank * nf * GB;
                    for (int i = tid; i < nf * GB / 4; i += WN_NT)
                        __stcg(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(src)[i]);
                    if (!last_owned) prefetch_taps(P.layers[l + 1], t, g, x0 + (par ^ 1u) * x0set);
                    cp_async_commit();
                }

                // (e) skip and residual 1x1 convs on the full gated vector
                const int ncol2 = ly.nres + ns;
                if (ncol2 > 0) {
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const Tile tl = make_tile(P.NB);
                    gemm_accum(acc, tl, W2, P.NB, yc, C);
                    gemm_store_partials(acc, tl, part);
                    __syncthreads();
                    for (int o = tid; o < ncol2 * GB; o += WN_NT) {
                        const int i = o / GB, p = o - i * GB;
                        const float v = gemm_reduce(part, P.NB, i, p) + b2[i];
                        if (i < ly.nres) {
                            float hn = x1[((size_t)rank * nf + i) * GB + p] + v;   // h_{l+1} = h_l + conv_res(y)
                            if (P.layerwise) hn += eown[i * GB + p];               // + inputs[0] (wavenet_v2.py:283-284)
                            slice[o] = hn;
                        } else {
                            const int j = i - ly.nres;
                            sacc[j * GB + p] = (l == 0) ? v : (v + sacc[j * GB + p]);  // skips = conv_skip(y) + skips
                        }
                    }
                    __syncthreads();
                }
                // next-layer input h_{l+1}: own slice is in `slice` (residual and / or layerwise input) or is y itself
                const bool own_slice = ly.nres > 0 || P.layerwise;
                if (P.layerwise && ly.nres == 0) {
                    for (int o = tid; o < nf * GB; o += WN_NT) slice[o] = yc[(size_t)rank * nf * GB + o] + eown[o];
                    __syncthreads();
                }
                const float* hnext_slice = own_slice ? slice : (yc + (size_t)rank * nf * GB);
                if (last_layer && ns == 0 && own_slice && head_on) {
                    // no skips: the head reads h_L = y + inputs[0] (layerwise) or x + conv_res(y) (a conv_res on the layer executed
                    // last: reverse_layer_order, blocks=()): gather it where the skip sum would have gone
                    scatter_slice<CS>(cluster, hin + (size_t)rank * nf * GB, slice, nf * GB / 4);
                    cluster_sync_all<CS>();
                }
                if (!last_layer) {
                    if (!last_owned) {
                        if (own_slice) {
                            scatter_slice<CS>(cluster, x1 + (size_t)rank * nf * GB, hnext_slice, nf * GB / 4);
                            cluster_sync_all<CS>();   // #2
                        } else {
                            // h_{l+1} = y: already gathered in yc; every CTA copies it locally
                            for (int i = tid; i < (int)(blk / 4); i += WN_NT)
                                reinterpret_cast<float4*>(x1)[i] = reinterpret_cast<const float4*>(yc)[i];
                            __syncthreads();
                        }
                    } else {
                        // hand the group to the next stage through the mailbox
                        const int ns_ = stage + 1;
                        wait_u32(P.ack + ns_ * P.G + g, delivery * CS, P.abort_flag);
                        float* mh = P.mail_h + ((size_t)ns_ * P.G + g) * blk + (size_t)rank * nf * GB;
                        for (int i = tid; i < nf * GB / 4; i += WN_NT)
                            __stcg(reinterpret_cast<float4*>(mh) + i, reinterpret_cast<const float4*>(hnext_slice)[i]);
                        if (ns > 0) {
                            float* ms = P.mail_s + ((size_t)ns_ * P.G + g) * (size_t)S * GB + (size_t)rank * ns * GB;
                            for (int i = tid; i < ns * GB; i += WN_NT) __stcg(ms + i, sacc[i]);
                        }
                        __syncthreads();
                        if (tid == 0) { __threadfence(); red_release_add(P.ready + ns_ * P.G + g, 1u); }
                    }
                }
                par ^= 1u;
            }  // layers

            // ---------------- head + sampler (last stage) ----------------
            if (last_stage && head_on) {
                const unsigned pp = par ^ 1u;                    // buffers used by the last layer
                const float* ylast = ybuf + pp * blk;
                const float* hW1 = head_s;                          // [Kh][NH]
                const float* hb1 = hW1 + (size_t)P.Kh * P.NH;       // [NH]
                const float* hW2 = hb1 + P.NH;                      // [Hh][NZ]
                const float* hb2 = hW2 + (size_t)P.Hh * P.NZ;       // [NZ]
                const float* hWh = hb2 + P.NZ;                      // [Hh][NH] the shared hidden Linear (n_hh > 0)
                const float* hbh = hWh + (size_t)P.Hh * P.NH;       // [NH]
                const float* head_in;
                if (ns > 0) {
                    scatter_slice<CS>(cluster, hin + (size_t)rank * ns * GB, sacc, ns * GB / 4);
                    cluster_sync_all<CS>();
                    head_in = hin;
                } else {
                    // no skips: the head reads h_L of the last layer — y itself, or the gathered y + inputs[0] / x + conv_res(y)
                    head_in = (P.layerwise || P.layers[P.L - 1].nres > 0) ? hin : ylast;
                }
                {   // hidden = mish(W1 x + b1)
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const Tile tl = make_tile(P.NH);
                    gemm_accum(acc, tl, hW1, P.NH, head_in, P.Kh);
                    gemm_store_partials(acc, tl, part);
                    __syncthreads();
                    for (int o = tid; o < P.nh * GB; o += WN_NT) {
                        const int i = o / GB, p = o - i * GB;
                        slice[o] = mish_acc(gemm_reduce(part, P.NH, i, p) + hb1[i]);
                    }
                    __syncthreads();
                    scatter_slice<CS>(cluster, hid + (size_t)rank * P.nh * GB, slice, P.nh * GB / 4);
                }
                cluster_sync_all<CS>();
                const float* hid_in = hid;
                for (int r = 0; r < P.n_hh; ++r) {   // hidden layers: ONE Linear(Hh, Hh) + Mish applied n_hh times (mlp.py:47-50)
                    float* hid_out = hid + (size_t)((r + 1) & 1) * P.Hh * GB;
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const Tile tl = make_tile(P.NH);
                    gemm_accum(acc, tl, hWh, P.NH, hid_in, P.Hh);
                    gemm_store_partials(acc, tl, part);
                    __syncthreads();
                    for (int o = tid; o < P.nh * GB; o += WN_NT) {
                        const int i = o / GB, p = o - i * GB;
                        slice[o] = mish_acc(gemm_reduce(part, P.NH, i, p) + hbh[i]);
                    }
                    __syncthreads();
                    scatter_slice<CS>(cluster, hid_out + (size_t)rank * P.nh * GB, slice, P.nh * GB / 4);
                    cluster_sync_all<CS>();
                    hid_in = hid_out;
                }
                {   // z = W2 hidden + b2 : Q+1 values, this CTA's columns [rank*nz, ...)
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const Tile tl = make_tile(P.NZ);
                    gemm_accum(acc, tl, hW2, P.NZ, hid_in, P.Hh);
                    gemm_store_partials(acc, tl, part);
                    __syncthreads();
                    float* z0 = (CS == 1) ? zbuf : cluster.map_shared_rank(zbuf, 0);
                    const int c_lo = rank * P.nz;
                    const int c_n = min(P.nz, P.Q + 1 - c_lo);
                    for (int o = tid; o < c_n * GB; o += WN_NT) {
                        const int i = o / GB, p = o - i * GB;
                        z0[(size_t)p * P.zrow + c_lo + i] = gemm_reduce(part, P.NZ, i, p) + hb2[i];
                    }
                }
                cluster_sync_all<CS>();
                if (rank == 0 && warp < GB) {
                    // one warp per prompt: learned temperature, argmax / inverse-CDF sampling
                    const int p = warp, b = g * GB + p;
                    if (b < P.B) {
                        const int Q = P.Q;
                        float* z = zbuf + (size_t)p * P.zrow;
                        const float temp = fmaxf(sigmoid_acc(z[Q]), P.min_temp);        // mlp.py:60-62
                        const long long hstep = t - P.t_head;
                        float* lout = P.logits_out ? P.logits_out + ((size_t)b * (P.t_end - P.t_head) + hstep) * Q : nullptr;
                        float best = -INFINITY;
                        int besti = 0x7fffffff;
                        __syncwarp();
                        for (int c = lane; c < Q; c += 32) {
                            const float v = z[c] / temp;
                            z[c] = v;
                            if (lout) __stcs(lout + c, v);
                            if (v > best) { best = v; besti = c; }   // strided visit keeps the lowest index per lane
                        }
                        for (int o = 16; o > 0; o >>= 1) {
                            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                        }
                        int choice = besti;                          // targets.py:42-43 (first maximal index)
                        if (P.temperature != nullptr) {
                            // inverse-CDF draw, blocked-scan order (oracle/restate.py: sample_inverse_cdf)
                            const float T = P.temperature[P.n_temperature == 1 ? 0 : b];
                            const float u = P.noise[(size_t)b * P.noise_stride + (t + 1 - P.noise_t0)];
                            const float m = __fdiv_rn(best, T);      // max_k (z_k / T) = (max_k z_k) / T for T > 0
                            float mm = m;
                            if (!(T > 0.0f)) {                       // non-positive T: take the true maximum
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
                                    z[c] = run;                      // in-lane inclusive prefix
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
                            choice = min(Q - 1, cnt);
                        }
                        if (lane == 0) {
                            if (P.decisions) P.decisions[(size_t)b * (P.t_end - P.t_head) + hstep] = choice;
                            if (!P.teacher_forced) __stcg(P.seq + (size_t)b * P.seq_stride + t + 1, (long long)choice);
                        }
                    }
                }
                if (rank == 0) {
                    __syncthreads();
                    if (tid == 0 && !P.teacher_forced) { __threadfence(); st_release_s64(P.avail + g, t + 2); }
                }
            }
            if (last_stage && rank == 0 && tid == 0 && g == P.n_groups - 1 && P.step_ts)
                P.step_ts[t - P.t_begin] = globaltimer();
        }  // groups
    }      // time
    cp_async_wait_all();
    cluster_sync_all<CS>();
}

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
static int pad4(int v) { return (v + 3) / 4 * 4; }

}  // namespace mmk

using namespace mmk;

struct mmk_wavenet_s {
    wn7_handle* v7 = nullptr;   // set when the layer-pipelined tensor-core kernel (wavenet7.cu) hosts this network (compute_mode 1)
    wn6_handle* v6 = nullptr;   // set when the layer-pipelined kernel (wavenet6.cu) hosts this network (the default)
    wn4_handle* v4 = nullptr;   // set when the bf16 tensor-core kernel (wavenet_tc.cu) hosts this network (compute_mode 1)
    WnParams p{};
    int device = 0;
    int max_batch = 0;
    int rf = 0;
    size_t smem_bytes = 0;
    void* d_wpack = nullptr; void* d_hpack = nullptr; void* d_E = nullptr; void* d_rings = nullptr;
    void* d_mail_h = nullptr; void* d_mail_s = nullptr; void* d_flags = nullptr;
    size_t flags_bytes = 0;
    const void* kernel = nullptr;
};

template <int CS>
static int wn_max_clusters(size_t smem_bytes, int* out) {
    const void* k = (const void*)wavenet_pipe_kernel<CS>;
    MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (CS > 8) MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (CS == 1) {
        int per_sm = 0, dev = 0, sms = 0;
        MMK_CUDA(cudaGetDevice(&dev));
        MMK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MMK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, WN_NT, smem_bytes));
        *out = per_sm * sms;
        return 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS * 8);
    cfg.blockDim = dim3(WN_NT);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *out = n;
    return 0;
}

static const void* wn_kernel_ptr(int CS) {
    switch (CS) {
        case 1: return (const void*)wavenet_pipe_kernel<1>;
        case 2: return (const void*)wavenet_pipe_kernel<2>;
        case 4: return (const void*)wavenet_pipe_kernel<4>;
        case 8: return (const void*)wavenet_pipe_kernel<8>;
        case 16: return (const void*)wavenet_pipe_kernel<16>;
    }
    return nullptr;
}
static int wn_query_clusters(int CS, size_t smem, int* out) {
    switch (CS) {
        case 1: return wn_max_clusters<1>(smem, out);
        case 2: return wn_max_clusters<2>(smem, out);
        case 4: return wn_max_clusters<4>(smem, out);
        case 8: return wn_max_clusters<8>(smem, out);
        case 16: return wn_max_clusters<16>(smem, out);
    }
    MMK_FAIL("bad cluster size");
}

// fills the geometry + smem carve-up for (CS, layers_per_stage); returns the dynamic smem bytes
static size_t wn_plan(WnParams& p, int CS, int max_layers_per_stage, bool has_head) {
    p.CS = CS;
    p.nf = p.C / CS; p.ns = p.S / CS; p.nh = p.Hh / CS; p.nz = (p.Q + 1 + CS - 1) / CS;
    p.NA = pad4(2 * p.nf); p.NB = std::max(4, pad4(p.nf + p.ns)); p.NH = pad4(p.nh); p.NZ = pad4(p.nz);
    if (p.kmax < 2) p.kmax = 2;
    p.NC = p.affine ? pad4(3 * p.nf) : 0;
    p.layer_block = pad4(p.kmax * p.C * p.NA + p.NA + p.C * p.NB + p.NB + p.C * p.NC + p.NC);
    p.head_block = pad4(p.Kh * p.NH + p.NH + p.Hh * p.NZ + p.NZ + (p.n_hh > 0 ? p.Hh * p.NH + p.NH : 0));
    const int n = (p.Q + 31) / 32;
    p.zrow = pad4(p.Q + 1 + 4);
    (void)n;
    int o = 0;
    auto take = [&](int floats) { int r = o; o += pad4(floats); return r; };
    p.off_w = take(max_layers_per_stage * p.layer_block);
    p.off_head = take(has_head ? p.head_block : 4);
    p.off_x1 = take(p.C * WN_GB);
    p.off_x0 = take(2 * (p.kmax - 1) * p.C * WN_GB);
    p.off_y = take(2 * p.C * WN_GB);
    p.off_sacc = take(std::max(4, p.ns * WN_GB));
    p.off_hin = take(std::max(4, p.Kh * WN_GB));
    p.off_hid = take(2 * p.Hh * WN_GB);
    p.off_z = take(WN_GB * p.zrow);
    p.off_part = take(WN_NT * 8);
    p.off_slice = take(std::max(std::max(p.nf, p.nh), p.nf + p.ns) * WN_GB + 8);
    p.off_e = take(std::max(4, p.nf * WN_GB));
    p.smem_floats = o;
    return (size_t)o * sizeof(float);
}

extern "C" int mmk_wavenet_create(const mmk_wavenet_desc* d, int max_batch, mmk_wavenet_t* out) {
    return mmk_wavenet_create_ex(d, max_batch, MMK_COMPUTE_FP32, out);
}

extern "C" int mmk_wavenet_create_ex(const mmk_wavenet_desc* d, int max_batch, int compute_mode, mmk_wavenet_t* out) {
    MMK_CHECK(d, "mmk_wavenet_create: null argument");
    mmk_wavenet_desc_ex dx{};
    dx.base = *d;
    return mmk_wavenet_create_cfg(&dx, max_batch, compute_mode, out);
}

extern "C" int mmk_wavenet_destroy(mmk_wavenet_t h);

extern "C" int mmk_wavenet_create_cfg(const mmk_wavenet_desc_ex* dx, int max_batch, int compute_mode, mmk_wavenet_t* out) {
    MMK_CHECK(dx && out, "mmk_wavenet_create: null argument");
    const mmk_wavenet_desc* d = &dx->base;
    MMK_CHECK(dx->head_hidden_layers >= 0 && dx->head_hidden_layers <= 8, "head_hidden_layers must be in [0, 8]");
    MMK_CHECK(dx->head_hidden_layers == 0 || (dx->head_wh && dx->head_bh), "missing head hidden-layer weights");
    int kmax = 2;
    MMK_CHECK((dx->aff_res_w != nullptr) == (dx->aff_res_b != nullptr), "aff_res_w and aff_res_b come together");
    MMK_CHECK(dx->act_f >= MMK_ACT_DEFAULT && dx->act_f <= MMK_ACT_COS && dx->act_g >= MMK_ACT_DEFAULT && dx->act_g <= MMK_ACT_COS,
              "act_f / act_g must be MMK_ACT_* codes");
    const int act_f = dx->act_f == MMK_ACT_DEFAULT ? MMK_ACT_TANH : dx->act_f;
    const int act_g = dx->act_g == MMK_ACT_DEFAULT ? MMK_ACT_SIGMOID : dx->act_g;
    bool plain = dx->layerwise_inputs == 0 && dx->head_hidden_layers == 0 && !dx->aff_res_w && act_f == MMK_ACT_TANH &&
                 act_g == MMK_ACT_SIGMOID;
    if (dx->kernel_sizes)
        for (int l = 0; l < d->n_layers; ++l) {
            MMK_CHECK(dx->kernel_sizes[l] >= 2 && dx->kernel_sizes[l] <= WN_MAX_K, "kernel sizes must be in [2, 4]");
            kmax = std::max(kmax, dx->kernel_sizes[l]);
            plain = plain && dx->kernel_sizes[l] == 2;
        }
    MMK_CHECK(plain || compute_mode == MMK_COMPUTE_FP32,
              "kernel sizes > 2, layerwise_inputs, hidden MLP layers, affine residuals and other activations run in the fp32 general kernel only");
    MMK_CHECK(compute_mode == MMK_COMPUTE_FP32 || compute_mode == MMK_COMPUTE_BF16_TC, "unknown compute_mode");
    MMK_CHECK(d->n_layers >= 1 && d->n_layers <= WN_MAX_LAYERS, "n_layers out of range [1, 96]");
    MMK_CHECK(d->dilated_dim >= 4 && d->dilated_dim % 4 == 0, "dilated_dim must be a positive multiple of 4");
    MMK_CHECK(d->skips_dim >= 0 && d->skips_dim % 4 == 0, "skips_dim must be 0 or a multiple of 4");
    MMK_CHECK(d->head_hidden >= 4 && d->head_hidden % 4 == 0, "head_hidden must be a positive multiple of 4");
    MMK_CHECK(d->q_levels >= 2 && d->q_levels <= 1024, "q_levels must be in [2, 1024]");
    MMK_CHECK(max_batch >= 1, "max_batch must be >= 1");
    MMK_CHECK(d->dilations && d->embedding && d->conv_dil_w && d->conv_dil_b && d->conv_res_w && d->conv_res_b &&
              d->head_w1 && d->head_b1 && d->head_w2 && d->head_b2, "missing weight pointers");
    MMK_CHECK(d->skips_dim == 0 || (d->conv_skip_w && d->conv_skip_b), "skips_dim > 0 needs conv_skip weights");
    // a conv_res on the layer executed last (reverse_layer_order, blocks=(): wavenet_v2.py:216, 270): without skips the head reads
    // h_L = x + conv_res(y) (:286-292); only the general kernel hosts it
    const bool last_res = d->conv_res_w[d->n_layers - 1] != nullptr;
    MMK_CHECK(!last_res || compute_mode == MMK_COMPUTE_FP32, "a residual conv on the last layer runs in the fp32 general kernel only");
    int ndev = 0;
    MMK_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device: mmk_b200 has no CPU fallback");

    auto* h = new mmk_wavenet_s();
    struct Guard {                 // every early return below (MMK_CHECK / MMK_CUDA / MMK_FAIL) releases the handle and what it owns
        mmk_wavenet_s* h;
        ~Guard() { if (h) mmk_wavenet_destroy(h); }
    } guard{h};
    WnParams& p = h->p;
    MMK_CUDA(cudaGetDevice(&h->device));
    if (compute_mode == MMK_COMPUTE_BF16_TC) {
        // explicit request: no silent fall-back to another precision
        int unsupported = 0;
        for (int l = 0; l < d->n_layers; ++l)
            if (!d->conv_dil_w[l] || !d->conv_dil_b[l]) { MMK_FAIL("missing conv_dil weights"); }
        {
            int unsup7 = 0;
            if (wn7_create(d, max_batch, &h->v7, &unsup7) == 0) {
                int rf7 = 1;
                for (int l = 0; l < d->n_layers; ++l) rf7 += d->dilations[l];
                h->rf = rf7; h->max_batch = max_batch;
                guard.h = nullptr;
                *out = h;
                return 0;
            }
            h->v7 = nullptr;
            if (!unsup7) { return 1; }
        }
        if (wn4_create(d, max_batch, &h->v4, &unsupported) != 0) { return 1; }
        int rf4 = 1;
        for (int l = 0; l < d->n_layers; ++l) rf4 += d->dilations[l];
        h->rf = rf4; h->max_batch = max_batch;
        guard.h = nullptr;
        *out = h;
        return 0;
    }
    {
        const char* force = (plain && !last_res) ? getenv("MMK_WN_KERNEL") : "1";   // "1" = general kernel only, "6" = layer-pipelined kernel only
        if (!force || atoi(force) == 6) {
            int unsupported = 0;
            if (wn6_create(d, max_batch, &h->v6, &unsupported) == 0) {
                int rf6 = 1;
                for (int l = 0; l < d->n_layers; ++l) rf6 += d->dilations[l];
                h->rf = rf6; h->max_batch = max_batch;
                guard.h = nullptr;
                *out = h;
                return 0;
            }
            h->v6 = nullptr;
            if (!unsupported) { return 1; }
            if (force) { MMK_FAIL("configuration not supported by the layer-pipelined fp32 kernel (MMK_WN_KERNEL=6)"); }
        }
    }
    p.L = d->n_layers; p.C = d->dilated_dim; p.S = d->skips_dim; p.Hh = d->head_hidden; p.Q = d->q_levels;
    p.Kh = p.S > 0 ? p.S : p.C;
    p.kmax = kmax; p.layerwise = dx->layerwise_inputs ? 1 : 0; p.n_hh = dx->head_hidden_layers;
    p.affine = dx->aff_res_w ? 1 : 0;
    p.act_f = act_f; p.act_g = act_g;
    p.min_temp = d->min_temperature;
    h->max_batch = max_batch;
    p.G = (max_batch + WN_GB - 1) / WN_GB;
    int rf = 1;
    for (int l = 0; l < p.L; ++l) {
        MMK_CHECK(d->dilations[l] >= 1, "dilation must be >= 1");
        rf += d->dilations[l] * ((dx->kernel_sizes ? dx->kernel_sizes[l] : 2) - 1);
        MMK_CHECK(d->conv_dil_w[l] && d->conv_dil_b[l], "missing conv_dil weights");
        MMK_CHECK(!p.affine || (dx->aff_res_w[l] && dx->aff_res_b[l]), "missing aff_res weights");
    }
    h->rf = rf;

    int dev_sms = 0, max_optin = 0;
    MMK_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device));
    MMK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));

    // choose the cluster size / number of stages: the largest cluster the dims divide into, as many stages as are
    // co-resident and useful; every stage's weights must fit in shared memory.
    int best_cs = 0, best_nst = 0;
    size_t best_smem = 0;
    const char* force_cs = getenv("MMK_WN_CLUSTER");
    const char* force_nst = getenv("MMK_WN_STAGES");
    for (int CS : {16, 8, 4, 2, 1}) {
        if (force_cs && atoi(force_cs) != CS) continue;
        if (p.C % CS || p.S % CS || p.Hh % CS) continue;
        {   // every contraction must fit one pass of (GB/2) x (columns/4) register tiles over the CTA's threads
            WnParams q = p;
            wn_plan(q, CS, 1, true);
            const int widest = std::max(std::max(std::max(q.NA, q.NB), std::max(q.NH, q.NZ)), q.NC);
            if ((WN_GB / 2) * (widest / 4) > WN_NT) continue;
        }
        // smallest number of stages for which the weights fit
        int nst_min = 0;
        for (int nst = 1; nst <= std::min(p.L, WN_MAX_STAGES); ++nst) {
            WnParams q = p;
            size_t smem = wn_plan(q, CS, (p.L + nst - 1) / nst, true);
            if (smem <= (size_t)max_optin) { nst_min = nst; break; }
        }
        if (!nst_min) continue;
        WnParams q = p;
        size_t smem_min = wn_plan(q, CS, (p.L + nst_min - 1) / nst_min, true);
        int max_clusters = 0;
        if (wn_query_clusters(CS, smem_min, &max_clusters)) { return 1; }
        if (max_clusters < nst_min) continue;
        int nst = std::min(std::min(max_clusters, p.L), WN_MAX_STAGES);
        if (force_nst) nst = std::max(nst_min, std::min(nst, atoi(force_nst)));
        else nst = std::min(nst, std::max(nst_min, std::max(1, p.G)));  // more stages than groups in flight only adds latency
        q = p;
        size_t smem = wn_plan(q, CS, (p.L + nst - 1) / nst, true);
        best_cs = CS; best_nst = nst; best_smem = smem;
        break;
    }
    if (!best_cs) { MMK_FAIL("WaveNet configuration does not fit the persistent kernel (shared memory / cluster limits)"); }
    const int CS = best_cs, NST = best_nst;
    p.NST = NST;
    const int per = (p.L + NST - 1) / NST;
    h->smem_bytes = wn_plan(p, CS, per, true);
    (void)best_smem;
    // balanced contiguous layer ranges; the last stage (which also runs the head) gets the short end
    {
        int base = p.L / NST, extra = p.L % NST, lo = 0;
        for (int s = 0; s < NST; ++s) { p.stage_lo[s] = lo; lo += base + (s < extra ? 1 : 0); }
        p.stage_lo[NST] = p.L;
    }
    h->kernel = wn_kernel_ptr(CS);
    MMK_CUDA(cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    if (CS > 8) MMK_CUDA(cudaFuncSetAttribute(h->kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    // ---- pack weights: per (layer, rank) block = W1[kmax C][NA] (tap j at rows [j C, (j+1) C)) | b1[NA] | W2[C][NB] | b2[NB]
    const int C = p.C, S = p.S, nf = p.nf, ns = p.ns;
    std::vector<float> wpack((size_t)p.L * CS * p.layer_block, 0.0f);
    long long ring_off = 0;
    for (int l = 0; l < p.L; ++l) {
        const bool has_res = d->conv_res_w[l] != nullptr;
        const int ks = dx->kernel_sizes ? dx->kernel_sizes[l] : 2;
        p.layers[l].dilation = d->dilations[l];
        p.layers[l].nres = has_res ? nf : 0;
        p.layers[l].ksize = ks;
        p.layers[l].ring_off = ring_off;
        ring_off += (long long)(ks - 1) * d->dilations[l] * p.G * C * WN_GB;
        const float* wd = d->conv_dil_w[l];   // (2C, C, ks): [o][c][tap], tap 0 = oldest sample
        const float* bd = d->conv_dil_b[l];
        for (int r = 0; r < CS; ++r) {
            float* blk = wpack.data() + ((size_t)l * CS + r) * p.layer_block;
            float* W1 = blk; float* b1 = W1 + (size_t)p.kmax * C * p.NA; float* W2 = b1 + p.NA; float* b2 = W2 + (size_t)C * p.NB;
            for (int j = 0; j < 2 * nf; ++j) {
                const int o = (j < nf) ? (r * nf + j) : (C + r * nf + (j - nf));   // f channels first, then g
                for (int c = 0; c < C; ++c)
                    for (int tap = 0; tap < ks; ++tap)
                        W1[(size_t)(tap * C + c) * p.NA + j] = wd[((size_t)o * C + c) * ks + tap];
                b1[j] = bd[o];
            }
            int col = 0;
            if (has_res) {
                for (int j = 0; j < nf; ++j, ++col) {
                    const int o = r * nf + j;
                    for (int c = 0; c < C; ++c) W2[(size_t)c * p.NB + col] = d->conv_res_w[l][(size_t)o * C + c];
                    b2[col] = d->conv_res_b[l][o];
                }
            }
            for (int j = 0; j < ns; ++j, ++col) {
                const int o = r * ns + j;
                for (int c = 0; c < C; ++c) W2[(size_t)c * p.NB + col] = d->conv_skip_w[l][(size_t)o * C + c];
                b2[col] = d->conv_skip_b[l][o];
            }
            if (p.affine) {   // aff_res.params (3C, C, 1): this CTA's channels of x_hat, then of a, then of b
                float* W3 = b2 + p.NB; float* b3 = W3 + (size_t)C * p.NC;
                for (int j = 0; j < 3 * nf; ++j) {
                    const int o = (j / nf) * C + r * nf + (j % nf);
                    for (int c = 0; c < C; ++c) W3[(size_t)c * p.NC + j] = dx->aff_res_w[l][(size_t)o * C + c];
                    b3[j] = dx->aff_res_b[l][o];
                }
            }
        }
    }
    std::vector<float> hpack((size_t)CS * p.head_block, 0.0f);
    for (int r = 0; r < CS; ++r) {
        float* blk = hpack.data() + (size_t)r * p.head_block;
        float* W1 = blk; float* b1 = W1 + (size_t)p.Kh * p.NH; float* W2 = b1 + p.NH; float* b2 = W2 + (size_t)p.Hh * p.NZ;
        for (int j = 0; j < p.nh; ++j) {
            const int o = r * p.nh + j;
            for (int c = 0; c < p.Kh; ++c) W1[(size_t)c * p.NH + j] = d->head_w1[(size_t)o * p.Kh + c];
            b1[j] = d->head_b1[o];
        }
        if (p.n_hh > 0) {
            float* Wh = b2 + p.NZ; float* bh = Wh + (size_t)p.Hh * p.NH;
            for (int j = 0; j < p.nh; ++j) {
                const int o = r * p.nh + j;
                for (int c = 0; c < p.Hh; ++c) Wh[(size_t)c * p.NH + j] = dx->head_wh[(size_t)o * p.Hh + c];
                bh[j] = dx->head_bh[o];
            }
        }
        for (int j = 0; j < p.nz; ++j) {
            const int o = r * p.nz + j;
            if (o > p.Q) break;
            for (int c = 0; c < p.Hh; ++c) W2[(size_t)c * p.NZ + j] = d->head_w2[(size_t)o * p.Hh + c];
            b2[j] = d->head_b2[o];
        }
    }
    const size_t ring_floats = (size_t)ring_off;
    const size_t mail_h_floats = (size_t)NST * p.G * C * WN_GB, mail_s_floats = (size_t)NST * p.G * std::max(S, 1) * WN_GB;
    h->flags_bytes = sizeof(unsigned) * (2 * (size_t)NST * p.G) + sizeof(long long) * (size_t)p.G + 64;
    MMK_CUDA(cudaMalloc(&h->d_wpack, wpack.size() * sizeof(float)));
    MMK_CUDA(cudaMalloc(&h->d_hpack, hpack.size() * sizeof(float)));
    MMK_CUDA(cudaMalloc(&h->d_E, (size_t)p.Q * C * sizeof(float)));
    MMK_CUDA(cudaMalloc(&h->d_rings, ring_floats * sizeof(float)));
    MMK_CUDA(cudaMalloc(&h->d_mail_h, mail_h_floats * sizeof(float)));
    MMK_CUDA(cudaMalloc(&h->d_mail_s, mail_s_floats * sizeof(float)));
    MMK_CUDA(cudaMalloc(&h->d_flags, h->flags_bytes));
    MMK_CUDA(cudaMemcpy(h->d_wpack, wpack.data(), wpack.size() * sizeof(float), cudaMemcpyHostToDevice));
    MMK_CUDA(cudaMemcpy(h->d_hpack, hpack.data(), hpack.size() * sizeof(float), cudaMemcpyHostToDevice));
    MMK_CUDA(cudaMemcpy(h->d_E, d->embedding, (size_t)p.Q * C * sizeof(float), cudaMemcpyHostToDevice));
    MMK_CUDA(cudaMemset(h->d_rings, 0, ring_floats * sizeof(float)));
    MMK_CUDA(cudaMemset(h->d_mail_h, 0, mail_h_floats * sizeof(float)));
    MMK_CUDA(cudaMemset(h->d_mail_s, 0, mail_s_floats * sizeof(float)));
    p.wpack = (const float*)h->d_wpack; p.hpack = (const float*)h->d_hpack; p.E = (const float*)h->d_E;
    p.rings = (float*)h->d_rings; p.mail_h = (float*)h->d_mail_h; p.mail_s = (float*)h->d_mail_s;
    // flags: [avail: G x s64][ready: NST*G x u32][ack: NST*G x u32][abort: u32]
    p.avail = (long long*)h->d_flags;
    p.ready = (unsigned*)(p.avail + p.G);
    p.ack = p.ready + (size_t)NST * p.G;
    p.abort_flag = p.ack + (size_t)NST * p.G;
    guard.h = nullptr;
    *out = h;
    return 0;
}

extern "C" int mmk_wavenet_destroy(mmk_wavenet_t h) {
    if (!h) return 0;
    if (h->v4) wn4_destroy(h->v4);
    if (h->v7) wn7_destroy(h->v7);
    if (h->v6) wn6_destroy(h->v6);
    cudaFree(h->d_wpack); cudaFree(h->d_hpack); cudaFree(h->d_E); cudaFree(h->d_rings);
    cudaFree(h->d_mail_h); cudaFree(h->d_mail_s); cudaFree(h->d_flags);
    delete h;
    return 0;
}

extern "C" int mmk_wavenet_sync_check(mmk_wavenet_t h, void* stream) {
    MMK_CHECK(h, "null handle");
    if (h->v4) return wn4_sync_check(h->v4, stream);
    if (h->v7) return wn7_sync_check(h->v7, stream);
    if (h->v6) return wn6_sync_check(h->v6, stream);
    unsigned aborted = 0;
    MMK_CUDA(cudaMemcpyAsync(&aborted, h->p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MMK_CHECK(aborted == 0, "WaveNet kernel watchdog fired: an inter-stage wait timed out (results invalid)");
    return 0;
}

extern "C" int mmk_wavenet_rf(mmk_wavenet_t h) { return h ? h->rf : -1; }

extern "C" int mmk_wavenet_launch_info(mmk_wavenet_t h, mmk_launch_info* out) {
    MMK_CHECK(h && out, "null argument");
    if (h->v4) return wn4_launch_info(h->v4, out);
    if (h->v7) return wn7_launch_info(h->v7, out);
    if (h->v6) return wn6_launch_info(h->v6, out);
    out->cluster_size = h->p.CS; out->n_stages = h->p.NST; out->group_size = WN_GB; out->threads = WN_NT;
    out->smem_bytes = (int)h->smem_bytes; out->sm_used = h->p.CS * h->p.NST;
    return 0;
}

__global__ void wn_init_flags_kernel(long long* avail, int G, long long avail0, unsigned* u32s, int n_u32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G) avail[i] = avail0;
    if (i < n_u32) u32s[i] = 0u;
}

extern "C" int mmk_wavenet_run(mmk_wavenet_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0,
                               int64_t t_begin, int64_t t_head, int64_t t_end, int teacher_forced, const float* d_temperature,
                               int n_temperature, const float* d_noise, int64_t noise_stride, int64_t noise_t0,
                               float* d_logits_out, int64_t* d_decisions, unsigned long long* d_step_ts, void* stream) {
    MMK_CHECK(h && d_seq, "mmk_wavenet_run: null argument");
    MMK_CHECK(B >= 1 && B <= h->max_batch, "batch exceeds the max_batch the handle was created for");
    MMK_CHECK(0 <= seq_t0 && seq_t0 <= t_begin && t_begin <= t_end && t_head >= t_begin,
              "need 0 <= seq_t0 <= t_begin <= t_head and t_begin <= t_end");
    MMK_CHECK(teacher_forced ? (t_end - seq_t0 <= seq_stride) : (t_end + 1 - seq_t0 <= seq_stride || t_head >= t_end),
              "sequence buffer too short for the requested steps");
    MMK_CHECK(d_temperature == nullptr || (n_temperature == 1 || n_temperature == B), "temperature must have 1 or B entries");
    MMK_CHECK(d_temperature == nullptr || d_noise != nullptr, "sampling (temperature given) needs a noise tensor");
    if (t_begin == t_end) return 0;
    if (h->v4)
        return wn4_run(h->v4, d_seq, B, seq_stride, seq_t0, t_begin, t_head, t_end, teacher_forced, d_temperature,
                       n_temperature, d_noise, noise_stride, noise_t0, d_logits_out, d_decisions, d_step_ts, stream);
    if (h->v7)
        return wn7_run(h->v7, d_seq, B, seq_stride, seq_t0, t_begin, t_head, t_end, teacher_forced, d_temperature,
                       n_temperature, d_noise, noise_stride, noise_t0, d_logits_out, d_decisions, d_step_ts, stream);
    if (h->v6)
        return wn6_run(h->v6, d_seq, B, seq_stride, seq_t0, t_begin, t_head, t_end, teacher_forced, d_temperature,
                       n_temperature, d_noise, noise_stride, noise_t0, d_logits_out, d_decisions, d_step_ts, stream);
    cudaStream_t st = (cudaStream_t)stream;
    WnParams p = h->p;
    p.seq = reinterpret_cast<long long*>(d_seq) - seq_t0;   // column j of d_seq holds time seq_t0 + j
    p.seq_stride = seq_stride; p.t_begin = t_begin; p.t_head = t_head; p.t_end = t_end;
    p.B = B; p.n_groups = (B + WN_GB - 1) / WN_GB; p.teacher_forced = teacher_forced ? 1 : 0;
    p.temperature = d_temperature; p.n_temperature = n_temperature;
    p.noise = d_noise; p.noise_stride = noise_stride; p.noise_t0 = noise_t0;
    p.logits_out = d_logits_out; p.decisions = reinterpret_cast<long long*>(d_decisions); p.step_ts = d_step_ts;
    const int n_u32 = 2 * p.NST * p.G + 1;
    const long long avail0 = teacher_forced ? (t_end + 1) : (t_head + 1);
    wn_init_flags_kernel<<<(std::max(p.G, n_u32) + 255) / 256, 256, 0, st>>>(p.avail, p.G, avail0, p.ready, n_u32);
    MMK_CUDA(cudaGetLastError());

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.CS * p.NST);
    cfg.blockDim = dim3(WN_NT);
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = p.CS > 1 ? 1 : 0;
    void* args[] = {&p};
    MMK_CUDA(cudaLaunchKernelExC(&cfg, h->kernel, args));
    return 0;
}

extern "C" int mmk_wavenet_generate(mmk_wavenet_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t prompt_len,
                                    int64_t n_steps, const float* d_temperature, int n_temperature,
                                    const float* d_noise, float* d_logits_out, unsigned long long* d_step_ts,
                                    void* stream) {
    MMK_CHECK(h, "null handle");
    MMK_CHECK(prompt_len >= h->rf, "prompt shorter than the receptive field (the reference raises RuntimeError too)");
    MMK_CHECK(n_steps >= 0 && prompt_len + n_steps <= seq_stride, "sequence buffer too short");
    if (n_steps == 0) return 0;
    return mmk_wavenet_run(h, d_seq, B, seq_stride, 0, prompt_len - h->rf, prompt_len - 1, prompt_len + n_steps - 1, 0,
                           d_temperature, n_temperature, d_noise, n_steps, prompt_len, d_logits_out, nullptr,
                           d_step_ts, stream);
}
