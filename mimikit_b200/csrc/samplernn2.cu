// Cluster SampleRNN generation kernel (sm_100a) — the fast path behind mmk_samplernn_*.  Same function as
// samplernn.cu (the general kernel; see that file for the reference lines: sample_rnn_v2.py:83-119, 226-260;
// modules/io.py:106-133, 185-198; modules/resamplers.py:13-23; networks/mlp.py:44-63; modules/targets.py:40-52).
//
// One persistent launch covers before_generate's warm-up over the prompt and every generated sample.
//   * the sample-level tier + MLP head + sampler never touch a grid barrier.  A thread-block cluster of CS CTAs owns a
//     group of prompts end to end: W1 and W2 are split along K over the cluster (a CTA needs only its H/CS rows of
//     the conditioning vector), partial sums meet through distributed shared memory (st.async + mbarrier
//     complete_tx): partial hidden -> reduce-scatter by row, Mish, partial logits -> reduce-scatter by prompt, every
//     thread scales one logit, one warp per prompt samples, the index is broadcast to the cluster.
//   * frame tiers stay weight-stationary over all CTAs (recurrent rows split by hidden index, up-sampler rows evenly), with
//     grid barriers only around tier firings.  Three engines (template parameter ENGINE of the kernel):
//       0  tile engine (round 1): activations [H][B] streamed through cp.async stages into 4x4 register tiles, weights in
//          shared memory.  Any hidden size that splits over the CTAs; GRU, zero initial state.
//       1  lane-major fp32 engine: activations [prompt][H] prefetched into registers, the lane's weights in registers for a
//          whole firing, transposing shuffle trees.  H in {128, 256, 512}; GRU and LSTM; bit-exact with the oracle.
//       2  tcgen05 engine (compute mode bf16): one tensor-core accumulation per contraction, activations kept as UMMA tiles,
//          frame Linear / head / conditioning folded into precomputed products.  H in {128, 256, 512}, <= 128 prompts.
#include "common.cuh"
#include "sampler.cuh"
#include "samplernn_impl.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

namespace mmk_sr2 {

using mmk::decide_warp;
using mmk::mish_acc;
using mmk::sigmoid_acc;

constexpr int NT = 512;          // threads per CTA
constexpr int HT = 256;          // threads that take tiles in the head contractions
constexpr int KC = 16;           // activation rows per streamed chunk at full width (x 2, 4, 8 for narrower batches)
constexpr int NSTAGE_DEFAULT = 2; // full-width cp.async stages (more stages when the chunks are small)
constexpr int KCFULL_DEFAULT = 64; // rows per full-width chunk (measured optimum of the tile engine)
constexpr int PBW = 128;         // prompts per streamed block
constexpr int WMAX = 32;         // widest streamed weight slice (columns per CTA)
constexpr int MAX_TIERS = 6;
// floats per stage (activations, streamed weights) = kcfull * PBW, kcfull * WMAX
constexpr int MAXST = 8;
constexpr unsigned long long WAIT_LIMIT_NS = 4000000000ull;

struct Tier {
    int fs, up, kdiv, NU, up_rows;
    int off_wih, off_whh, off_wup;   // float offsets in the CTA's packed global block: W[H][N] followed by bias[N]
    int off_inw;                     // frame Linear: in_w (H, fs) followed by in_b (H)
    int so_wih, so_whh, so_wup;      // float offsets in shared memory, or -1: streamed from L2 with the activations
    int so_inw;
    const float* in_w; const float* in_b;
    float* hbuf; float* obuf;
    // lane-major engine (Params::fast): weights packed per lane, see pack_fast() — GRU [NC][24][H/4] float4, up-sampler
    // [NC][NV][H/4] float4, frame Linear [fs][H/4] float4 + bias [H/4] float4, gate biases [NC][24], up-sampler biases [NC][NV]
    const float4* wg4; const float4* wu4; const float4* iw4; const float4* ib4;
    const float* gb; const float* ub;
    int NV;                          // up-sampler rows of a CTA (4 * up)
    // tensor-core engine (Params::tc): bf16 images in the UMMA K-major SWIZZLE_128B layout — [128 prompts x H] activations
    // (16 KB atoms of 64 k), [16 columns x K] weight tiles per CTA
    float* cbuf;                     // LSTM cell state, fp32 [2][prompt][H] (ping-pong like hbuf); null for GRU tiers
    unsigned char* himg;             // hidden state, two images (ping-pong like hbuf)
    unsigned char* oimg;             // up-sampler output (the next tier's conditioning), one image per slot; null for the bottom frame tier
    const unsigned char* wgimg;      // [NC][fills][4 KB]: r | z | n_i | n_h columns over K = [conditioning | hidden]
    const unsigned char* wuimg;      // [NC][H / 128][4 KB]
    const unsigned char* wximg;      // fold_gx: [NC][H / 128][up * 4 KB] tiles of W_ih(next tier) . W_up(slot), up * 16 columns
    const float* xb;                 // fold_gx: [NC][up][16] = W_ih(next tier) . b_up(slot)
    float* gx;                       // fold_gx: [up][128 prompts][NC][16] fp32, written and read by the same CTA
    const float* wf;                 // [NC][16][fs]: W_ih . in_w rows of the columns (frame Linear folded into the gate)
    const float* bfold;              // [NC][16]: b_ih + b_hh + W_ih . in_b (n_i: without b_hh; n_h: b_hh alone)
};

struct Params {
    int n_ft, H, Hh, Q, NC, CS, GP, JP, NG, fs_last, fs_lt;
    int n_res, res_goff[3 * MAX_TIERS + 3], res_soff[3 * MAX_TIERS + 3], res_len[3 * MAX_TIERS + 3];   // resident pieces
    int KS, RS, ZR;                  // head: x rows per CTA (H/CS), hidden rows per CTA (Hh/CS), padded logit rows
    int off_w1, off_b1, off_w2, cta_block;     // global block offsets
    int so_w1, so_b1, so_w2;
    int s_region, s_gi, s_bar, s_b2, smem_floats;
    int xstage, wstage, xregion, wregion, region;    // per stage; floats: activation stages, streamed-weight stages, both (= partial sums / head buffers)
    int h_x, h_inh, h_hid, h_un, h_zs;      // head buffers inside the region (float offsets)
    const float* wpack; const float* conv_w; const float* conv_b; const float* b2;
    unsigned long long* bar; unsigned* abort_flag;
    float min_temp;
    Tier tiers[MAX_TIERS];
    // this run
    int B, Bp, teacher_forced, n_temperature;
    int hsel[MAX_TIERS];
    long long* seq;
    long long seq_stride, warm_begin, warm_end, warm_off, gen_begin, gen_end;
    const float* temperature; const float* noise;
    long long noise_stride, noise_t0;
    float* logits_out; long long* decisions; unsigned long long* step_ts;
    unsigned long long* dbg;         // MMK_SR_DEBUG: time (ns) spent by CTA 0 per section
    int exp;                         // MMK_SR_EXP: timing experiments (results invalid)
    // lane-major frame-tier engine
    int fast, NKQ, CHP;              // K quarters (H / 128), prompts per block of the walk (2 per warp)
    int s_part, s_hold, s_lin;
    int tc, s_tcbar;                 // tensor-core frame tiers (bf16 operands, fp32 accumulation): compute mode MMK_COMPUTE_BF16_TC
    // tensor-core mode, fold_head: the bottom tier's up-sampler emits W1 (up(h) + conv_b) + b1 directly (W1 . W_up precomputed), so a
    // head step starts from `pre` [slot][128 prompts][Hh] and adds (W1 conv_w) lin(q): no x rows, no W1 contraction, no exchange
    int lstm;                        // tensor-core mode only: nn.LSTM tiers (gates i, f, g, o; the reference's default rnn_class)
    int fold_head, NVh;              // NVh: folded up-sampler columns per CTA (up * Hh / NC)
    // tensor-core mode, fold_gx: a frame tier's up-sampler emits, per slot, the INPUT-side gate pre-activations of the tier below for
    // this CTA's own 16 columns (W_ih . (W_up h + b_up), products precomputed): the tier below then contracts over its hidden
    // state only (K halves), and no conditioning image is written or pulled.  tc_bbytes: B area of a stage (4 KB, or 16 KB with fold_gx)
    int fold_gx, tc_bbytes;
    float* pre; const float* hu; const float* hb;     // hu [Hh][fs_last] = W1 . conv_w; hb [NC][NVh] folded biases
};

// ------------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void st_async_v4(unsigned raddr, float4 v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_u32(unsigned raddr, unsigned v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];"
                 ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// mbarrier wait with the watchdog out of line
__device__ __noinline__ bool mbar_wait_slow(unsigned bar, unsigned parity, unsigned* abort_flag) {
    const unsigned long long t0 = globaltimer();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 63u) == 0u) {
            if (ld_relaxed_u32(abort_flag) != 0u) return false;
            if (globaltimer() - t0 > WAIT_LIMIT_NS) { atomicExch(abort_flag, 1u); return false; }
        }
    }
    return true;
}
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity, unsigned* abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    return mbar_wait_slow(bar, parity, abort_flag);
}

// All CTAs of the grid meet here (co-resident by construction).  Returns false when the launch was aborted.
__device__ __forceinline__ bool grid_barrier(const Params& P, unsigned long long& epoch) {
    __shared__ int ok_s;
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += (unsigned long long)P.NC;
        __threadfence();
        red_release_add_u64(P.bar, 1ull);
        int ok = 1;
        if (ld_acquire_u64(P.bar) < epoch) {
            const unsigned long long t0 = globaltimer();
            unsigned spins = 0;
            while (ld_acquire_u64(P.bar) < epoch) {
                if ((++spins & 63u) == 0u) {
                    if (ld_relaxed_u32(P.abort_flag) != 0u) { ok = 0; break; }
                    if (globaltimer() - t0 > WAIT_LIMIT_NS) { atomicExch(P.abort_flag, 1u); ok = 0; break; }
                }
            }
        }
        ok_s = ok;
    }
    __syncthreads();
    return ok_s != 0;
}

// Linearizer (modules/io.py:111-112): ((q / Q) - .5) * 2 in fp32
__device__ __forceinline__ float linearize(long long q, float Qf) {
    return __fmul_rn(__fsub_rn(__fdiv_rn((float)q, Qf), 0.5f), 2.0f);
}

// ------------------------------------------------------------------------------------------------------------
// Streamed contraction: acc[4 cols][4 prompts] += W[k][4 cq ..] * x[k][4 pq ..], x rows streamed from global
// ([K][ld] floats, prompts pb0 .. pb0 + pbw) through the cp.async stages.  The K rows of a chunk are dealt to
// `nslice` thread slices; the caller reduces the slices in slice order (fixed summation order).
// ------------------------------------------------------------------------------------------------------------
struct Map {
    int tiles, nslice, tile, slice, cq, pq, npq, kc;
    bool active;
};
// kc: rows per streamed chunk — 16 at full width, more for narrow batches so that a chunk stays ~8 KB and the
// per-chunk barrier is amortised (kc * pbw <= XSTAGE, kc divides K)
__device__ __forceinline__ Map make_map(int ncol4, int pbw, int part_cap, int K, bool wstream, int XSTAGE, int WSTAGE) {
    Map m;
    m.npq = pbw >> 2;
    m.tiles = ncol4 * m.npq;
    int kc = KC;
    while (kc * 2 * pbw <= XSTAGE && (!wstream || kc * 2 * ncol4 * 4 <= WSTAGE) && kc < 128 && K % (kc * 2) == 0) kc *= 2;
    m.kc = kc;
    int ns = NT / m.tiles, p2 = 1;
    while (p2 * 2 <= ns && p2 * 2 <= 16 && p2 * 2 <= kc && (p2 * 2 - 1) * m.tiles * 16 <= part_cap) p2 *= 2;
    m.nslice = p2;
    m.active = (int)threadIdx.x < m.tiles * p2;
    m.tile = threadIdx.x % m.tiles;
    m.slice = threadIdx.x / m.tiles;
    m.cq = m.tile / m.npq;
    m.pq = m.tile - m.cq * m.npq;
    return m;
}

// The frame-linear input term x = Linear(frame) + bias (+ conditioning already in the landed chunk), FramedLinearIO
// (modules/io.py:106-133).  It is added to a chunk by the thread that copied the element; the Linear rows of the
// Linear table (in_w | in_b) stays resident in shared memory.
struct FrameTerm {
    const float* inw_s; const float* inb_s; const float* lin_s;   // [H][fs], [H], [fs][pbw] — all in shared memory
    int fs;
    bool has_cond;
};

template <int R>
__device__ __forceinline__ void tile_rows(float (&acc)[16], const float* __restrict__ wp, int ldw,
                                          const float* __restrict__ xp, int pbw) {
#pragma unroll (R >= 4 ? 4 : R)
    for (int j = 0; j < R; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(wp + j * ldw);
        const float4 x = *reinterpret_cast<const float4*>(xp + j * pbw);
        const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
#pragma unroll
            for (int pi = 0; pi < 4; ++pi) acc[ci * 4 + pi] = fmaf(wv[ci], xv[pi], acc[ci * 4 + pi]);
    }
}

__device__ __forceinline__ void cp_async_wait_n(int n) {   // at most n groups still in flight
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    }
}

// W: resident weights in shared memory, or nullptr when Wg (the same [K][ldw] matrix in global memory) is streamed.
// Slice s of the map takes the kc / nslice consecutive rows s * (kc / nslice) .. of every chunk.
template <bool FRAME>
__device__ __forceinline__ void gemm_stream(float (&acc)[16], const Map& m, const float* __restrict__ W,
                                            const float* __restrict__ Wg, int ldw,
                                            const float* __restrict__ src, int ld, int pb0, int pbw, int K,
                                            float* stage, int XREGION, int WREGION, const FrameTerm ft, int rot,
                                            int exp = 0) {
    // rot: this CTA walks the K chunks starting at chunk `rot` (every CTA reads the same activations: starting them at
    // different rows keeps 128 SMs from asking the same L2 lines in the same cycle; the order is fixed per CTA)
    const int tid = threadIdx.x;
    const int kc = m.kc, n4row = pbw >> 2, n4 = kc * n4row, nchunk = K / kc, n4w = (kc * ldw) >> 2;
    const int xstride = kc * pbw, wstride = kc * ldw;
    // stages: as many as fit (chunks are small for narrow batches), so that the L2 latency hides behind >= 3 chunks
    int nst = min(MAXST, XREGION / xstride);
    if (W == nullptr) nst = min(nst, WREGION / wstride);
    float* wstage = stage + XREGION;
    if (rot >= nchunk) rot %= nchunk;
    // thread `tid` copies / fixes the float4 elements tid, tid + NT, ... of a chunk (row e / n4row, prompt quad e % n4row)
    auto chunk_of = [&](int c) { const int cc = c + rot; return cc >= nchunk ? cc - nchunk : cc; };
    int st_i = 0;                                              // stage of the chunk being issued
    auto issue = [&](int c) {
        if (c < nchunk && !(exp & 1)) {
            const int cc = chunk_of(c);
            if (src != nullptr)
                for (int e = tid; e < n4; e += NT) {
                    const int r = e / n4row, q = e - r * n4row;
                    cp_async16(stage + st_i * xstride + r * pbw + 4 * q, src + ((size_t)cc * kc + r) * ld + pb0 + 4 * q);
                }
            if (W == nullptr) {
                float* wb = wstage + st_i * wstride;
                for (int i = tid; i < n4w; i += NT) cp_async16(wb + 4 * i, Wg + (size_t)cc * wstride + 4 * i);
            }
        }
        cp_async_commit();
        st_i = st_i + 1 == nst ? 0 : st_i + 1;
    };
    for (int c = 0; c < nst - 1; ++c) issue(c);
    const int R = kc / m.nslice;
    int st_c = 0;                                              // stage of the chunk being consumed
    for (int c = 0; c < nchunk; ++c) {
        cp_async_wait_n(nst - 2);                           // this thread's copies of chunk c have landed
        float* buf = stage + st_c * xstride;
        if (FRAME) for (int e = tid; e < n4; e += NT) {
            const int r = e / n4row, ec = e - r * n4row;
            float4* x4 = reinterpret_cast<float4*>(buf + r * pbw + 4 * ec);
            const int k = chunk_of(c) * kc + r;
            const float* wr = ft.inw_s + (size_t)k * ft.fs;
            float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int f = 0; f < ft.fs; ++f) {
                const float wv = wr[f];
                const float4 l = *reinterpret_cast<const float4*>(ft.lin_s + f * pbw + 4 * ec);
                a[0] = fmaf(l.x, wv, a[0]); a[1] = fmaf(l.y, wv, a[1]);
                a[2] = fmaf(l.z, wv, a[2]); a[3] = fmaf(l.w, wv, a[3]);
            }
            const float bv = ft.inb_s[k];
            float4 xv = make_float4(a[0] + bv, a[1] + bv, a[2] + bv, a[3] + bv);
            if (ft.has_cond) { const float4 cv = *x4; xv.x += cv.x; xv.y += cv.y; xv.z += cv.z; xv.w += cv.w; }
            *x4 = xv;
        }
        if (!(exp & 4)) __syncthreads();                    // chunk c complete; everyone is done with chunk c - 1
        issue(c + nst - 1);                                 // into the stage chunk c - 1 occupied
        if (m.active && !(exp & 2)) {
            const float* wp = (W != nullptr ? W + (size_t)(chunk_of(c) * kc) * ldw : wstage + st_c * wstride)
                              + (size_t)(m.slice * R) * ldw + m.cq * 4;
            const float* xp = buf + (m.slice * R) * pbw + m.pq * 4;
            for (int r0 = 0; r0 < R; r0 += 16) {
                switch (R) {
                    case 1: tile_rows<1>(acc, wp, ldw, xp, pbw); break;
                    case 2: tile_rows<2>(acc, wp, ldw, xp, pbw); break;
                    case 4: tile_rows<4>(acc, wp, ldw, xp, pbw); break;
                    case 8: tile_rows<8>(acc, wp, ldw, xp, pbw); break;
                    default: tile_rows<16>(acc, wp + (size_t)r0 * ldw, ldw, xp + r0 * pbw, pbw); break;
                }
            }
        }
        st_c = st_c + 1 == nst ? 0 : st_c + 1;
    }
    __syncthreads();                                        // stage buffers are free again
}

// Slices >= 1 park their tile sums; after a __syncthreads the slice-0 thread of every tile adds them to its own
// registers in slice order (fixed summation order).
__device__ __forceinline__ void store_partials(const float (&acc)[16], const Map& m, float* part) {
    if (m.active && m.slice > 0) {
        float4* d = reinterpret_cast<float4*>(part + ((size_t)(m.slice - 1) * m.tiles + m.tile) * 16);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) d[ci] = make_float4(acc[ci * 4], acc[ci * 4 + 1], acc[ci * 4 + 2], acc[ci * 4 + 3]);
    }
}
__device__ __forceinline__ void reduce_partials(float (&acc)[16], const Map& m, const float* part) {
    for (int sl = 1; sl < m.nslice; sl += 2) {              // two slices per round: independent loads, ordered adds
        const float4* q0 = reinterpret_cast<const float4*>(part + ((size_t)(sl - 1) * m.tiles + m.tile) * 16);
        const float4* q1 = reinterpret_cast<const float4*>(part + ((size_t)sl * m.tiles + m.tile) * 16);
        const bool two = sl + 1 < m.nslice;
        float4 v0[4], v1[4];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) { v0[ci] = q0[ci]; v1[ci] = two ? q1[ci] : make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            acc[ci * 4] += v0[ci].x; acc[ci * 4 + 1] += v0[ci].y; acc[ci * 4 + 2] += v0[ci].z; acc[ci * 4 + 3] += v0[ci].w;
            if (two) { acc[ci * 4] += v1[ci].x; acc[ci * 4 + 1] += v1[ci].y; acc[ci * 4 + 2] += v1[ci].z; acc[ci * 4 + 3] += v1[ci].w; }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Lane-major frame-tier engine (Params::fast; H = 4 * NC, H in {128, 256, 512}, frame sizes and up-sampling factors
// in {1, 2, 4, 8, 16} / {1, 2, 4, 8}; GRU and LSTM tiers).  The column split over the CTAs stays (CTA c owns hidden indices 4c .. 4c + 3 and the
// up-sampler rows NV c .. NV c + NV - 1), but a contraction is laid out the other way round:
//   * activations live in HBM as [prompt][H]; lane l of K-quarter warp q owns k = 128 q + 4 l .. + 3, so a warp reads
//     its 512 bytes of a prompt's x and h rows with one coalesced 16-byte load per lane straight from L2 into
//     REGISTERS, one iteration (two prompts) ahead of their use — no shared-memory staging, no barrier per chunk.
//     (Measured, scripts/ubench/bulk_stream.cu: a cp.async.bulk costs its issuing thread ~600 cycles whatever its size,
//     so one elected thread feeding 8 KB chunks tops out at 14 B/clk per SM — the first version of this engine; cluster
//     multicast of 4 saves no L2 traffic; LDS.128 costs 4 cycles per warp instruction whatever the broadcast, which
//     bound the tile engine.)
//   * the lane's weights (24 float4 for the GRU, NV for the up-sampler) sit in REGISTERS for the whole firing, loaded
//     from L2 before the grid barrier that precedes the firing — no shared-memory weight traffic at all;
//   * the sums over k meet in a transposing shuffle tree: 16 values per lane -> 1 (a different one per lane pair) in
//     16 SHFL + 16 FADD.  The tree levels that fold hidden indices need no selects because the weights are packed
//     per lane with the indices pre-swapped (slot a on lane l is hidden index a ^ jm(l)); the two levels that fold
//     gates (r, z, n_i, n_h) use selects so that the zero blocks (W_ih has no n_h column) cost no FMA;
//   * K-quarter partials meet in shared memory; gates / biases run once per (prompt, hidden index).
// Every CTA reads the same rows: each starts at a different prompt block (rotation), so that the grid does not ask the
// same L2 lines in the same cycle — prompts are independent, any order gives the same bits.
// ------------------------------------------------------------------------------------------------------------
constexpr int NTF = 256;         // 8 warps: NKQ K-quarters x 8 / NKQ prompt groups (9 warps would cap the registers at 168)

__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Packed fp32 pairs (FFMA2 / FADD2, sm_100): one issue slot for two IEEE fp32 operations (the FMA pipe still spends two cycles —
// scripts/ubench/ffma2_bench.cu — so this buys issue slots for the shuffles, not FMA throughput)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 shfl2(u64 v, int bit) { return __shfl_xor_sync(0xffffffffu, v, bit); }

struct Fast {                    // shared-memory map of the lane-major engine
    float* part; float4* hold; float* lin;
    long long tl, ta[12];        // MMK_SR_DEBUG: SM cycles spent by CTA 0 per section of a firing
    bool timing;
};
__device__ __forceinline__ void fast_lap(Fast& F, int slot) {
    if (F.timing) { const long long n = clock64(); F.ta[slot] += n - F.tl; F.tl = n; }
}

// Row fast_col<NV>(l) is what lane l holds after the transposing tree over NV homogeneous values (up-sampler rows): the
// fold level that uses lane bit (16 >> j) pairs slot i with slot i + (NV >> (j + 1)); the lane with the bit set holds
// its slots pre-swapped (pack_up), so every lane keeps its lower half.
template <int NV>
__device__ __host__ __forceinline__ int fast_col(int lane) {
    int m = 0, j = 0;
    for (int half = NV / 2; half >= 1; half >>= 1, ++j) m |= ((lane >> (4 - j)) & 1) * half;
    return m;
}
template <int NV>
__device__ __forceinline__ bool fast_writer(int lane) {   // one lane per row writes: the lanes whose unused bits are 0
    int used = 0, j = 0;
    for (int half = NV / 2; half >= 1; half >>= 1, ++j) used |= 16 >> j;
    return (lane & ~used & 31) == 0;
}

// GRU cell of frame tier T at window end tw on this CTA's 4 hidden indices (sample_rnn_v2.py:226-260, modules/io.py:106-133).
template <int FS>
__device__ __forceinline__ bool fast_gru(const Params& P, const Tier& T, const float* cond, const float* hcur, float* hnext,
                                         long long tw, bool pre_barrier, unsigned long long& epoch, Fast& F) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    const int H = P.H, NKQ = P.NKQ, CHP = P.CHP, Bp = P.Bp, TW = NKQ * 32;
    const int kq = warp % NKQ, pg = warp / NKQ, NPG = 8 / NKQ;
    // 48 weight pairs per lane (pack_fast): [0,16) (W_ir, W_iz)[slot a][k], [16,24) (W_in[2a'], W_in[2a'+1])[k], [24,48) the same of W_hh
    constexpr bool HOIST = FS <= 8;                            // longer frames: the frame Linear rows are re-read from L1 per prompt
    ulonglong2 wq[24], iw[HOIST ? FS : 1], ib;
    const ulonglong2* iwp = reinterpret_cast<const ulonglong2*>(T.iw4) + kq * 32 + lane;
    {                                                          // issued before the barrier wait: the latency hides behind it
        const ulonglong2* w = reinterpret_cast<const ulonglong2*>(T.wg4) + (size_t)c * 24 * TW + kq * 32 + lane;
#pragma unroll
        for (int i = 0; i < 24; ++i) wq[i] = __ldg(w + i * TW);
        if (HOIST) {
#pragma unroll
            for (int f = 0; f < FS; ++f) iw[f] = __ldg(iwp + f * TW);
        }
        ib = __ldg(reinterpret_cast<const ulonglong2*>(T.ib4) + kq * 32 + lane);
    }
#define WQ(e) (((e) & 1) ? wq[(e) >> 1].y : wq[(e) >> 1].x)
    fast_lap(F, 0);
    if (pre_barrier && !grid_barrier(P, epoch)) return false;
    fast_lap(F, 1);
    const int Bl = (P.B + CHP - 1) / CHP * CHP, nchunk = Bl / CHP;
    const int rot = (int)(((long long)c * nchunk) / P.NC);
    auto rotated = [&](int ch) { const int x = ch + rot; return x >= nchunk ? x - nchunk : x; };
    const int koff = kq * 128 + 4 * lane;
    // the first two prompts' rows are on their way while the frame is linearised
    float4 hn[2], cn[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const size_t row = (size_t)(rotated(0) * CHP + pg + q * NPG) * H + koff;
        hn[q] = __ldcg(reinterpret_cast<const float4*>(hcur + row));
        cn[q] = cond != nullptr ? __ldcg(reinterpret_cast<const float4*>(cond + row)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    const float Qf = (float)P.Q;
    for (int idx = tid; idx < Bl * FS; idx += NTF) {           // Linearizer of the frame (modules/io.py:111-112)
        const int p = idx / FS, f = idx - p * FS;
        long long q = 0;
        if (p < P.B) q = __ldcg(P.seq + (size_t)p * P.seq_stride + (tw - FS + f));
        F.lin[idx] = linearize(q, Qf);
    }
    __syncthreads();
    fast_lap(F, 2);
    {
        const int j_lo = c * 4, hold_kq = j_lo / 128, hold_lane = (j_lo % 128) / 4;
        const bool b4 = (lane & 4) != 0, b2 = (lane & 2) != 0;
        const int out_slot = (b4 ? 8 : 0) + (b2 ? 4 : 0) + (((lane >> 4) & 1) << 1 | ((lane >> 3) & 1));   // gate * 4 + hidden index
        const bool nofma = (P.exp & 2) != 0;
        for (int ch = 0; ch < nchunk; ++ch) {                   // the warp's two prompts of the block, interleaved for ILP
            const int pr[2] = {rotated(ch) * CHP + pg, rotated(ch) * CHP + pg + NPG};
            float4 h4[2], c4[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) { h4[q] = hn[q]; c4[q] = cn[q]; }
            if (ch + 1 < nchunk && !(P.exp & 1)) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const size_t row = (size_t)(rotated(ch + 1) * CHP + pg + q * NPG) * H + koff;
                    hn[q] = __ldcg(reinterpret_cast<const float4*>(hcur + row));
                    if (cond != nullptr) cn[q] = __ldcg(reinterpret_cast<const float4*>(cond + row));
                }
            }
            if (nofma) continue;
            float kk[2];
            u64 rz[2][4], ni[2][2], nh[2][2];                   // (r, z) of slot a; (n_i, n_h) of slots 2a', 2a'+1; slot a = hidden index ^ jm(lane)
            u64 xd[2][4], hd[2][4];                             // (x_k, x_k), (h_k, h_k)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float* lp = F.lin + pr[q] * FS;
                u64 x01 = 0ull, x23 = 0ull;
#pragma unroll
                for (int f = 0; f < FS; ++f) {                 // FramedLinearIO: Linear(frame) + bias (+ conditioning)
                    const float l = lp[f];
                    const u64 ll = pack2(l, l);
                    const ulonglong2 wf = HOIST ? iw[HOIST ? f : 0] : __ldg(iwp + f * TW);
                    x01 = fma2(ll, wf.x, x01); x23 = fma2(ll, wf.y, x23);
                }
                x01 = add2(x01, ib.x); x23 = add2(x23, ib.y);
                if (cond != nullptr) { x01 = add2(x01, pack2(c4[q].x, c4[q].y)); x23 = add2(x23, pack2(c4[q].z, c4[q].w)); }
                float x0, x1, x2, x3;
                unpack2(x01, x0, x1); unpack2(x23, x2, x3);
                xd[q][0] = pack2(x0, x0); xd[q][1] = pack2(x1, x1); xd[q][2] = pack2(x2, x2); xd[q][3] = pack2(x3, x3);
                hd[q][0] = pack2(h4[q].x, h4[q].x); hd[q][1] = pack2(h4[q].y, h4[q].y);
                hd[q][2] = pack2(h4[q].z, h4[q].z); hd[q][3] = pack2(h4[q].w, h4[q].w);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int a = 0; a < 4; ++a) rz[q][a] = 0ull;
                ni[q][0] = ni[q][1] = nh[q][0] = nh[q][1] = 0ull;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)                         // input side, k ascending per accumulator ...
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) rz[q][a] = fma2(WQ(a * 4 + k), xd[q][k], rz[q][a]);
                    ni[q][0] = fma2(WQ(16 + k), xd[q][k], ni[q][0]);
                    ni[q][1] = fma2(WQ(20 + k), xd[q][k], ni[q][1]);
                }
#pragma unroll
            for (int k = 0; k < 4; ++k)                         // ... then the hidden side
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) rz[q][a] = fma2(WQ(24 + a * 4 + k), hd[q][k], rz[q][a]);
                    nh[q][0] = fma2(WQ(40 + k), hd[q][k], nh[q][0]);
                    nh[q][1] = fma2(WQ(44 + k), hd[q][k], nh[q][1]);
                }
            // hidden-index folds (pre-swapped slots), then gate folds (selects), then the last lane bit
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                rz[q][0] = add2(rz[q][0], shfl2(rz[q][2], 16)); rz[q][1] = add2(rz[q][1], shfl2(rz[q][3], 16));
                ni[q][0] = add2(ni[q][0], shfl2(ni[q][1], 16)); nh[q][0] = add2(nh[q][0], shfl2(nh[q][1], 16));
            }
            float g4[2][4];                                     // r, z, n_i, n_h of hidden index jm(lane)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                rz[q][0] = add2(rz[q][0], shfl2(rz[q][1], 8));
                float n0, n1, m0, m1;
                unpack2(ni[q][0], n0, n1); unpack2(nh[q][0], m0, m1);
                n0 += __shfl_xor_sync(0xffffffffu, n1, 8);
                m0 += __shfl_xor_sync(0xffffffffu, m1, 8);
                unpack2(rz[q][0], g4[q][0], g4[q][1]);
                g4[q][2] = n0; g4[q][3] = m0;
            }
            float k0[2], k1[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float s0 = b4 ? g4[q][0] : g4[q][2], s1 = b4 ? g4[q][1] : g4[q][3];
                k0[q] = b4 ? g4[q][2] : g4[q][0]; k1[q] = b4 ? g4[q][3] : g4[q][1];
                k0[q] += __shfl_xor_sync(0xffffffffu, s0, 4);
                k1[q] += __shfl_xor_sync(0xffffffffu, s1, 4);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float s2 = b2 ? k0[q] : k1[q];
                kk[q] = b2 ? k1[q] : k0[q];
                kk[q] += __shfl_xor_sync(0xffffffffu, s2, 2);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                kk[q] += __shfl_xor_sync(0xffffffffu, kk[q], 1);
                if ((lane & 1) == 0) F.part[((size_t)kq * Bp + pr[q]) * 16 + out_slot] = kk[q];
                if (kq == hold_kq && lane == hold_lane) F.hold[pr[q]] = h4[q];
            }
        }
    }
#undef WQ
    __syncthreads();
    fast_lap(F, 3);
    const float* gb = T.gb + (size_t)c * 24;
    for (int o = tid; o < Bl * 4; o += NTF) {                  // PyTorch GRU cell, gates r, z, n
        const int p = o >> 2, jj = o & 3;
        float sg[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) sg[g] = F.part[(size_t)p * 16 + g * 4 + jj];
        for (int q = 1; q < NKQ; ++q)
#pragma unroll
            for (int g = 0; g < 4; ++g) sg[g] += F.part[((size_t)q * Bp + p) * 16 + g * 4 + jj];
        const float r = sigmoid_acc((sg[0] + __ldg(gb + jj)) + __ldg(gb + 12 + jj));
        const float zg = sigmoid_acc((sg[1] + __ldg(gb + 4 + jj)) + __ldg(gb + 16 + jj));
        const float n = tanhf((sg[2] + __ldg(gb + 8 + jj)) + r * (sg[3] + __ldg(gb + 20 + jj)));
        const float hold = reinterpret_cast<const float*>(F.hold + p)[jj];
        const float hnew = (1.0f - zg) * n + zg * hold;
        if (p < P.B) __stcg(hnext + (size_t)p * H + c * 4 + jj, hnew);
    }
    fast_lap(F, 4);
    return true;
}

// nn.LSTM cell of frame tier T (the reference's default rnn_class, sample_rnn_v2.py:40-66) on this CTA's 4 hidden indices, lane-major
// fp32 form: the 16 gate rows (i, f, g, o of 4 indices) are 16 homogeneous columns fed by both operands — the contraction of fast_up<16>
// with [x | h] as its input.  64 weight pairs per lane (pack_fast: entry m * 32 + i2 * 4 + k = (W_m[slot 2 i2][k], W_m[slot 2 i2 + 1][k]),
// slot i = column i ^ fast_col<16>(lane), column = gate * 4 + hidden index).  The frame Linear rows are re-read from L1 per prompt
// when the frame is longer than 2 (the register file holds 128 weight registers here).
template <int FS>
__device__ __forceinline__ bool fast_lstm(const Params& P, const Tier& T, const float* cond, const float* hcur, float* hnext,
                                          const float* ccur, float* cnext, long long tw, bool pre_barrier, unsigned long long& epoch, Fast& F) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    const int H = P.H, NKQ = P.NKQ, CHP = P.CHP, Bp = P.Bp, TW = NKQ * 32;
    const int kq = warp % NKQ, pg = warp / NKQ, NPG = 8 / NKQ;
    constexpr bool HOIST = FS <= 2;
    ulonglong2 wq[32], iw[HOIST ? FS : 1], ib;
    const ulonglong2* iwp = reinterpret_cast<const ulonglong2*>(T.iw4) + kq * 32 + lane;
    {
        const ulonglong2* w = reinterpret_cast<const ulonglong2*>(T.wg4) + (size_t)c * 32 * TW + kq * 32 + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) wq[i] = __ldg(w + i * TW);
        if (HOIST) {
#pragma unroll
            for (int f = 0; f < FS; ++f) iw[f] = __ldg(iwp + f * TW);
        }
        ib = __ldg(reinterpret_cast<const ulonglong2*>(T.ib4) + kq * 32 + lane);
    }
#define WL(e) (((e) & 1) ? wq[(e) >> 1].y : wq[(e) >> 1].x)
    fast_lap(F, 0);
    if (pre_barrier && !grid_barrier(P, epoch)) return false;
    fast_lap(F, 1);
    const int Bl = (P.B + CHP - 1) / CHP * CHP, nchunk = Bl / CHP;
    const int rot = (int)(((long long)c * nchunk) / P.NC);
    auto rotated = [&](int ch) { const int x = ch + rot; return x >= nchunk ? x - nchunk : x; };
    const int koff = kq * 128 + 4 * lane;
    float4 hn[2], cn[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const size_t row = (size_t)(rotated(0) * CHP + pg + q * NPG) * H + koff;
        hn[q] = __ldcg(reinterpret_cast<const float4*>(hcur + row));
        cn[q] = cond != nullptr ? __ldcg(reinterpret_cast<const float4*>(cond + row)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    const float Qf = (float)P.Q;
    for (int idx = tid; idx < Bl * FS; idx += NTF) {           // Linearizer of the frame (modules/io.py:111-112)
        const int p = idx / FS, f = idx - p * FS;
        long long q = 0;
        if (p < P.B) q = __ldcg(P.seq + (size_t)p * P.seq_stride + (tw - FS + f));
        F.lin[idx] = linearize(q, Qf);
    }
    __syncthreads();
    fast_lap(F, 2);
    {
        const int col = fast_col<16>(lane);
        const bool writer = fast_writer<16>(lane);
        for (int ch = 0; ch < nchunk; ++ch) {
            const int pr[2] = {rotated(ch) * CHP + pg, rotated(ch) * CHP + pg + NPG};
            float4 h4[2], c4[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) { h4[q] = hn[q]; c4[q] = cn[q]; }
            if (ch + 1 < nchunk) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const size_t row = (size_t)(rotated(ch + 1) * CHP + pg + q * NPG) * H + koff;
                    hn[q] = __ldcg(reinterpret_cast<const float4*>(hcur + row));
                    if (cond != nullptr) cn[q] = __ldcg(reinterpret_cast<const float4*>(cond + row));
                }
            }
            u64 a2[2][8];                                       // (slot 2 i2, slot 2 i2 + 1)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float* lp = F.lin + pr[q] * FS;
                u64 x01 = 0ull, x23 = 0ull;
#pragma unroll
                for (int f = 0; f < FS; ++f) {                 // FramedLinearIO: Linear(frame) + bias (+ conditioning)
                    const float l = lp[f];
                    const u64 ll = pack2(l, l);
                    const ulonglong2 w = HOIST ? iw[HOIST ? f : 0] : __ldg(iwp + f * TW);
                    x01 = fma2(ll, w.x, x01); x23 = fma2(ll, w.y, x23);
                }
                x01 = add2(x01, ib.x); x23 = add2(x23, ib.y);
                if (cond != nullptr) { x01 = add2(x01, pack2(c4[q].x, c4[q].y)); x23 = add2(x23, pack2(c4[q].z, c4[q].w)); }
                float x0, x1, x2, x3;
                unpack2(x01, x0, x1); unpack2(x23, x2, x3);
                const u64 xd[4] = {pack2(x0, x0), pack2(x1, x1), pack2(x2, x2), pack2(x3, x3)};
                const u64 hd[4] = {pack2(h4[q].x, h4[q].x), pack2(h4[q].y, h4[q].y), pack2(h4[q].z, h4[q].z), pack2(h4[q].w, h4[q].w)};
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2) {
                    u64 a = 0ull;
#pragma unroll
                    for (int k = 0; k < 4; ++k) a = fma2(WL(i2 * 4 + k), xd[k], a);          // input side, k ascending ...
#pragma unroll
                    for (int k = 0; k < 4; ++k) a = fma2(WL(32 + i2 * 4 + k), hd[k], a);     // ... then the hidden side
                    a2[q][i2] = a;
                }
            }
            float acc[2];
            int bit = 16;
#pragma unroll
            for (int half = 4; half >= 1; half >>= 1, bit >>= 1)      // folds of whole pairs
#pragma unroll
                for (int i = 0; i < half; ++i)
#pragma unroll
                    for (int q = 0; q < 2; ++q) a2[q][i] = add2(a2[q][i], shfl2(a2[q][half + i], bit));
#pragma unroll
            for (int q = 0; q < 2; ++q) {                       // the last fold: slot 1 into slot 0
                float lo, hi;
                unpack2(a2[q][0], lo, hi);
                acc[q] = lo + __shfl_xor_sync(0xffffffffu, hi, 2);
                acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
            }
            if (writer)
#pragma unroll
                for (int q = 0; q < 2; ++q) F.part[((size_t)kq * Bp + pr[q]) * 16 + col] = acc[q];
        }
    }
#undef WL
    __syncthreads();
    fast_lap(F, 3);
    const float* gb = T.gb + (size_t)c * 32;                    // b_ih (16 columns), then b_hh
    for (int o = tid; o < Bl * 4; o += NTF) {                  // PyTorch LSTM cell, gates i, f, g, o
        const int p = o >> 2, jj = o & 3;
        float sg[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) sg[g] = F.part[(size_t)p * 16 + g * 4 + jj];
        for (int q = 1; q < NKQ; ++q)
#pragma unroll
            for (int g = 0; g < 4; ++g) sg[g] += F.part[((size_t)q * Bp + p) * 16 + g * 4 + jj];
        const float ig = sigmoid_acc((sg[0] + __ldg(gb + jj)) + __ldg(gb + 16 + jj));
        const float fg = sigmoid_acc((sg[1] + __ldg(gb + 4 + jj)) + __ldg(gb + 20 + jj));
        const float gg = tanhf((sg[2] + __ldg(gb + 8 + jj)) + __ldg(gb + 24 + jj));
        const float og = sigmoid_acc((sg[3] + __ldg(gb + 12 + jj)) + __ldg(gb + 28 + jj));
        const float cold = __ldcg(ccur + (size_t)p * H + c * 4 + jj);
        const float cnew = fg * cold + ig * gg;
        if (p < P.B) {
            __stcg(cnext + (size_t)p * H + c * 4 + jj, cnew);
            __stcg(hnext + (size_t)p * H + c * 4 + jj, og * tanhf(cnew));
        }
    }
    fast_lap(F, 4);
    return true;
}

// LinearResampler rows of this CTA (modules/resamplers.py:13-23) on the freshly written hidden state: always behind a
// grid barrier.  Row u = NV c + col of the (up * H) rows goes to obuf[u / H][prompt][u % H].
template <int NV>
__device__ __forceinline__ bool fast_up(const Params& P, const Tier& T, const float* hsrc, unsigned long long& epoch, Fast& F) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    const int H = P.H, NKQ = P.NKQ, CHP = P.CHP, Bp = P.Bp, TW = NKQ * 32;
    const int kq = warp % NKQ, pg = warp / NKQ, NPG = 8 / NKQ;
    constexpr int PS = NV > 16 ? 32 : 16;   // floats per (K quarter, prompt) in the partial-sum buffer
    ulonglong2 wu[NV];           // 2 NV weight pairs per lane (pack_up): entry i2 * 4 + k = (W[slot 2 i2][k], W[slot 2 i2 + 1][k])
    {
        const ulonglong2* w = reinterpret_cast<const ulonglong2*>(T.wu4) + (size_t)c * NV * TW + kq * 32 + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) wu[i] = __ldg(w + i * TW);
    }
#define WU(e) (((e) & 1) ? wu[(e) >> 1].y : wu[(e) >> 1].x)
    fast_lap(F, 5);
    if (!grid_barrier(P, epoch)) return false;
    fast_lap(F, 6);
    const int Bl = (P.B + CHP - 1) / CHP * CHP, nchunk = Bl / CHP;
    const int rot = (int)(((long long)c * nchunk) / P.NC);
    auto rotated = [&](int ch) { const int x = ch + rot; return x >= nchunk ? x - nchunk : x; };
    const int koff = kq * 128 + 4 * lane;
    {
        const int col = fast_col<NV>(lane);
        const bool writer = fast_writer<NV>(lane);
        const bool nofma = (P.exp & 2) != 0;
        float4 hn[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) hn[q] = __ldcg(reinterpret_cast<const float4*>(hsrc + (size_t)(rotated(0) * CHP + pg + q * NPG) * H + koff));
        for (int ch = 0; ch < nchunk; ++ch) {
            float4 h4[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) h4[q] = hn[q];
            if (ch + 1 < nchunk && !(P.exp & 1)) {
#pragma unroll
                for (int q = 0; q < 2; ++q) hn[q] = __ldcg(reinterpret_cast<const float4*>(hsrc + (size_t)(rotated(ch + 1) * CHP + pg + q * NPG) * H + koff));
            }
            if (nofma) continue;
            u64 a2[2][NV / 2];                                  // (slot 2 i2, slot 2 i2 + 1)
            float acc[2][1];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const u64 hd[4] = {pack2(h4[q].x, h4[q].x), pack2(h4[q].y, h4[q].y), pack2(h4[q].z, h4[q].z), pack2(h4[q].w, h4[q].w)};
#pragma unroll
                for (int i2 = 0; i2 < NV / 2; ++i2) {
                    u64 a = 0ull;
#pragma unroll
                    for (int k = 0; k < 4; ++k) a = fma2(WU(i2 * 4 + k), hd[k], a);
                    a2[q][i2] = a;
                }
            }
            int bit = 16;
#pragma unroll
            for (int half = NV / 4; half >= 1; half >>= 1, bit >>= 1)      // folds of whole pairs
#pragma unroll
                for (int i = 0; i < half; ++i)
#pragma unroll
                    for (int q = 0; q < 2; ++q) a2[q][i] = add2(a2[q][i], shfl2(a2[q][half + i], bit));
#pragma unroll
            for (int q = 0; q < 2; ++q) {                       // the last fold: slot 1 into slot 0
                float lo, hi;
                unpack2(a2[q][0], lo, hi);
                acc[q][0] = lo + __shfl_xor_sync(0xffffffffu, hi, bit);
            }
            bit >>= 1;
#pragma unroll
            for (; bit >= 1; bit >>= 1)
#pragma unroll
                for (int q = 0; q < 2; ++q) acc[q][0] += __shfl_xor_sync(0xffffffffu, acc[q][0], bit);
            if (writer)
#pragma unroll
                for (int q = 0; q < 2; ++q) F.part[((size_t)kq * Bp + rotated(ch) * CHP + pg + q * NPG) * PS + col] = acc[q][0];
        }
    }
#undef WU
    __syncthreads();
    fast_lap(F, 7);
    const float* ub = T.ub + (size_t)c * NV;
    for (int o = tid; o < Bl * NV; o += NTF) {
        const int p = o / NV, col = o - p * NV;
        float sv = F.part[(size_t)p * PS + col];
        for (int q = 1; q < NKQ; ++q) sv += F.part[((size_t)q * Bp + p) * PS + col];
        sv += __ldg(ub + col);
        const int urow = c * NV + col, slot = urow / H, kk = urow - slot * H;
        if (p < P.B) __stcg(T.obuf + ((size_t)slot * Bp + p) * H + kk, sv);
    }
    fast_lap(F, 8);
    return true;
}


// ------------------------------------------------------------------------------------------------------------
// Tensor-core frame-tier engine (Params::tc; compute mode bf16): the same column split (CTA c owns hidden indices
// 4c .. 4c + 3 and NV up-sampler rows), but a contraction is ONE tcgen05 accumulation: D[128 prompts x 16 columns] (TMEM,
// fp32) += A[128 x K] (activations, bf16) . B[16 x K]^T (this CTA's weight rows, bf16), K = [conditioning | hidden].
//   * the producers keep the activations in HBM as ready-made UMMA tiles (bf16, K-major, SWIZZLE_128B: 16 KB atoms of
//     128 prompts x 64 k): a CTA writes the 8 bytes of a prompt row it owns, every CTA pulls whole atoms with
//     cp.async.bulk — three issuing threads, one per stage, because a bulk copy costs its issuer ~600 cycles;
//   * the frame Linear is folded into the gate algebraically (W_ih (in_w lin + in_b) = (W_ih in_w) lin + W_ih in_b,
//     precomputed in fp32), so the A operand needs no per-CTA rebuild: it is the conditioning image and the hidden image;
//   * the epilogue thread of a prompt reads its 16 accumulator columns (r, z, n_i, n_h of 4 hidden indices) with one
//     tcgen05.ld, adds the folded terms, applies the GRU cell in fp32 and writes h' twice: fp32 (state, carry) and bf16
//     (the next operand).  The head stays the fp32 cluster head.
// ------------------------------------------------------------------------------------------------------------
constexpr int TC_STAGES = 3;
constexpr unsigned TC_A_BYTES = 32768u, TC_B_BYTES = 4096u;   // per fill of 128 k: 128 prompts x 128 k of A, 16 columns x 128 k of B
constexpr unsigned TC_BX_BYTES = 16384u;                      // widest B fill (fold_gx: up to 4 slots x 16 columns)
enum { TCB_FULL = 0, TCB_EMPTY = TC_STAGES, TCB_ACC = 2 * TC_STAGES, TCB_COUNT };

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned dst_smem, unsigned cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (as csrc/wavenet7.cu): a tile of R rows x K bf16 = K / 64 atoms of R x 128-byte
// lines, 8 rows = 1024 bytes (SBO), chunk c of line r at position c ^ (r & 7); atoms R * 128 bytes apart.
__host__ __device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3ffffu) >> 4);
    d |= (unsigned long long)1u << 16;
    d |= (unsigned long long)(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
__host__ __device__ __forceinline__ unsigned kstep16(int kk, int R) { return (unsigned)((kk >> 2) * (R * 8) + (kk & 3) * 2); }
__host__ __device__ __forceinline__ unsigned umma_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}
__host__ __device__ __forceinline__ unsigned tile_off(int m, int kc, int R) {   // byte offset of row m, 16-byte chunk kc
    return (unsigned)((kc >> 3) * (R * 128) + (m >> 3) * 1024 + (m & 7) * 128 + (((kc & 7) ^ (m & 7)) << 4));
}
__device__ __forceinline__ unsigned bf16x2_bits(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&v);
}

struct Tc {
    unsigned stage0, bar0, tmem;     // shared-memory address of stage 0 (1024-byte aligned), of the mbarriers; TMEM base
    unsigned gf, na;                 // fills streamed / accumulations finished so far (same in every thread)
    float* lin;
};
__device__ __forceinline__ unsigned tc_bar(const Tc& X, int i) { return X.bar0 + 8u * (unsigned)i; }

// Warps 5..7 (one issuing lane each, stage = warp - 5) stream the fills, warp 4 issues the MMAs: D (TMEM columns 0..15)
// = sum over the fills of A_fill[128 x 128] . B_fill[16 x 128]^T.  Fill j < n0 comes from img0, the rest from img1.
__device__ __forceinline__ bool tc_contract(const Params& P, Tc& X, const unsigned char* img0, int n0, const unsigned char* img1, int n1,
                                            const unsigned char* wimg, int warp, int lane, int ncols = 16) {
    const int nfill = n0 + n1;
    const unsigned bbytes = (unsigned)ncols * 256u;             // ncols rows x 128 k of bf16
    const unsigned sbytes = TC_A_BYTES + (unsigned)P.tc_bbytes; // stage stride
    bool ok = true;
    if (warp >= 5) {
        if (lane == 0) {
            const unsigned s = (unsigned)(warp - 5);
            fence_proxy_async_global();   // rows written with st.global by other CTAs before the grid barrier -> async-proxy reads
            for (int j = 0; j < nfill && ok; ++j) {
                const unsigned g = X.gf + (unsigned)j;
                if (g % TC_STAGES != s) continue;
                const unsigned u = g / TC_STAGES;
                if (u > 0) ok = mbar_wait(tc_bar(X, TCB_EMPTY + s), (u - 1u) & 1u, P.abort_flag);
                mbar_expect_tx(tc_bar(X, TCB_FULL + s), TC_A_BYTES + bbytes);
                const unsigned char* src = j < n0 ? img0 + (size_t)j * TC_A_BYTES : img1 + (size_t)(j - n0) * TC_A_BYTES;
                bulk_g2s(X.stage0 + s * sbytes, src, TC_A_BYTES, tc_bar(X, TCB_FULL + s));
                bulk_g2s(X.stage0 + s * sbytes + TC_A_BYTES, wimg + (size_t)j * bbytes, bbytes, tc_bar(X, TCB_FULL + s));
            }
        }
    } else if (warp == 4) {
        const unsigned idesc = umma_idesc(ncols);
        for (int j = 0; j < nfill; ++j) {
            const unsigned g = X.gf + (unsigned)j, s = g % TC_STAGES, u = g / TC_STAGES;
            ok = __all_sync(0xffffffffu, (mbar_wait(tc_bar(X, TCB_FULL + s), u & 1u, P.abort_flag) && ok) ? 1 : 0) != 0;
            if (!ok) break;
            tc_fence_after();
            unsigned pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) {
                const unsigned long long dA = umma_desc(X.stage0 + s * sbytes), dB = umma_desc(X.stage0 + s * sbytes + TC_A_BYTES);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) umma_bf16(X.tmem, dA + kstep16(kk, 128), dB + kstep16(kk, ncols), idesc, (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(tc_bar(X, TCB_EMPTY + s));
                if (j == nfill - 1) umma_commit(tc_bar(X, TCB_ACC));
            }
            __syncwarp();
        }
    }
    return ok;
}

// GRU cell of frame tier T on this CTA's 4 hidden indices (sample_rnn_v2.py:226-260, modules/io.py:106-133), bf16 tensor-core form.
template <int FS>
__device__ __forceinline__ bool tc_gru(const Params& P, const Tier& T, const unsigned char* cimg, const unsigned char* himg, unsigned char* himg_next,
                                       const float* hcur, float* hnext, long long tw, bool pre_barrier, unsigned long long& epoch, Tc& X, int hs,
                                       const float* gxrow) {   // fold_gx: this CTA's [128 prompts][NC][16] input-side pre-activations of the slot
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x, H = P.H;
    if (pre_barrier && !grid_barrier(P, epoch)) return false;
    const int nimg = H / 128;                                   // fills per image
    // weight fills of the CTA: [conditioning part | hidden part] (the top tier: hidden part first); with fold_gx only the hidden part runs
    const unsigned char* wimg = T.wgimg + (size_t)c * (2 * nimg) * TC_B_BYTES + (gxrow ? (size_t)nimg * TC_B_BYTES : 0);
    bool ok = tc_contract(P, X, cimg ? cimg : himg, nimg, himg, cimg ? nimg : 0, wimg, warp, lane);
    if (warp < 4) {
        const int p = tid;
        // the state this thread carries: h_old (GRU) or c_old (LSTM: cbuf ping-pongs with the hidden state, side hs)
        const float* sold = P.lstm ? T.cbuf + (size_t)hs * H * P.Bp : hcur;
        const float4 hold = __ldcg(reinterpret_cast<const float4*>(sold + (size_t)p * H + 4 * c));
        {   // Linearizer of the prompt's frame (modules/io.py:111-112), under the fills and the MMAs
            const float Qf = (float)P.Q;
            long long q[FS];
#pragma unroll
            for (int f = 0; f < FS; ++f) q[f] = p < P.B ? __ldcg(P.seq + (size_t)p * P.seq_stride + (tw - FS + f)) : 0;
#pragma unroll
            for (int f = 0; f < FS; ++f) X.lin[p * FS + f] = linearize(q[f], Qf);
        }
        const float* wf = T.wf + (size_t)c * 16 * FS;
        const float* bfo = T.bfold + (size_t)c * 16;
        float4 gxv[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        if (gxrow) {
#pragma unroll
            for (int i = 0; i < 4; ++i) gxv[i] = __ldcg(reinterpret_cast<const float4*>(gxrow + ((size_t)p * P.NC + c) * 16) + i);
        }
        float pre[16];
#pragma unroll
        for (int col = 0; col < 16; ++col) {
            float a = __ldg(bfo + col);
            if (col < 12 || P.lstm) {                           // GRU: the n_h column takes no input term
#pragma unroll
                for (int f = 0; f < FS; ++f) a = fmaf(__ldg(wf + col * FS + f), X.lin[p * FS + f], a);
            }
            pre[col] = a + reinterpret_cast<const float*>(gxv)[col];
        }
        ok = mbar_wait(tc_bar(X, TCB_ACC), X.na & 1u, P.abort_flag) && ok;
        tc_fence_after();
        float v[16];
        tmem_ld16(X.tmem + ((unsigned)(32 * warp) << 16), v);
        tmem_ld_wait();
        tc_fence_before();
        const float ho[4] = {hold.x, hold.y, hold.z, hold.w};
        float hn[4], cn[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (P.lstm) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {                    // PyTorch LSTM cell, gates i, f, g, o
                const float ig = sigmoid_acc(v[jj] + pre[jj]);
                const float fg = sigmoid_acc(v[4 + jj] + pre[4 + jj]);
                const float gg = tanhf(v[8 + jj] + pre[8 + jj]);
                const float og = sigmoid_acc(v[12 + jj] + pre[12 + jj]);
                cn[jj] = fg * ho[jj] + ig * gg;
                hn[jj] = og * tanhf(cn[jj]);
            }
            if (p < P.B)
                __stcg(reinterpret_cast<float4*>(T.cbuf + (size_t)(hs ^ 1) * H * P.Bp + (size_t)p * H + 4 * c), make_float4(cn[0], cn[1], cn[2], cn[3]));
        } else {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {                    // PyTorch GRU cell, gates r, z, n
                const float r = sigmoid_acc(v[jj] + pre[jj]);
                const float zg = sigmoid_acc(v[4 + jj] + pre[4 + jj]);
                const float n = tanhf((v[8 + jj] + pre[8 + jj]) + r * (v[12 + jj] + pre[12 + jj]));
                hn[jj] = (1.0f - zg) * n + zg * ho[jj];
            }
        }
        if (p < P.B) {
            __stcg(reinterpret_cast<float4*>(hnext + (size_t)p * H + 4 * c), make_float4(hn[0], hn[1], hn[2], hn[3]));
            const int k0 = 4 * c;
            __stcg(reinterpret_cast<uint2*>(himg_next + tile_off(p, k0 >> 3, 128) + (k0 & 7) * 2),
                   make_uint2(bf16x2_bits(hn[0], hn[1]), bf16x2_bits(hn[2], hn[3])));
        }
    }
    X.gf += (unsigned)(cimg ? 2 * nimg : nimg);
    X.na += 1u;
    return __syncthreads_or(ok ? 0 : 1) == 0;
}

// fold_gx: the up-sampler of tier T as the input-side gate pre-activations of the tier below, for this CTA's own columns: one
// accumulation D[128 x up * 16] = h_new . (W_ih(next) W_up(slot))^T, written to T.gx [slot][prompt][CTA][16] (fp32; same CTA reads it).
__device__ __forceinline__ bool tc_up_gx(const Params& P, const Tier& T, const unsigned char* himg, unsigned long long& epoch, Tc& X) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x, H = P.H;
    const int ncols = T.up * 16;
    if (!grid_barrier(P, epoch)) return false;
    const int nimg = H / 128;
    bool ok = tc_contract(P, X, himg, nimg, himg, 0, T.wximg + (size_t)c * nimg * ((size_t)ncols * 256), warp, lane, ncols);
    if (warp < 4) {
        const int p = tid;
        ok = mbar_wait(tc_bar(X, TCB_ACC), X.na & 1u, P.abort_flag) && ok;
        tc_fence_after();
        for (int sl = 0; sl < T.up; ++sl) {
            float v[16];
            tmem_ld16(X.tmem + ((unsigned)(32 * warp) << 16) + 16u * (unsigned)sl, v);
            tmem_ld_wait();
            const float* xb = T.xb + ((size_t)c * T.up + sl) * 16;
            float4* dst = reinterpret_cast<float4*>(T.gx + (((size_t)sl * 128 + p) * P.NC + c) * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                __stcg(dst + i, make_float4(v[4 * i] + __ldg(xb + 4 * i), v[4 * i + 1] + __ldg(xb + 4 * i + 1), v[4 * i + 2] + __ldg(xb + 4 * i + 2),
                                            v[4 * i + 3] + __ldg(xb + 4 * i + 3)));
        }
        tc_fence_before();
    }
    X.gf += (unsigned)nimg;
    X.na += 1u;
    return __syncthreads_or(ok ? 0 : 1) == 0;
}

// LinearResampler rows of this CTA (modules/resamplers.py:13-23) on the new hidden image: always behind a grid barrier.
__device__ __forceinline__ bool tc_up(const Params& P, const Tier& T, const unsigned char* himg, unsigned long long& epoch, Tc& X, bool fold) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x, H = P.H;
    const int NV = fold ? P.NVh : T.NV;                         // columns of this CTA
    const int RL = fold ? P.Hh : H;                             // row length of the destination
    if (!grid_barrier(P, epoch)) return false;
    const int nimg = H / 128;
    bool ok = tc_contract(P, X, himg, nimg, himg, 0, T.wuimg + (size_t)c * nimg * TC_B_BYTES, warp, lane);
    if (warp < 4) {
        const int p = tid;
        ok = mbar_wait(tc_bar(X, TCB_ACC), X.na & 1u, P.abort_flag) && ok;
        tc_fence_after();
        float v[16];
        tmem_ld16(X.tmem + ((unsigned)(32 * warp) << 16), v);
        tmem_ld_wait();
        tc_fence_before();
        const float* ub = fold ? P.hb + (size_t)c * NV : T.ub + (size_t)c * NV;
        const int urow = c * NV, slot = urow / RL, k0 = urow - slot * RL;   // NV divides the row length: the CTA's rows share a slot
#pragma unroll
        for (int col = 0; col < 16; ++col) v[col] += col < NV ? __ldg(ub + col) : 0.0f;
        if (p < P.B) {
            if (!fold && T.oimg != nullptr) {                   // conditioning of the next frame tier: bf16 image of the slot
                unsigned char* img = T.oimg + (size_t)slot * ((size_t)H * 256);
                if (NV == 4) {
                    __stcg(reinterpret_cast<uint2*>(img + tile_off(p, k0 >> 3, 128) + (k0 & 7) * 2), make_uint2(bf16x2_bits(v[0], v[1]), bf16x2_bits(v[2], v[3])));
                } else {
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8)
                        if (h8 * 8 < NV)
                            __stcg(reinterpret_cast<uint4*>(img + tile_off(p, (k0 >> 3) + h8, 128)),
                                   make_uint4(bf16x2_bits(v[h8 * 8], v[h8 * 8 + 1]), bf16x2_bits(v[h8 * 8 + 2], v[h8 * 8 + 3]),
                                              bf16x2_bits(v[h8 * 8 + 4], v[h8 * 8 + 5]), bf16x2_bits(v[h8 * 8 + 6], v[h8 * 8 + 7])));
                }
            } else {                                            // the head reads fp32: [slot][prompt][H] rows of the up-sampler, or (folded)
                float* dst = (fold ? P.pre : T.obuf) + ((size_t)slot * P.Bp + p) * RL + k0;   // [slot][prompt][Hh] hidden pre-activations
                if (NV >= 4) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        if (q4 * 4 < NV) __stcg(reinterpret_cast<float4*>(dst + 4 * q4), make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]));
                } else {
#pragma unroll
                    for (int q1 = 0; q1 < 4; ++q1)
                        if (q1 < NV) __stcg(dst + q1, v[q1]);
                }
            }
        }
    }
    X.gf += (unsigned)nimg;
    X.na += 1u;
    return __syncthreads_or(ok ? 0 : 1) == 0;
}

enum { BAR_HID = 0, BAR_Z, BAR_Q, BAR_COUNT };

// ------------------------------------------------------------------------------------------------------------
template <int ENGINE>   // 0: tile engine, 1: lane-major fp32 engine, 2: tensor-core bf16 engine (frame tiers; the head is the same)
__global__ void __launch_bounds__(ENGINE ? NTF : NT, 1) samplernn_cluster_kernel(const __grid_constant__ Params P) {
    constexpr bool FAST = ENGINE != 0;
    constexpr int NTK = FAST ? NTF : NT;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, NC = P.NC, H = P.H, Bp = P.Bp, CS = P.CS;
    const int rank = (int)cluster_ctarank();
    const int cluster = c / CS, n_clusters = NC / CS;
    const float* w_s = smem;
    float* region = smem + P.s_region;
    float* gi_s = smem + P.s_gi;          // [NG][pbw] input-side gate pre-activations; before that: lin_s[fs][pbw]
    const unsigned sbase = smem_u32(smem);
    const unsigned off_bar0 = (unsigned)P.s_bar * 4u;
    auto bar = [&](int i) { return sbase + off_bar0 + 8u * (unsigned)i; };
    unsigned* abort_flag = P.abort_flag;

    const float* gblock = P.wpack + (size_t)c * P.cta_block;   // this CTA's packed weights in global memory
    for (int r = 0; r < P.n_res; ++r) {   // resident pieces (what does not fit is streamed from L2 when used)
        const float4* src = reinterpret_cast<const float4*>(gblock + P.res_goff[r]);
        float4* dst = reinterpret_cast<float4*>(smem + P.res_soff[r]);
        for (int i = tid; i < P.res_len[r] / 4; i += NTK) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i <= P.Q; i += NTK) smem[P.s_b2 + i] = __ldg(P.b2 + i);
    const int GP = P.GP, npq_h = GP >> 2;
    const unsigned hid_bytes = (unsigned)(CS * P.RS * GP) * 4u;
    const unsigned z_bytes = (unsigned)((GP / CS) * CS * P.ZR) * 4u;
    const unsigned q_bytes = (unsigned)GP * 4u;
    if (tid == 0) {
        for (int i = 0; i < BAR_COUNT; ++i) mbar_init(bar(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar(BAR_HID), hid_bytes);
        mbar_expect_tx(bar(BAR_Z), z_bytes);
        mbar_expect_tx(bar(BAR_Q), q_bytes);
    }
    __syncthreads();
    cluster_sync_all();

    Tc X{};
    if (ENGINE == 2) {
        X.stage0 = (sbase + (unsigned)P.s_region * 4u + 1023u) & ~1023u;      // the region carries 1 KB of slack for this
        X.bar0 = sbase + (unsigned)P.s_tcbar * 4u;
        X.lin = smem + P.s_lin;
        unsigned* s_tmem = reinterpret_cast<unsigned*>(smem + P.s_tcbar) + 2 * TCB_COUNT;
        if (tid == 0) {
            for (int i = 0; i < TCB_COUNT; ++i) mbar_init(tc_bar(X, i), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 4) { tmem_alloc(smem_u32(s_tmem), 64); tmem_relinquish(); }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        X.tmem = *reinterpret_cast<volatile unsigned*>(s_tmem);
    }
    Fast F{};
    if (ENGINE == 1) {
        F.part = smem + P.s_part; F.hold = reinterpret_cast<float4*>(smem + P.s_hold); F.lin = smem + P.s_lin;
        F.timing = P.dbg && c == 0 && tid == 0; F.tl = clock64();
        for (int i = 0; i < 12; ++i) F.ta[i] = 0;
    }
    const int j_lo = c * P.JP;
    const int rot = (int)(((unsigned)c * 11u) % (unsigned)(H / KC));   // reduced modulo the chunk count where used
    int hsel[MAX_TIERS];
#pragma unroll
    for (int i = 0; i < MAX_TIERS; ++i) hsel[i] = P.hsel[i];
    unsigned long long epoch = 0;
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = globaltimer();
    auto lap = [&](int slot) { if (!FAST && P.dbg && c == 0 && tid == 0) { const unsigned long long n = globaltimer(); tacc[slot] += n - tlast; tlast = n; } };
    long long hta[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, htl = clock64();   // MMK_SR_DEBUG: SM cycles of CTA 0 per head section (+ everything else)
    auto hlap = [&](int slot) { if (FAST && P.dbg && c == 0 && tid == 0) { const long long n = clock64(); hta[slot] += n - htl; htl = n; } };
    unsigned head_phase = 0;              // completed head exchanges (parity of the three mbarriers)
    bool dead = false;
    bool heads_pending = false;           // head steps ran since the last grid barrier
    const float Qf = (float)P.Q;
    const int n_groups = (P.B + GP - 1) / GP;
    const int Bl = (P.B + 3) & ~3;        // live prompts, padded to the float4 granule

    for (int phase = 0; phase < 2 && !dead; ++phase) {
        const bool gen = phase == 1;
        const long long t_lo = gen ? P.gen_begin : P.warm_begin, t_hi = gen ? P.gen_end : P.warm_end;
        const long long off = gen ? 0 : P.warm_off;
        for (long long t = t_lo; t < t_hi && !dead; ++t) {
            const long long tw = t + off;   // the window ends at data index tw (exclusive)
            // ---------------- frame tiers (weight-stationary over the whole grid) ----------------
            bool fired = false;
            if constexpr (ENGINE == 2) {
                bool pending_up = false;
                for (int i = 0; i < P.n_ft && !dead; ++i) {
                    const Tier& T = P.tiers[i];
                    if (t % T.fs != 0) continue;
                    const bool pre = pending_up || heads_pending;
                    heads_pending = false;
                    fired = true;
                    const size_t img = (size_t)H * 256;           // bytes of a [128 x H] bf16 image
                    const bool gx_in = P.fold_gx && i > 0;                  // the tier above emitted this tier's input-side pre-activations
                    const unsigned char* cimg = (i > 0 && !gx_in) ? P.tiers[i - 1].oimg + (size_t)((t / T.fs) % T.kdiv) * img : nullptr;
                    const float* gxrow = gx_in ? P.tiers[i - 1].gx + (size_t)((t / T.fs) % T.kdiv) * 128 * P.NC * 16 : nullptr;
                    const unsigned char* himg = T.himg + (size_t)hsel[i] * img;
                    unsigned char* himg_next = T.himg + (size_t)(hsel[i] ^ 1) * img;
                    const float* hcur = T.hbuf + (size_t)hsel[i] * H * Bp;
                    float* hnext = T.hbuf + (size_t)(hsel[i] ^ 1) * H * Bp;
                    bool ok = false;
                    switch (T.fs) {
                        case 1: ok = tc_gru<1>(P, T, cimg, himg, himg_next, hcur, hnext, tw, pre, epoch, X, hsel[i], gxrow); break;
                        case 2: ok = tc_gru<2>(P, T, cimg, himg, himg_next, hcur, hnext, tw, pre, epoch, X, hsel[i], gxrow); break;
                        case 4: ok = tc_gru<4>(P, T, cimg, himg, himg_next, hcur, hnext, tw, pre, epoch, X, hsel[i], gxrow); break;
                        case 8: ok = tc_gru<8>(P, T, cimg, himg, himg_next, hcur, hnext, tw, pre, epoch, X, hsel[i], gxrow); break;
                        default: ok = tc_gru<16>(P, T, cimg, himg, himg_next, hcur, hnext, tw, pre, epoch, X, hsel[i], gxrow); break;
                    }
                    hsel[i] ^= 1;
                    if (ok) ok = (P.fold_gx && i < P.n_ft - 1) ? tc_up_gx(P, T, himg_next, epoch, X)
                                                               : tc_up(P, T, himg_next, epoch, X, P.fold_head != 0 && i == P.n_ft - 1);
                    if (!ok) { dead = true; break; }
                    pending_up = true;
                }
                if (pending_up && !dead) {
                    if (gen || t == t_hi - 1) { if (!grid_barrier(P, epoch)) dead = true; }
                    else heads_pending = true;
                }
            } else if constexpr (ENGINE == 1) {
                bool pending_up = false;
                for (int i = 0; i < P.n_ft && !dead; ++i) {
                    const Tier& T = P.tiers[i];
                    if (t % T.fs != 0) continue;
                    fast_lap(F, 11);
                    const bool pre = pending_up || heads_pending;     // last head samples / last up-sampler rows must be visible
                    heads_pending = false;
                    fired = true;
                    const float* cond = nullptr;
                    if (i > 0) cond = P.tiers[i - 1].obuf + (size_t)((t / T.fs) % T.kdiv) * H * Bp;
                    const float* hcur = T.hbuf + (size_t)hsel[i] * H * Bp;
                    float* hnext = T.hbuf + (size_t)(hsel[i] ^ 1) * H * Bp;
                    bool ok = false;
                    if (P.lstm) {
                        const float* ccur = T.cbuf + (size_t)hsel[i] * H * Bp;
                        float* cnext = T.cbuf + (size_t)(hsel[i] ^ 1) * H * Bp;
                        switch (T.fs) {
                            case 1: ok = fast_lstm<1>(P, T, cond, hcur, hnext, ccur, cnext, tw, pre, epoch, F); break;
                            case 2: ok = fast_lstm<2>(P, T, cond, hcur, hnext, ccur, cnext, tw, pre, epoch, F); break;
                            case 4: ok = fast_lstm<4>(P, T, cond, hcur, hnext, ccur, cnext, tw, pre, epoch, F); break;
                            case 8: ok = fast_lstm<8>(P, T, cond, hcur, hnext, ccur, cnext, tw, pre, epoch, F); break;
                            default: ok = fast_lstm<16>(P, T, cond, hcur, hnext, ccur, cnext, tw, pre, epoch, F); break;
                        }
                    } else
                    switch (T.fs) {
                        case 1: ok = fast_gru<1>(P, T, cond, hcur, hnext, tw, pre, epoch, F); break;
                        case 2: ok = fast_gru<2>(P, T, cond, hcur, hnext, tw, pre, epoch, F); break;
                        case 4: ok = fast_gru<4>(P, T, cond, hcur, hnext, tw, pre, epoch, F); break;
                        case 8: ok = fast_gru<8>(P, T, cond, hcur, hnext, tw, pre, epoch, F); break;
                        default: ok = fast_gru<16>(P, T, cond, hcur, hnext, tw, pre, epoch, F); break;
                    }
                    hsel[i] ^= 1;
                    if (ok) switch (T.NV) {
                        case 4: ok = fast_up<4>(P, T, hnext, epoch, F); break;
                        case 8: ok = fast_up<8>(P, T, hnext, epoch, F); break;
                        case 16: ok = fast_up<16>(P, T, hnext, epoch, F); break;
                        default: ok = fast_up<32>(P, T, hnext, epoch, F); break;
                    }
                    if (!ok) { dead = true; break; }
                    pending_up = true;
                }
                if (pending_up && !dead) {
                    if (gen || t == t_hi - 1) { if (!grid_barrier(P, epoch)) dead = true; }   // the head reads the bottom tier's rows
                    else heads_pending = true;                               // warm-up: the next firing's barrier covers it
                    fast_lap(F, 9);
                }
            } else {
            for (int i = 0; i < P.n_ft && !dead; ++i) {
                const Tier& T = P.tiers[i];
                if (t % T.fs != 0) continue;
                lap(0);
                if (!fired && heads_pending) {              // the samples of the last head steps must be visible
                    if (!grid_barrier(P, epoch)) { dead = true; break; }
                    heads_pending = false;
                }
                lap(1);
                fired = true;
                const float* cond = nullptr;
                if (i > 0) cond = P.tiers[i - 1].obuf + (size_t)((t / T.fs) % T.kdiv) * H * Bp;
                const float* hcur = T.hbuf + (size_t)hsel[i] * H * Bp;
                float* hnext = T.hbuf + (size_t)(hsel[i] ^ 1) * H * Bp;
                const float* Wih_g = gblock + T.off_wih;
                const float* Wih = T.so_wih >= 0 ? w_s + T.so_wih : nullptr;
                const float* bih = T.so_wih >= 0 ? Wih + (size_t)H * P.NG : Wih_g + (size_t)H * P.NG;
                const float* Whh_g = gblock + T.off_whh;
                const float* Whh = T.so_whh >= 0 ? w_s + T.so_whh : nullptr;
                const float* bhh = T.so_whh >= 0 ? Whh + (size_t)H * P.NG : Whh_g + (size_t)H * P.NG;
                const float* inw_s = w_s + T.so_inw;
                const int fs = T.fs;
                // ---- GRU cell on this CTA's hidden indices ----
                const int cap_g = min(PBW, (NTK / (P.NG >> 2)) << 2);
                float* gh_s = region + P.region - P.NG * PBW;          // [NG][pbw] hidden-side gate pre-activations
                for (int pb0 = 0; pb0 < Bl; pb0 += cap_g) {
                    const int pbw = min(cap_g, Bl - pb0);
                    float* lin_s = gi_s;
                    for (int idx = tid; idx < fs * pbw; idx += NTK) {
                        const int f = idx / pbw, p = idx - f * pbw, b = pb0 + p;
                        long long q = 0;
                        if (b < P.B) q = __ldcg(P.seq + (size_t)b * P.seq_stride + (tw - fs + f));
                        lin_s[idx] = linearize(q, Qf);
                    }
                    __syncthreads();
                    const Map m = make_map(P.NG >> 2, pbw, P.region - P.NG * PBW, H, Whh == nullptr || Wih == nullptr, P.xstage, P.wstage);
                    float acc[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
                    {
                        FrameTerm ft{inw_s, inw_s + (size_t)H * fs, lin_s, fs, cond != nullptr};
                        gemm_stream<true>(acc, m, Wih, Wih_g, P.NG, cond, Bp, pb0, pbw, H, region, P.xregion, P.wregion, ft, rot, P.exp);
                    }
                    lap(2);
                    store_partials(acc, m, region);
                    __syncthreads();
                    if (m.active && m.slice == 0) {                   // gi = W_ih x + b_ih  -> gi_s[col][p]
                        reduce_partials(acc, m, region);
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) {
                            const int col = m.cq * 4 + ci;
                            const float bv = bih[col];
                            *reinterpret_cast<float4*>(gi_s + col * pbw + 4 * m.pq) =
                                make_float4(acc[ci * 4] + bv, acc[ci * 4 + 1] + bv, acc[ci * 4 + 2] + bv, acc[ci * 4 + 3] + bv);
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
                    gemm_stream<false>(acc, m, Whh, Whh_g, P.NG, hcur, Bp, pb0, pbw, H, region, P.xregion, P.wregion, FrameTerm{}, rot, P.exp);
                    store_partials(acc, m, region);
                    __syncthreads();
                    if (m.active && m.slice == 0) {                   // gh = W_hh h + b_hh  -> gh_s[col][p]
                        reduce_partials(acc, m, region);
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) {
                            const int col = m.cq * 4 + ci;
                            const float bv = bhh[col];
                            *reinterpret_cast<float4*>(gh_s + col * pbw + 4 * m.pq) =
                                make_float4(acc[ci * 4] + bv, acc[ci * 4 + 1] + bv, acc[ci * 4 + 2] + bv, acc[ci * 4 + 3] + bv);
                        }
                    }
                    __syncthreads();
                    for (int o = tid; o < P.JP * pbw; o += NTK) {      // PyTorch GRU cell, gates r, z, n
                        const int jj = o / pbw, p = o - jj * pbw;
                        const int cr = jj, cz = P.JP + jj, cn = 2 * P.JP + jj;
                        const float r = sigmoid_acc(gi_s[cr * pbw + p] + gh_s[cr * pbw + p]);
                        const float zg = sigmoid_acc(gi_s[cz * pbw + p] + gh_s[cz * pbw + p]);
                        const float n = tanhf(gi_s[cn * pbw + p] + r * gh_s[cn * pbw + p]);
                        const float hold = __ldcg(hcur + (size_t)(j_lo + jj) * Bp + pb0 + p);
                        const float hnew = (1.0f - zg) * n + zg * hold;
                        if (pb0 + p < P.B) __stcg(hnext + (size_t)(j_lo + jj) * Bp + pb0 + p, hnew);
                    }
                    __syncthreads();
                }
                hsel[i] ^= 1;
                lap(3);
                if (!grid_barrier(P, epoch)) { dead = true; break; }
                lap(4);
                // ---- LinearResampler rows of this CTA (modules/resamplers.py:13-23) ----
                {
                    const int nu = T.up_rows / NC, u_lo = c * nu;
                    const float* Wup_g = gblock + T.off_wup;
                    const float* Wup = T.so_wup >= 0 ? w_s + T.so_wup : nullptr;
                    const float* bup = T.so_wup >= 0 ? Wup + (size_t)H * T.NU : Wup_g + (size_t)H * T.NU;
                    const int cap_u = min(PBW, (NTK / (T.NU >> 2)) << 2);
                    for (int pb0 = 0; pb0 < Bl; pb0 += cap_u) {
                        const int pbw = min(cap_u, Bl - pb0);
                        const Map m = make_map(T.NU >> 2, pbw, P.region, H, Wup == nullptr, P.xstage, P.wstage);
                        float acc[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
                        gemm_stream<false>(acc, m, Wup, Wup_g, T.NU, hnext, Bp, pb0, pbw, H, region, P.xregion, P.wregion, FrameTerm{}, rot, P.exp);
                        store_partials(acc, m, region);
                        __syncthreads();
                        if (m.active && m.slice == 0) {
                            reduce_partials(acc, m, region);
#pragma unroll
                            for (int ci = 0; ci < 4; ++ci) {
                                const int col = m.cq * 4 + ci;
                                if (col < nu) {                       // padded prompts (>= B) receive values nobody reads
                                    const float bv = bup[col];
                                    __stcg(reinterpret_cast<float4*>(T.obuf + (size_t)(u_lo + col) * Bp + pb0 + 4 * m.pq),
                                           make_float4(acc[ci * 4] + bv, acc[ci * 4 + 1] + bv, acc[ci * 4 + 2] + bv, acc[ci * 4 + 3] + bv));
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
                lap(5);
                if (!grid_barrier(P, epoch)) { dead = true; break; }
                lap(6);
            }
            }
            if (!gen || dead) continue;

            // ---------------- sample-level tier + head + sampler: cluster-local, 8 prompts per group ----------------
            const Tier& TL = P.tiers[P.n_ft - 1];
            const float* condL = TL.obuf + (size_t)(t % TL.fs) * H * Bp;     // outputs[-1][:, (t % fs[-2]) - fs[-2]]
            const float* W1s = w_s + P.so_w1;       // [KS][Hh]  rows rank*KS .. of x
            const float* b1s = w_s + P.so_b1;       // [RS]      biases of this CTA's hidden rows
            const float* W2s = w_s + P.so_w2;       // [RS][ZR]  hidden rows rank*RS ..
            float* xh = region + P.h_x;             // [KS][GP]
            float* inbox_h = region + P.h_inh;      // [CS][RS][GP]
            float* hid_s = region + P.h_hid;        // [RS][GP]
            float* part_h = region + P.h_un;        // [nq][tiles][16]   (dead before the logits arrive)
            float* inbox_z = region + P.h_un;       // [GP/CS][CS][ZR]
            float* zs = region + P.h_zs;            // [GP/CS][ZR + 4]
            const float* b2s = smem + P.s_b2;       // [Q + 1] output biases
            unsigned* qbuf = reinterpret_cast<unsigned*>(smem + P.s_bar) + 2 * BAR_COUNT;   // [groups per cluster][GP]
            const int KS = P.KS, RS = P.RS, ZR = P.ZR, Hh = P.Hh, fsl = P.fs_last;
            const long long hstep = t - P.gen_begin, n_gen = P.gen_end - P.gen_begin;
            int gl = 0;
            for (int g = cluster; g < n_groups && !dead; g += n_clusters, ++gl) {
                const int b0 = g * GP;
                const unsigned par = head_phase & 1u;
                hlap(5);
                if (ENGINE == 2 && P.fold_head) {
                    // -- 1'-3'. hidden rows of this CTA straight from the folded up-sampler output: mish(pre + (W1 conv_w) lin(q))
                    const float* preL = P.pre + (size_t)(t % TL.fs) * Bp * Hh;
                    for (int i = tid; i < RS * GP; i += NTK) {
                        const int r = i / GP, p = i - r * GP, row = rank * RS + r, b = b0 + p;
                        float a = (b < P.B) ? __ldcg(preL + (size_t)b * Hh + row) : 0.0f;
                        for (int f = 0; f < fsl; ++f) {
                            long long q = 0;
                            if (b < P.B) {
                                if (f == fsl - 1 && !P.teacher_forced && t > P.gen_begin) q = (long long)qbuf[gl * GP + p];
                                else q = __ldcg(P.seq + (size_t)b * P.seq_stride + (tw - fsl + f));
                            }
                            a = fmaf(linearize(q, Qf), __ldg(P.hu + (size_t)row * fsl + f), a);
                        }
                        hid_s[i] = mish_acc(a);
                    }
                    __syncthreads();
                } else {
                // -- 1. this CTA's rows of x = Conv1d(lin(q[t-fs:t])) + conditioning (modules/io.py:185-198)
                for (int idx = tid; idx < KS * GP; idx += NTK) {
                    const int kl = FAST ? idx % KS : idx / GP, p = FAST ? idx / KS : idx - kl * GP, k = rank * KS + kl, b = b0 + p;
                    float a = 0.0f;
                    for (int f = 0; f < fsl; ++f) {
                        long long q = 0;
                        if (b < P.B) {
                            if (f == fsl - 1 && !P.teacher_forced && t > P.gen_begin) q = (long long)qbuf[gl * GP + p];
                            else q = __ldcg(P.seq + (size_t)b * P.seq_stride + (tw - fsl + f));
                        }
                        a = fmaf(linearize(q, Qf), __ldg(P.conv_w + (size_t)k * fsl + f), a);
                    }
                    a += __ldg(P.conv_b + k);
                    a += (b < P.B) ? __ldcg(FAST ? condL + (size_t)b * H + k : condL + (size_t)k * Bp + b) : 0.0f;
                    xh[kl * GP + p] = a;
                }
                __syncthreads();
                hlap(6);
                // -- 2. partial hidden over this CTA's K slice, reduce-scattered by hidden row
                {
                    const int tiles = (Hh >> 2) * npq_h;
                    int nq = HT / tiles, p2 = 1;
                    while (p2 * 2 <= nq && p2 * 2 <= KS) p2 *= 2;
                    nq = p2;
                    for (int tb = 0; tb < tiles * nq; tb += NTK) {   // one pass unless the head is very wide
                        const int id = tb + tid;
                        if (id < tiles * nq) {
                            const int tile = id % tiles, s = id / tiles, rq = tile / npq_h, pq = tile - rq * npq_h;
                            float acc[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
#pragma unroll 4
                            for (int kl = s; kl < KS; kl += nq) {
                                const float4 w = *reinterpret_cast<const float4*>(W1s + (size_t)kl * Hh + 4 * rq);
                                const float4 x = *reinterpret_cast<const float4*>(xh + kl * GP + 4 * pq);
                                const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                                    for (int pi = 0; pi < 4; ++pi) acc[ci * 4 + pi] = fmaf(wv[ci], xv[pi], acc[ci * 4 + pi]);
                            }
                            float4* d = reinterpret_cast<float4*>(part_h + ((size_t)s * tiles + tile) * 16);
#pragma unroll
                            for (int ci = 0; ci < 4; ++ci) d[ci] = make_float4(acc[ci * 4], acc[ci * 4 + 1], acc[ci * 4 + 2], acc[ci * 4 + 3]);
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < Hh * npq_h; i += NTK) {
                        const int row = i / npq_h, half = i - row * npq_h, tile = (row >> 2) * npq_h + half;
                        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        for (int s = 0; s < nq; ++s) {
                            const float4 a = *reinterpret_cast<const float4*>(part_h + ((size_t)s * tiles + tile) * 16 + (row & 3) * 4);
                            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                        }
                        const unsigned dst = (unsigned)(row / RS);
                        const unsigned win = mapa(sbase, dst) - sbase;
                        const unsigned off = (unsigned)(P.s_region + P.h_inh + ((rank * RS + row % RS) * GP + half * 4)) * 4u;
                        st_async_v4(win + sbase + off, v, win + bar(BAR_HID));
                    }
                }
                hlap(7);
                dead |= !mbar_wait(bar(BAR_HID), par, abort_flag);
                if (tid == 0) mbar_expect_tx(bar(BAR_HID), hid_bytes);
                hlap(8);
                // -- 3. hidden rows of this CTA: sum the partials in rank order, bias, Mish (mlp.py:44-53)
                for (int i = tid; i < RS * GP; i += NTK) {
                    float s = 0.0f;
                    for (int src = 0; src < CS; ++src) s += inbox_h[src * RS * GP + i];
                    hid_s[i] = mish_acc(s + b1s[i / GP]);
                }
                __syncthreads();
                }
                hlap(0);
                // -- 4. partial logits over this CTA's hidden rows, reduce-scattered by prompt
                {
                    const int tiles = (ZR >> 2) * npq_h;
                    for (int tile = tid; tile < tiles; tile += NTK) {
                        const int rq = tile / npq_h, pq = tile - rq * npq_h;
                        float acc[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
#pragma unroll 4
                        for (int k = 0; k < RS; ++k) {
                            const float4 w = *reinterpret_cast<const float4*>(W2s + (size_t)k * ZR + 4 * rq);
                            const float4 x = *reinterpret_cast<const float4*>(hid_s + k * GP + 4 * pq);
                            const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                            for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                                for (int pi = 0; pi < 4; ++pi) acc[ci * 4 + pi] = fmaf(wv[ci], xv[pi], acc[ci * 4 + pi]);
                        }
#pragma unroll
                        for (int pi = 0; pi < 4; ++pi) {
                            const int p = 4 * pq + pi;
                            const unsigned dst = (unsigned)(p % CS), slot = (unsigned)(p / CS);
                            const unsigned win = mapa(sbase, dst) - sbase;
                            const unsigned off = (unsigned)(P.s_region + P.h_un + ((slot * CS + rank) * ZR + 4 * rq)) * 4u;
                            st_async_v4(win + sbase + off, make_float4(acc[pi], acc[4 + pi], acc[8 + pi], acc[12 + pi]),
                                        win + bar(BAR_Z));
                        }
                    }
                }
                hlap(1);
                dead |= !mbar_wait(bar(BAR_Z), par, abort_flag);
                if (tid == 0) mbar_expect_tx(bar(BAR_Z), z_bytes);
                hlap(2);
                // -- 5a. every thread: one logit of one of this CTA's prompts — partial sums in rank order, bias, learned temperature
                //        (mlp.py:60-62; the temperature row is re-derived per thread: 4 loads and a sigmoid instead of a barrier)
                {
                    const int D = GP / CS;
                    for (int idx = tid; idx < D * P.Q; idx += NTK) {
                        const int w = idx / P.Q, o = idx - w * P.Q;
                        const float* iz = inbox_z + (size_t)(w * CS) * ZR;
                        float sv = 0.0f, st = 0.0f;
                        for (int src = 0; src < CS; ++src) { sv += iz[(size_t)src * ZR + o]; st += iz[(size_t)src * ZR + P.Q]; }
                        const float temp = fmaxf(sigmoid_acc(st + b2s[P.Q]), P.min_temp);
                        const float v = (sv + b2s[o]) / temp;
                        zs[w * (ZR + 4) + o] = v;
                        const int b = b0 + w * CS + rank;
                        if (P.logits_out && b < P.B) __stcs(P.logits_out + ((size_t)b * n_gen + hstep) * P.Q + o, v);
                    }
                    __syncthreads();
                }
                // -- 5b. one warp per prompt: the decision
                if (warp < GP / CS) {
                    const int p = warp * CS + rank, b = b0 + p;
                    float* zr = zs + warp * (ZR + 4);
                    int choice = 0;
                    if (b < P.B) {
                        const bool sample = P.temperature != nullptr;
                        float Tt = 1.0f, u = 0.0f;
                        if (sample) {
                            Tt = P.temperature[P.n_temperature == 1 ? 0 : b];
                            u = P.noise[(size_t)b * P.noise_stride + (t - P.noise_t0)];
                        }
                        choice = mmk::decide_scaled_warp(zr, P.Q, sample, Tt, u);
                        if (lane == 0) {
                            if (P.decisions) P.decisions[(size_t)b * n_gen + hstep] = choice;
                            if (!P.teacher_forced) __stcg(P.seq + (size_t)b * P.seq_stride + t, (long long)choice);
                        }
                    }
                    // broadcast the index to every CTA of the cluster (their next x needs it)
                    if (lane < CS) {
                        const unsigned win = mapa(sbase, (unsigned)lane) - sbase;
                        const unsigned off = off_bar0 + (unsigned)(2 * BAR_COUNT + gl * GP + p) * 4u;
                        st_async_u32(win + sbase + off, (unsigned)choice, win + bar(BAR_Q));
                    }
                }
                hlap(3);
                dead |= !mbar_wait(bar(BAR_Q), par, abort_flag);
                if (tid == 0) mbar_expect_tx(bar(BAR_Q), q_bytes);
                ++head_phase;
                if (__syncthreads_or(dead ? 1 : 0)) dead = true;
                hlap(4);
            }
            heads_pending = true;
            lap(7);
            if (ENGINE == 1) fast_lap(F, 10);
            if (c == 0 && tid == 0 && P.step_ts) P.step_ts[hstep] = globaltimer();
        }
    }
    if (P.dbg && c == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) P.dbg[i] += tacc[i];
    if (ENGINE == 1 && F.timing)
        for (int i = 0; i < 12; ++i) P.dbg[8 + i] += (unsigned long long)F.ta[i];
    if (FAST && P.dbg && c == 0 && tid == 0)
        for (int i = 0; i < 9; ++i) P.dbg[20 + i] += (unsigned long long)hta[i];
    // no CTA may exit while peers can still store into its shared memory
    if (ENGINE == 2) tc_fence_before();
    __syncthreads();
    if (ENGINE == 2 && warp == 4) { tc_fence_after(); tmem_dealloc(X.tmem, 64); }
    cluster_sync_all();
}

static int pad4(int v) { return (v + 3) / 4 * 4; }

}  // namespace mmk_sr2

using namespace mmk_sr2;

struct sr2_handle {
    Params p{};
    int device = 0, max_batch = 0, rf = 0;
    size_t smem_bytes = 0;
    std::vector<void*> allocs;
    size_t hbuf_floats[MAX_TIERS] = {0};
};

int sr2_destroy(sr2_handle* h) {
    if (!h) return 0;
    for (void* a : h->allocs) cudaFree(a);
    delete h;
    return 0;
}

static const void* sr2_kernel(int engine) {
    return engine == 2 ? (const void*)samplernn_cluster_kernel<2> : engine == 1 ? (const void*)samplernn_cluster_kernel<1> : (const void*)samplernn_cluster_kernel<0>;
}
static int sr2_max_clusters(int CS, size_t smem, int sms, int engine) {
    const void* k = sr2_kernel(engine);
    const int NTH = engine ? NTF : NT;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (CS == 1) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, NTH, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        return per_sm * sms;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS * 4);
    cfg.blockDim = dim3(NTH);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}


// ---- lane-major engine: plan (shared-memory map, cluster size) and per-lane weight packing
static bool fast_supported(const mmk_samplernn_desc* d, int sms, bool tc) {
    const int n_ft = d->n_tiers - 1, H = d->hidden_dim;
    if (getenv("MMK_SR_ENGINE") && atoi(getenv("MMK_SR_ENGINE")) == 2) return false;     // 2: the tile engine
    if (H % 128 != 0 || (H != 128 && H != 256 && H != 512) || H / 4 > sms || d->head_hidden % 4 != 0 || n_ft > MAX_TIERS) return false;
    for (int i = 0; i < n_ft; ++i) {
        const int fs = d->frame_sizes[i], nxt = i < n_ft - 1 ? d->frame_sizes[i + 1] : 1;
        // both engines also host the reference's default geometry (16, 8, 8): frames of 16, up-sampling by 8 — the tensor-core one only
        // when the bottom tier's up-sampler is folded with the head's first Linear (16 columns per CTA instead of 32)
        if (fs != 1 && fs != 2 && fs != 4 && fs != 8 && fs != 16) return false;
        if (fs % nxt != 0) return false;
        const int up = fs / nxt;
        if (up != 1 && up != 2 && up != 4 && up != 8) return false;
        if (up == 8 && tc) {
            const int rows = up * d->head_hidden, NC = H / 4;
            const bool fold = i == n_ft - 1 && !(getenv("MMK_SR_FOLD") && atoi(getenv("MMK_SR_FOLD")) == 0) && rows % NC == 0 && rows / NC <= 16 &&
                              d->head_hidden % (rows / NC) == 0;
            if (!fold) return false;
        }
    }
    return true;
}

static bool plan_fast(const mmk_samplernn_desc* d, int max_batch, int CS, int sms, int max_optin, bool tc, Params* out, size_t* out_smem) {
    const int n_ft = d->n_tiers - 1, H = d->hidden_dim, Hh = d->head_hidden, Q = d->q_levels, NC = H / 4;
    if (NC % CS || H % CS || Hh % CS) return false;
    Params p{};
    const int GP = CS == 8 ? 8 : 4;
    p.fast = 1; p.NKQ = H / 128; p.CHP = 16 / p.NKQ;
    p.tc = tc ? 1 : 0;
    if (tc && max_batch > 128) return false;                 // one M = 128 tile of prompts
    p.n_ft = n_ft; p.H = H; p.Hh = Hh; p.Q = Q; p.NC = NC; p.CS = CS; p.GP = GP;
    p.JP = 4; p.NG = 12;
    p.fs_last = d->frame_sizes[n_ft]; p.fs_lt = d->frame_sizes[n_ft - 1];
    p.KS = H / CS; p.RS = Hh / CS; p.ZR = pad4(Q + 1);
    p.min_temp = d->min_temperature;
    p.Bp = tc ? 128 : (max_batch + p.CHP - 1) / p.CHP * p.CHP;
    int o = 0, fs_max = 1;
    auto take = [&](int floats) { int r = o; o += pad4(floats); return r; };
    for (int i = 0; i < n_ft; ++i) {
        Tier& T = p.tiers[i];
        T.fs = d->frame_sizes[i];
        T.up = T.fs / (i < n_ft - 1 ? d->frame_sizes[i + 1] : 1);
        T.kdiv = i > 0 ? d->frame_sizes[i - 1] / T.fs : 1;
        T.up_rows = T.up * H;
        T.NV = 4 * T.up; T.NU = T.NV;
        T.so_wih = T.so_whh = T.so_wup = T.so_inw = -1;
        fs_max = std::max(fs_max, T.fs);
    }
    p.off_w1 = take(p.KS * Hh);
    p.off_b1 = take(p.RS);
    p.off_w2 = take(p.RS * p.ZR);
    p.cta_block = o;
    // head buffers
    int ho = 0;
    auto htake = [&](int floats) { int r = ho; ho += pad4(floats); return r; };
    p.h_x = htake(p.KS * GP);
    p.h_inh = htake(Hh * GP);
    p.h_hid = htake(p.RS * GP);
    p.h_un = ho;
    const int tiles_h = (Hh / 4) * (GP / 4);
    int nq = std::max(1, HT / tiles_h), p2 = 1;
    while (p2 * 2 <= nq && p2 * 2 <= p.KS) p2 *= 2;
    const int part_floats = tiles_h * p2 * 16;
    const int inz_floats = GP * p.ZR;
    p.h_zs = p.h_un + pad4(inz_floats);
    const int head_floats = ho + std::max(part_floats, pad4(inz_floats) + (GP / CS) * (p.ZR + 4));
    const int budget = max_optin / (int)sizeof(float) - 256;
    // tensor-core engine: which folds apply (MMK_SR_FOLD=0: none)
    const bool folds = tc && !(getenv("MMK_SR_FOLD") && atoi(getenv("MMK_SR_FOLD")) == 0);
    {
        const int rows = p.tiers[n_ft - 1].up * Hh;           // the head's first Linear into the bottom tier's up-sampler
        if (folds && rows % NC == 0 && rows / NC >= 1 && rows / NC <= 16 && Hh % (rows / NC) == 0) { p.fold_head = 1; p.NVh = rows / NC; }
    }
    bool gx_ok = folds && n_ft >= 2;                           // a tier's up-sampler into the input side of the tier below
    for (int i = 0; i + 1 < n_ft; ++i) gx_ok = gx_ok && p.tiers[i].up <= (int)(TC_BX_BYTES / TC_B_BYTES);
    for (int attempt = gx_ok ? 0 : 1; attempt < 2; ++attempt) {
        p.fold_gx = attempt == 0 ? 1 : 0;
        p.tc_bbytes = p.fold_gx ? (int)TC_BX_BYTES : (int)TC_B_BYTES;
        o = 0;
        // the stages (1024-byte aligned inside the region: 1 KB of slack) share the region with the head buffers
        p.region = tc ? std::max(head_floats, (int)(TC_STAGES * (TC_A_BYTES + (unsigned)p.tc_bbytes) + 1024) / 4) : head_floats;
        p.xregion = p.wregion = p.xstage = p.wstage = 0;
        p.s_region = take(p.region);
        p.s_gi = o;
        p.s_bar = take(2 * BAR_COUNT + 16 * GP);
        p.s_b2 = take(Q + 1);
        int nv_max = 16;
        for (int i = 0; i < n_ft; ++i) nv_max = std::max(nv_max, p.tiers[i].NV);
        p.s_part = take(tc ? 4 : p.NKQ * p.Bp * (nv_max > 16 ? 32 : 16));
        p.s_hold = take(tc ? 4 : p.Bp * 4);
        p.s_lin = take(p.Bp * fs_max);
        p.s_tcbar = take(2 * TCB_COUNT + 4);
        p.n_res = 0;
        bool ok = true;
        auto resident = [&](int goff, int len, int* soff) {
            len = pad4(len);
            if (o + len > budget) { ok = false; return; }
            *soff = o;
            p.res_goff[p.n_res] = goff; p.res_soff[p.n_res] = o; p.res_len[p.n_res] = len; ++p.n_res;
            o += len;
        };
        if (p.fold_head) p.so_w1 = p.so_b1 = 0;              // the folded head never touches W1 / b1
        else { resident(p.off_w1, p.KS * Hh, &p.so_w1); resident(p.off_b1, p.RS, &p.so_b1); }
        resident(p.off_w2, p.RS * p.ZR, &p.so_w2);
        if (!ok) continue;
        p.smem_floats = o;
        const size_t smem = (size_t)o * sizeof(float);
        const int max_clusters = sr2_max_clusters(CS, smem, sms, tc ? 2 : 1);
        if (getenv("MMK_SR_DEBUG"))
            fprintf(stderr, "[sr2] %s engine: CS=%d NC=%d GP=%d smem=%zu max_clusters=%d fold_head=%d fold_gx=%d\n", tc ? "tcgen05" : "lane-major", CS, NC, GP, smem,
                    max_clusters, p.fold_head, p.fold_gx);
        if (max_clusters * CS < NC) continue;
        *out = p; *out_smem = smem;
        return true;
    }
    return false;
}

// Per-lane packing.  Thread t = 32 q + l of a weight row owns k = 4 t .. 4 t + 3 (q = K quarter, l = lane).
//   GRU, CTA c, slot a of lane l: hidden index 4 c + (a ^ jm(l)), jm(l) = (lane bit 16) * 2 + (lane bit 8) — the two
//   hidden-index folds of the shuffle tree keep the lower half on every lane; weights stored as the FFMA2 operand pairs.
//   up-sampler, CTA c, slot i: row NV c + (i ^ fast_col<NV>(l)); pairs of adjacent slots.
template <int NV>
static void pack_up(float* dst, const float* up_w, int c, int H) {
    const int TW = H / 4;
    for (int t = 0; t < TW; ++t)
        for (int e = 0; e < 2 * NV; ++e) {                      // pair e = i2 * 4 + k: (W[slot 2 i2][k], W[slot 2 i2 + 1][k]); float4 e / 2 of the lane
            const int i2 = e / 4, k = e % 4, m = fast_col<NV>(t & 31);
            float* o = dst + ((size_t)(e / 2) * TW + t) * 4 + (e & 1) * 2;
            o[0] = up_w[(size_t)(c * NV + ((2 * i2) ^ m)) * H + 4 * t + k];
            o[1] = up_w[(size_t)(c * NV + ((2 * i2 + 1) ^ m)) * H + 4 * t + k];
        }
}

int sr2_create(const mmk_samplernn_desc* d, int max_batch, int tc, int lstm, int need_set_hidden, sr2_handle** out, int* unsupported) {
    *unsupported = 1;
    const int n_ft = d->n_tiers - 1, H = d->hidden_dim, Hh = d->head_hidden, Q = d->q_levels;
    if (H % KC != 0 || Hh % 4 != 0 || n_ft > MAX_TIERS) return 1;
    int dev = 0, sms = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const char* force_cs = getenv("MMK_SR_CLUSTER");
    const char* force_nc = getenv("MMK_SR_CTAS");

    auto* h = new sr2_handle();
    h->device = dev;
    Params best{};
    size_t best_smem = 0;
    bool found = false;
    const int groups_per_cluster_max = 16;
    if (fast_supported(d, sms, tc != 0))
        for (int CS : {4, 8, 2, 1}) {
            if (force_cs && atoi(force_cs) != CS) continue;
            if (plan_fast(d, max_batch, CS, sms, max_optin, tc != 0, &best, &best_smem)) { found = true; break; }
        }
    if (tc && !found) { sr2_destroy(h); return 1; }            // the tensor-core engine hosts H in {128, 256, 512}, <= 128 prompts
    if ((lstm || need_set_hidden) && !found) { sr2_destroy(h); return 1; }   // LSTM tiers / a non-zero initial state: never the tile engine
    best.lstm = lstm ? 1 : 0;
    for (int CS : {8, 4, 2, 1}) {
        if (found) break;
        if (force_cs && atoi(force_cs) != CS) continue;
        if (H % CS || Hh % CS) continue;
        // NC: the largest multiple of CS that divides H (GRU rows split evenly by hidden index) and fits the device
        int NC = 0;
        for (int n = std::min(sms, H) / CS * CS; n >= CS; n -= CS)
            if (H % n == 0 && (!force_nc || n <= atoi(force_nc))) { NC = n; break; }
        if (!NC) continue;
        Params p{};
        const int GP = CS == 8 ? 8 : 4;
        p.n_ft = n_ft; p.H = H; p.Hh = Hh; p.Q = Q; p.NC = NC; p.CS = CS; p.GP = GP;
        p.JP = H / NC; p.NG = pad4(3 * p.JP);
        p.fs_last = d->frame_sizes[n_ft]; p.fs_lt = d->frame_sizes[n_ft - 1];
        p.KS = H / CS; p.RS = Hh / CS; p.ZR = pad4(Q + 1);
        p.min_temp = d->min_temperature;
        // ---- the CTA's packed block in global memory holds everything
        int o = 0, fs_max = 1;
        auto take = [&](int floats) { int r = o; o += pad4(floats); return r; };
        bool ok = true;
        for (int i = 0; i < n_ft; ++i) {
            Tier& T = p.tiers[i];
            T.fs = d->frame_sizes[i];
            T.up = T.fs / (i < n_ft - 1 ? d->frame_sizes[i + 1] : 1);     // sample_rnn_v2.py:155-158
            T.kdiv = i > 0 ? d->frame_sizes[i - 1] / T.fs : 1;
            T.up_rows = T.up * H;
            if (T.up_rows % NC) ok = false;
            T.NU = pad4(T.up_rows / NC);
            T.off_wih = take(H * p.NG + p.NG);
            T.off_whh = take(H * p.NG + p.NG);
            T.off_wup = take(H * T.NU + T.NU);
            T.off_inw = take(H * T.fs + H);
            T.so_wih = T.so_whh = T.so_wup = T.so_inw = -1;
            fs_max = std::max(fs_max, T.fs);
            if ((T.NU >> 2) > NT || (p.NG >> 2) > NT) ok = false;
        }
        if (!ok) continue;
        p.off_w1 = take(p.KS * Hh);
        p.off_b1 = take(p.RS);
        p.off_w2 = take(p.RS * p.ZR);
        p.cta_block = o;
        // ---- shared memory: buffers first, then the resident weights by priority until the budget is spent
        o = 0;
        int nstage = NSTAGE_DEFAULT;
        int kcfull = KCFULL_DEFAULT;
        int wmax = p.NG;
        for (int i = 0; i < n_ft; ++i) wmax = std::max(wmax, p.tiers[i].NU);
        p.xstage = kcfull * PBW; p.wstage = kcfull * std::min(WMAX, wmax);
        p.xregion = nstage * p.xstage; p.wregion = nstage * p.wstage; p.region = p.xregion + p.wregion;
        const int REGION = p.region;
        p.s_region = take(REGION);
        p.s_gi = take(std::max(p.NG, fs_max) * PBW);
        p.s_bar = take(2 * BAR_COUNT + groups_per_cluster_max * GP);
        p.s_b2 = take(Q + 1);
        p.n_res = 0;
        const int budget = max_optin / (int)sizeof(float) - 256;     // floats; 1 KB of slack for static shared memory
        auto resident = [&](int goff, int len, int* soff) {
            len = pad4(len);
            if (o + len > budget) return false;
            *soff = o;
            p.res_goff[p.n_res] = goff; p.res_soff[p.n_res] = o; p.res_len[p.n_res] = len; ++p.n_res;
            o += len;
            return true;
        };
        // must be resident: the head slices (used every sample) and the frame Linear tables
        ok = resident(p.off_w1, p.KS * Hh, &p.so_w1) && resident(p.off_b1, p.RS, &p.so_b1) &&
             resident(p.off_w2, p.RS * p.ZR, &p.so_w2);
        for (int i = n_ft - 1; i >= 0 && ok; --i)
            ok = resident(p.tiers[i].off_inw, H * p.tiers[i].fs + H, &p.tiers[i].so_inw);
        if (!ok) continue;
        // optional, most frequently firing tier first (input-side GRU matrix, hidden-side, up-sampler); the rest
        // streams from L2 next to the activations
        for (int i = n_ft - 1; i >= 0; --i) {
            Tier& T = p.tiers[i];
            if (!resident(T.off_wih, H * p.NG + p.NG, &T.so_wih) && p.NG > WMAX) ok = false;
            if (!resident(T.off_whh, H * p.NG + p.NG, &T.so_whh) && p.NG > WMAX) ok = false;
            if (!resident(T.off_wup, H * T.NU + T.NU, &T.so_wup) && T.NU > WMAX) ok = false;
        }
        if (!ok) continue;
        p.smem_floats = o;
        // head buffers inside the region
        int ho = 0;
        auto htake = [&](int floats) { int r = ho; ho += pad4(floats); return r; };
        p.h_x = htake(p.KS * GP);
        p.h_inh = htake(Hh * GP);
        p.h_hid = htake(p.RS * GP);
        p.h_un = ho;
        const int tiles_h = (Hh / 4) * (GP / 4);
        int nq = std::max(1, HT / tiles_h), p2 = 1;
        while (p2 * 2 <= nq && p2 * 2 <= p.KS) p2 *= 2;
        const int part_floats = tiles_h * p2 * 16;
        const int inz_floats = GP * p.ZR;                      // (GP/CS) slots x CS sources
        p.h_zs = p.h_un + pad4(inz_floats);
        const int un_floats = std::max(part_floats, pad4(inz_floats) + (GP / CS) * (p.ZR + 4));
        if (ho + un_floats > REGION || p.NG * PBW * 2 > REGION) continue;
        const size_t smem = (size_t)o * sizeof(float);
        const int max_clusters = sr2_max_clusters(CS, smem, sms, 0);
        if (getenv("MMK_SR_DEBUG")) {
            fprintf(stderr, "[sr2] CS=%d NC=%d GP=%d smem=%zu max_clusters=%d resident:", CS, NC, GP, smem, max_clusters);
            for (int i = 0; i < n_ft; ++i) fprintf(stderr, " t%d(ih=%d hh=%d up=%d)", i, p.tiers[i].so_wih >= 0, p.tiers[i].so_whh >= 0, p.tiers[i].so_wup >= 0);
            fprintf(stderr, "\n");
        }
        if (max_clusters * CS < NC) continue;   // the grid could not be co-resident with this cluster size
        best = p; best_smem = smem; found = true;
        break;
    }
    if (!found) { sr2_destroy(h); return 1; }
    *unsupported = 0;
    Params& p = h->p;
    p = best;
    h->smem_bytes = best_smem;
    h->max_batch = max_batch; h->rf = d->frame_sizes[0];
    if (!p.fast) p.Bp = (max_batch + 3) / 4 * 4;
    const int NC = p.NC, CS = p.CS, GP = p.GP;
    MMK_CUDA(cudaFuncSetAttribute(sr2_kernel(p.tc ? 2 : p.fast), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    {
        const int n_groups = (max_batch + GP - 1) / GP, n_clusters = NC / CS;
        if ((n_groups + n_clusters - 1) / n_clusters > groups_per_cluster_max) { sr2_destroy(h); *unsupported = 1; return 1; }
    }

    // ---- pack per-CTA weight blocks
    std::vector<float> wpack((size_t)NC * p.cta_block, 0.0f);
    for (int c = 0; c < NC; ++c) {
        float* blk = wpack.data() + (size_t)c * p.cta_block;
        const int j_lo = c * p.JP, rank = c % CS;
        for (int i = 0; i < n_ft && !p.fast; ++i) {
            const Tier& T = p.tiers[i];
            for (int m = 0; m < 2; ++m) {
                const float* W = m == 0 ? d->w_ih[i] : d->w_hh[i];
                const float* b = m == 0 ? d->b_ih[i] : d->b_hh[i];
                float* Ws = blk + (m == 0 ? T.off_wih : T.off_whh);
                float* bs = Ws + (size_t)H * p.NG;
                for (int g = 0; g < 3; ++g)
                    for (int jj = 0; jj < p.JP; ++jj) {
                        const int col = g * p.JP + jj, row = g * H + j_lo + jj;
                        for (int k = 0; k < H; ++k) Ws[(size_t)k * p.NG + col] = W[(size_t)row * H + k];
                        bs[col] = b[row];
                    }
            }
            std::copy(d->in_w[i], d->in_w[i] + (size_t)H * T.fs, blk + T.off_inw);
            std::copy(d->in_b[i], d->in_b[i] + H, blk + T.off_inw + (size_t)H * T.fs);
            const int nu = T.up_rows / NC, u_lo = c * nu;
            float* Wu = blk + T.off_wup; float* bu = Wu + (size_t)H * T.NU;
            for (int col = 0; col < nu; ++col) {
                for (int k = 0; k < H; ++k) Wu[(size_t)k * T.NU + col] = d->up_w[i][(size_t)(u_lo + col) * H + k];
                bu[col] = d->up_b[i][u_lo + col];
            }
        }
        // head: K slices.  W1s[kl][row] = W1[row][rank*KS + kl];  W2s[kl][o] = W2[o][rank*RS + kl]
        float* W1s = blk + p.off_w1; float* b1s = blk + p.off_b1; float* W2s = blk + p.off_w2;
        for (int kl = 0; kl < p.KS; ++kl)
            for (int row = 0; row < Hh; ++row) W1s[(size_t)kl * Hh + row] = d->head_w1[(size_t)row * H + rank * p.KS + kl];
        for (int r = 0; r < p.RS; ++r) b1s[r] = d->head_b1[rank * p.RS + r];
        for (int kl = 0; kl < p.RS; ++kl)
            for (int o2 = 0; o2 <= Q; ++o2) W2s[(size_t)kl * p.ZR + o2] = d->head_w2[(size_t)o2 * Hh + rank * p.RS + kl];
    }
    auto dev_alloc = [&](size_t bytes, const void* src) -> void* {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, bytes) != cudaSuccess) return nullptr;
        h->allocs.push_back(ptr);
        if (src) cudaMemcpy(ptr, src, bytes, cudaMemcpyHostToDevice); else cudaMemset(ptr, 0, bytes);
        return ptr;
    };
    bool ok = true;
    auto up = [&](const float* src, size_t n) { void* q = dev_alloc(n * sizeof(float), src); ok = ok && q; return (float*)q; };
    p.wpack = up(wpack.data(), wpack.size());
    for (int i = 0; i < n_ft; ++i) {
        Tier& T = p.tiers[i];
        T.in_w = up(d->in_w[i], (size_t)H * T.fs);
        T.in_b = up(d->in_b[i], H);
        h->hbuf_floats[i] = (size_t)2 * H * p.Bp;
        T.hbuf = up(nullptr, h->hbuf_floats[i]);
        T.obuf = up(nullptr, (size_t)T.up_rows * p.Bp);
        if (!p.fast) continue;
        const int TW = H / 4;
        const int NWG = p.lstm ? 32 : 24;                       // float4 of recurrent weights per lane
        std::vector<float> wg((size_t)NC * NWG * TW * 4), wu((size_t)NC * T.NV * TW * 4), iw((size_t)T.fs * TW * 4), gb((size_t)NC * NWG),
            ub((size_t)NC * T.NV);
        for (int c = 0; c < NC && p.lstm; ++c) {
            // LSTM: 64 pairs per lane, e = m * 32 + i2 * 4 + k: (W_m[slot 2 i2][k], W_m[slot 2 i2 + 1][k]); slot i = column i ^ fast_col<16>(l),
            // column = gate * 4 + hidden index (gates i, f, g, o)
            auto row_of = [&](int col) { return (col / 4) * H + 4 * c + (col % 4); };
            for (int t = 0; t < TW; ++t) {
                const int mask = fast_col<16>(t & 31);
                for (int e = 0; e < 64; ++e) {
                    const int m = e / 32, r = e % 32, i2 = r / 4, k = r % 4;
                    const float* W = m == 0 ? d->w_ih[i] : d->w_hh[i];
                    float* dst = wg.data() + (((size_t)c * 32 + e / 2) * TW + t) * 4 + (e & 1) * 2;
                    dst[0] = W[(size_t)row_of((2 * i2) ^ mask) * H + 4 * t + k];
                    dst[1] = W[(size_t)row_of((2 * i2 + 1) ^ mask) * H + 4 * t + k];
                }
            }
            for (int col = 0; col < 16; ++col) {
                gb[(size_t)c * 32 + col] = d->b_ih[i][row_of(col)];
                gb[(size_t)c * 32 + 16 + col] = d->b_hh[i][row_of(col)];
            }
        }
        for (int c = 0; c < NC && !p.lstm; ++c) {
            // 48 pairs per lane, e = 0..47: [0,16) (W_ir, W_iz)[slot e / 4][k = e % 4], [16,24) (W_in[slot 2a'], W_in[slot 2a'+1])[k],
            // [24,48) the same of W_hh; slot a of lane l is hidden index 4c + (a ^ jm(l)).  Pair e lives in float4 e / 2 of the lane.
            for (int t = 0; t < TW; ++t) {
                const int l = t & 31, jm = ((l >> 4) & 1) << 1 | ((l >> 3) & 1);
                for (int e = 0; e < 48; ++e) {
                    const int m = e / 24, r = e % 24, k = r % 4;
                    const float* W = m == 0 ? d->w_ih[i] : d->w_hh[i];
                    int row_lo, row_hi;
                    if (r < 16) { const int a = r / 4; row_lo = 0 * H + 4 * c + (a ^ jm); row_hi = 1 * H + 4 * c + (a ^ jm); }
                    else { const int a2 = (r - 16) / 4; row_lo = 2 * H + 4 * c + ((2 * a2) ^ jm); row_hi = 2 * H + 4 * c + ((2 * a2 + 1) ^ jm); }
                    float* dst = wg.data() + (((size_t)c * 24 + e / 2) * TW + t) * 4 + (e & 1) * 2;
                    dst[0] = W[(size_t)row_lo * H + 4 * t + k];
                    dst[1] = W[(size_t)row_hi * H + 4 * t + k];
                }
            }
            for (int m = 0; m < 2; ++m) {
                const float* b = m == 0 ? d->b_ih[i] : d->b_hh[i];
                for (int g = 0; g < 3; ++g)
                    for (int a = 0; a < 4; ++a) gb[(size_t)c * 24 + m * 12 + g * 4 + a] = b[g * H + 4 * c + a];
            }
        }
        for (int c = 0; c < NC; ++c) {
            float* wuc = wu.data() + (size_t)c * T.NV * TW * 4;
            if (T.NV == 4) pack_up<4>(wuc, d->up_w[i], c, H);
            else if (T.NV == 8) pack_up<8>(wuc, d->up_w[i], c, H);
            else if (T.NV == 16) pack_up<16>(wuc, d->up_w[i], c, H);
            else pack_up<32>(wuc, d->up_w[i], c, H);
            for (int col = 0; col < T.NV; ++col) ub[(size_t)c * T.NV + col] = d->up_b[i][c * T.NV + col];
        }
        for (int f = 0; f < T.fs; ++f)
            for (int k = 0; k < H; ++k) iw[(size_t)f * H + k] = d->in_w[i][(size_t)k * T.fs + f];
        T.wg4 = reinterpret_cast<const float4*>(up(wg.data(), wg.size()));
        T.wu4 = reinterpret_cast<const float4*>(up(wu.data(), wu.size()));
        T.iw4 = reinterpret_cast<const float4*>(up(iw.data(), iw.size()));
        T.ib4 = reinterpret_cast<const float4*>(T.in_b);
        T.gb = up(gb.data(), gb.size());
        T.ub = up(ub.data(), ub.size());
        T.cbuf = p.lstm ? up(nullptr, h->hbuf_floats[i]) : nullptr;
        if (!p.tc) continue;
        // ---- tensor-core engine: bf16 weight tiles [16 columns x 128 k] per fill (K = [conditioning | hidden], the top tier: hidden only),
        //      columns gate * 4 + jj with gates r, z, n_i, n_h; the frame Linear and all biases folded into wf / bfold (fp64 -> fp32)
        const int nimg = H / 128, nfill = 2 * nimg;
        // fold the head's first Linear into the bottom tier's up-sampler when its rows split evenly over the CTAs (see Params::fold_head)
        const bool fold_here = p.fold_head && i == n_ft - 1;   // decided by the plan (shared-memory map)
        auto bf = [](float v) { __nv_bfloat16 b = __float2bfloat16_rn(v); unsigned short u; memcpy(&u, &b, 2); return u; };
        std::vector<unsigned char> wgi((size_t)NC * nfill * TC_B_BYTES, 0), wui((size_t)NC * nimg * TC_B_BYTES, 0);
        std::vector<float> wfv((size_t)NC * 16 * T.fs, 0.0f), bfv((size_t)NC * 16, 0.0f);
        const float* Wih = d->w_ih[i]; const float* Whh = d->w_hh[i];
        for (int c = 0; c < NC; ++c) {
            for (int col = 0; col < 16; ++col) {
                const int g4 = col / 4, jj = col % 4, j = 4 * c + jj;
                // row of W_ih / W_hh: GRU (r, z, n; columns n_i and n_h both read row n), LSTM (i, f, g, o)
                const int grow = (p.lstm ? g4 : (g4 == 0 ? 0 : g4 == 1 ? 1 : 2)) * H + j;
                for (int part = 0; part < 2; ++part) {            // 0: conditioning (W_ih), 1: hidden (W_hh)
                    if (part == 0 && i == 0) continue;            // the top tier has no conditioning
                    const bool zero = !p.lstm && ((part == 0 && g4 == 3) || (part == 1 && g4 == 2));
                    const float* W = part == 0 ? Wih : Whh;
                    const int fill0 = (i == 0) ? 0 : part * nimg;
                    for (int k = 0; k < H; ++k) {
                        const unsigned short v = zero ? 0 : bf(W[(size_t)grow * H + k]);
                        unsigned char* tile = wgi.data() + ((size_t)c * nfill + fill0 + k / 128) * TC_B_BYTES;
                        memcpy(tile + tile_off(col, (k % 128) / 8, 16) + (k % 8) * 2, &v, 2);
                    }
                }
                // folded terms (the input x = in_w lin + in_b + conditioning enters W_ih only: columns r, z, n_i)
                double b = 0.0;
                if (g4 < 3 || p.lstm) {
                    b = d->b_ih[i][grow];
                    for (int k = 0; k < H; ++k) b += (double)Wih[(size_t)grow * H + k] * (double)d->in_b[i][k];
                    for (int f = 0; f < T.fs; ++f) {
                        double a = 0.0;
                        for (int k = 0; k < H; ++k) a += (double)Wih[(size_t)grow * H + k] * (double)d->in_w[i][(size_t)k * T.fs + f];
                        wfv[((size_t)c * 16 + col) * T.fs + f] = (float)a;
                    }
                }
                if (g4 < 2 || g4 == 3 || p.lstm) b += d->b_hh[i][grow];
                bfv[(size_t)c * 16 + col] = (float)b;
            }
            if (fold_here) continue;                              // the bottom tier's tiles are packed below from W1 . W_up
            for (int col = 0; col < T.NV; ++col)
                for (int k = 0; k < H; ++k) {
                    const unsigned short v = bf(d->up_w[i][(size_t)(c * T.NV + col) * H + k]);
                    unsigned char* tile = wui.data() + ((size_t)c * nimg + k / 128) * TC_B_BYTES;
                    memcpy(tile + tile_off(col, (k % 128) / 8, 16) + (k % 8) * 2, &v, 2);
                }
        }
        if (fold_here) {
            // W'[slot][j][k] = sum_m W1[j][m] W_up[slot H + m][k];  b'[slot][j] = sum_m W1[j][m] (b_up[slot H + m] + conv_b[m]) + b1[j];
            // hu[j][f] = sum_m W1[j][m] conv_w[m][f]   (fp64 accumulation)
            const int fsl = p.fs_last, NVh = p.NVh;
            std::vector<double> wp((size_t)T.up * Hh * H, 0.0);
            for (int sl = 0; sl < T.up; ++sl)
                for (int j = 0; j < Hh; ++j) {
                    double* row = wp.data() + ((size_t)sl * Hh + j) * H;
                    for (int m = 0; m < H; ++m) {
                        const double w1 = d->head_w1[(size_t)j * H + m];
                        const float* wu = d->up_w[i] + ((size_t)sl * H + m) * H;
                        for (int k = 0; k < H; ++k) row[k] += w1 * (double)wu[k];
                    }
                }
            std::vector<float> hbv((size_t)NC * NVh), huv((size_t)Hh * fsl);
            for (int c = 0; c < NC; ++c)
                for (int col = 0; col < NVh; ++col) {
                    const int urow = c * NVh + col, sl = urow / Hh, j = urow % Hh;
                    for (int k = 0; k < H; ++k) {
                        const unsigned short v = bf((float)wp[((size_t)sl * Hh + j) * H + k]);
                        unsigned char* tile = wui.data() + ((size_t)c * nimg + k / 128) * TC_B_BYTES;
                        memcpy(tile + tile_off(col, (k % 128) / 8, 16) + (k % 8) * 2, &v, 2);
                    }
                    double b = d->head_b1[j];
                    for (int m = 0; m < H; ++m) b += (double)d->head_w1[(size_t)j * H + m] * ((double)d->up_b[i][(size_t)sl * H + m] + (double)d->conv_b[m]);
                    hbv[(size_t)c * NVh + col] = (float)b;
                }
            for (int j = 0; j < Hh; ++j)
                for (int f = 0; f < fsl; ++f) {
                    double a = 0.0;
                    for (int m = 0; m < H; ++m) a += (double)d->head_w1[(size_t)j * H + m] * (double)d->conv_w[(size_t)m * fsl + f];
                    huv[(size_t)j * fsl + f] = (float)a;
                }
            p.hb = up(hbv.data(), hbv.size());
            p.hu = up(huv.data(), huv.size());
            p.pre = up(nullptr, (size_t)T.up * p.Bp * Hh);
        }
        std::vector<unsigned char> wxi;
        std::vector<float> xbv;
        if (p.fold_gx && i < n_ft - 1) {
            // gx of the tier below (i + 1): column (slot s, col) of CTA c = row grow(col) of W_ih(i + 1) applied to slot s of this tier's
            // up-sampler: W''[k] = sum_m W_ih1[grow][m] W_up[s H + m][k];  xb = sum_m W_ih1[grow][m] b_up[s H + m]   (fp64)
            const int ncols = T.up * 16;
            const size_t bfill = (size_t)ncols * 256;
            wxi.assign((size_t)NC * nimg * bfill, 0);
            xbv.assign((size_t)NC * T.up * 16, 0.0f);
            const float* Wih1 = d->w_ih[i + 1];
            std::vector<double> row(H);
            for (int sl = 0; sl < T.up; ++sl)
                for (int c = 0; c < NC; ++c)
                    for (int col = 0; col < 16; ++col) {
                        const int g4 = col / 4, jj = col % 4;
                        if (!p.lstm && g4 == 3) continue;                    // GRU: the n_h column takes no input term
                        const int grow = (p.lstm ? g4 : g4) * H + 4 * c + jj;   // rows r, z, n (GRU) / i, f, g, o (LSTM) of the tier below
                        std::fill(row.begin(), row.end(), 0.0);
                        double bsum = 0.0;
                        for (int m = 0; m < H; ++m) {
                            const double w1 = Wih1[(size_t)grow * H + m];
                            const float* wu = d->up_w[i] + ((size_t)sl * H + m) * H;
                            for (int k = 0; k < H; ++k) row[k] += w1 * (double)wu[k];
                            bsum += w1 * (double)d->up_b[i][(size_t)sl * H + m];
                        }
                        for (int k = 0; k < H; ++k) {
                            const unsigned short v = bf((float)row[k]);
                            unsigned char* tile = wxi.data() + ((size_t)c * nimg + k / 128) * bfill;
                            memcpy(tile + tile_off(sl * 16 + col, (k % 128) / 8, ncols) + (k % 8) * 2, &v, 2);
                        }
                        xbv[((size_t)c * T.up + sl) * 16 + col] = (float)bsum;
                    }
        }
        auto upb = [&](const void* src, size_t bytes) { void* q = dev_alloc(bytes, src); ok = ok && q; return (unsigned char*)q; };
        if (!wxi.empty()) {
            T.wximg = upb(wxi.data(), wxi.size());
            T.xb = up(xbv.data(), xbv.size());
            T.gx = up(nullptr, (size_t)T.up * 128 * NC * 16);
        }
        T.wgimg = upb(wgi.data(), wgi.size());
        T.wuimg = upb(wui.data(), wui.size());
        T.wf = up(wfv.data(), wfv.size());
        T.bfold = up(bfv.data(), bfv.size());
        T.himg = upb(nullptr, (size_t)2 * H * 256);
        T.oimg = i < n_ft - 1 ? upb(nullptr, (size_t)T.up * H * 256) : nullptr;
    }
    p.conv_w = up(d->conv_w, (size_t)H * p.fs_last);
    p.conv_b = up(d->conv_b, H);
    p.b2 = up(d->head_b2, (size_t)Q + 1);
    p.bar = (unsigned long long*)dev_alloc(64, nullptr);
    ok = ok && p.bar;
    if (!ok) { sr2_destroy(h); MMK_FAIL("cudaMalloc failed while creating the SampleRNN handle"); }
    p.abort_flag = (unsigned*)(p.bar + 1);
    if (getenv("MMK_SR_DEBUG")) p.dbg = (unsigned long long*)dev_alloc(256, nullptr);
    if (const char* e = getenv("MMK_SR_EXP")) p.exp = atoi(e);
    MMK_CUDA(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

// A non-zero initial state (SampleRNNTier._init_h0, sample_rnn_v2.py:101-119) for the lane-major / tensor-core engines: rows (B, H)
// fp32 go to the [prompt][H] state buffer as they are; the tensor-core engine also needs them as its bf16 operand image.
__global__ void sr2_image_rows_kernel(unsigned char* img, const float* __restrict__ v, int B, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk (8 k) of one prompt row
    if (i >= B * (H / 8)) return;
    const int p = i / (H / 8), kc = i - p * (H / 8);
    const float* s = v + (size_t)p * H + kc * 8;
    *reinterpret_cast<uint4*>(img + mmk_sr2::tile_off(p, kc, 128)) =
        make_uint4(mmk_sr2::bf16x2_bits(s[0], s[1]), mmk_sr2::bf16x2_bits(s[2], s[3]), mmk_sr2::bf16x2_bits(s[4], s[5]), mmk_sr2::bf16x2_bits(s[6], s[7]));
}

int sr2_set_hidden(sr2_handle* h, int tier, int which, const float* d_values, int B, void* stream) {
    const Params& p = h->p;
    MMK_CHECK(p.fast, "this geometry runs the tile engine of the cluster kernel (zero initial state only)");
    MMK_CHECK(tier >= 0 && tier < p.n_ft, "tier out of range");
    MMK_CHECK(which == 0 || (which == 1 && p.lstm), "which: 0 = hidden state, 1 = LSTM cell state");
    MMK_CHECK(B >= 1 && B <= h->max_batch, "batch exceeds the max_batch the handle was created for");
    cudaStream_t st = (cudaStream_t)stream;
    const Tier& T = p.tiers[tier];
    float* dst = (which == 0 ? T.hbuf : T.cbuf) + (size_t)p.hsel[tier] * p.H * p.Bp;
    MMK_CUDA(cudaMemcpyAsync(dst, d_values, (size_t)B * p.H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (which == 0 && p.tc) {
        unsigned char* img = T.himg + (size_t)p.hsel[tier] * p.H * 256;
        const int n = B * (p.H / 8);
        sr2_image_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(img, d_values, B, p.H);
        MMK_CUDA(cudaGetLastError());
    }
    return 0;
}

int sr2_launch_info(sr2_handle* h, mmk_launch_info* out) {
    out->cluster_size = h->p.CS; out->n_stages = h->p.NC / h->p.CS; out->group_size = h->p.GP; out->threads = h->p.fast ? NTF : NT;
    out->smem_bytes = (int)h->smem_bytes; out->sm_used = h->p.NC;
    return 0;
}

int sr2_sync_check(sr2_handle* h, void* stream) {
    unsigned aborted = 0;
    MMK_CUDA(cudaMemcpyAsync(&aborted, h->p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MMK_CHECK(aborted == 0, "SampleRNN kernel watchdog fired: a barrier wait timed out (results invalid)");
    if (h->p.dbg) {
        unsigned long long t[29];
        MMK_CUDA(cudaMemcpy(t, h->p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
        MMK_CUDA(cudaMemset(h->p.dbg, 0, sizeof(t)));
        if (h->p.fast) {
            fprintf(stderr, "[sr2] CTA0 head Mcycles: hidden rows %.1f (unfolded head: x rows %.1f, W1 partials + send %.1f, wait %.1f, then Mish) logit partials + send %.1f "
                            "wait logits %.1f decide %.1f wait index %.1f | outside the head %.1f\n",
                    (t[20] + t[26] + t[27] + t[28]) / 1e6, t[26] / 1e6, t[27] / 1e6, t[28] / 1e6, t[21] / 1e6, t[22] / 1e6, t[23] / 1e6, t[24] / 1e6, t[25] / 1e6);
            if (h->p.tc) return 0;
            static const char* names[12] = {"weights", "pre-barrier", "frame", "gru stream", "gates", "up weights", "barrier", "up stream", "up rows",
                                            "last barrier", "head", "other"};
            fprintf(stderr, "[sr2] CTA0 Mcycles:");
            for (int i = 0; i < 12; ++i) fprintf(stderr, " %s %.1f", names[i], t[8 + i] / 1e6);
            fprintf(stderr, "\n");
            return 0;
        }
        fprintf(stderr, "[sr2] CTA0 ms: other %.2f pre-barrier %.2f gemm_ih %.2f gemm_hh+gate %.2f barrier %.2f up %.2f barrier %.2f head %.2f\n",
                t[0] / 1e6, t[1] / 1e6, t[2] / 1e6, t[3] / 1e6, t[4] / 1e6, t[5] / 1e6, t[6] / 1e6, t[7] / 1e6);
    }
    return 0;
}

int sr2_run(sr2_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t warm_begin,
            int64_t warm_end, int64_t warm_offset, int64_t gen_begin, int64_t gen_end, int reset_hidden,
            int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (reset_hidden) {
        for (int i = 0; i < h->p.n_ft; ++i) {
            MMK_CUDA(cudaMemsetAsync(h->p.tiers[i].hbuf, 0, h->hbuf_floats[i] * sizeof(float), st));
            if (h->p.tc) MMK_CUDA(cudaMemsetAsync(h->p.tiers[i].himg, 0, (size_t)2 * h->p.H * 256, st));
            if (h->p.tiers[i].cbuf) MMK_CUDA(cudaMemsetAsync(h->p.tiers[i].cbuf, 0, h->hbuf_floats[i] * sizeof(float), st));
            h->p.hsel[i] = 0;
        }
    }
    if (warm_end == warm_begin && gen_end == gen_begin) return 0;
    Params p = h->p;
    MMK_CUDA(cudaMemsetAsync(p.bar, 0, 64, st));
    p.seq = reinterpret_cast<long long*>(d_seq) - seq_t0;
    p.seq_stride = seq_stride;
    p.warm_begin = warm_begin; p.warm_end = warm_end; p.warm_off = warm_offset;
    p.gen_begin = gen_begin; p.gen_end = gen_end;
    p.B = B; p.teacher_forced = teacher_forced ? 1 : 0;
    p.temperature = d_temperature; p.n_temperature = n_temperature;
    p.noise = d_noise; p.noise_stride = noise_stride; p.noise_t0 = noise_t0;
    p.logits_out = d_logits_out; p.decisions = reinterpret_cast<long long*>(d_decisions); p.step_ts = d_step_ts;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.NC);
    cfg.blockDim = dim3(p.fast ? NTF : NT);
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[] = {&p};
    MMK_CUDA(cudaLaunchKernelExC(&cfg, sr2_kernel(p.tc ? 2 : p.fast), args));
    // the hidden ping-pong advances once per tier firing: keep the handle's view in step with the device
    for (int i = 0; i < p.n_ft; ++i) {
        const long long fs = p.tiers[i].fs;
        auto firings = [&](long long lo, long long hi) { return hi > lo ? (hi + fs - 1) / fs - (lo + fs - 1) / fs : 0; };
        const long long n = firings(warm_begin, warm_end) + firings(gen_begin, gen_end);
        h->p.hsel[i] ^= (int)(n & 1);
    }
    return 0;
}
