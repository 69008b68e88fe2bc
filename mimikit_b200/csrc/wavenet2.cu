// Latency-engineered persistent WaveNet generation kernel (sm_100a) — the fast path for channel counts that are
// multiples of 32.  Same function as wavenet.cu (the general kernel); see that file for the reference lines.
//
// The step loop of autoregressive WaveNet is a dependency chain (30 layers x 2 dependent contractions + head +
// sampler per sample), so the design minimises the latency of one (prompt group, layer) unit:
//   * stage = thread-block cluster of CS CTAs holding a contiguous range of layers' fp32 weights in shared memory;
//     every contraction is split by output channel over the CTAs; prompt groups of 8 flow through the stages.
//   * all-gathers inside a cluster (gated output y, next layer input h') are st.async stores into every peer's
//     shared memory that signal the peer's mbarrier (complete_tx); consumers wait on their LOCAL mbarrier.  No
//     cluster barrier, no fence, no CTA barrier on that path.
//   * a contraction is K-split over the 32 lanes of a warp: a warp owns 4 output columns x 8 prompts, each lane
//     accumulates its k = lane + 32 j slice in registers (weights pre-packed so that a lane's 4 columns are one
//     conflict-free LDS.128), and a 31-shuffle butterfly transposes/reduces the 32 partial sums so that lane
//     (c * 8 + p) ends up with output (column c, prompt p).  No shared-memory partials.
//   * the tap-0 half of the dilated conv only needs ring data (h(t - d), fetched with a TMA bulk copy one unit
//     ahead), so it runs on its own warps before the layer input even arrives; the tap-1 half is the only
//     contraction between the arrival of h(t) and the gate.
//   * skip sums stay in shared memory per CTA, residual adds are in-lane, ring writes ride on an idle warp.
#include "common.cuh"
#include "sampler.cuh"
#include "wavenet_impl.h"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace mmk2 {

using mmk::sigmoid_acc;
using mmk::mish_acc;

constexpr int NT = 256;   // threads per CTA
constexpr int NW = NT / 32;
constexpr int GB = 8;     // prompts per pipeline group
constexpr int MAX_LAYERS = 96;
constexpr int MAX_STAGES = 32;
constexpr int TRACE_EV = 40;   // events per (stage, group) in the debug timeline

struct Layer {
    int dilation, has_res;
    long long ring_off;   // float offset of this layer's ring
};

struct Params {
    int L, C, S, Hh, Q, Kh, CS, NST, G;
    int nf, ns, nh, nz;          // per-CTA gate channels, skip channels, head hidden units, head logits (padded)
    int GQ, RQ, SQ, HQ, ZQ;      // quads (4 output columns each)
    int KJ, KJh, KJ2;            // K / 32 of the three contraction depths (C, Kh, Hh)
    int o_g1, o_g0, o_gb, o_r, o_rb, o_s, o_sb, layer_block;   // float offsets inside a (layer, rank) block
    int o_h1, o_h1b, o_h2, o_h2b, head_block;
    int s_w, s_head, s_x1, s_x0, s_y, s_hin, s_hid, s_z, s_zrows, s_gate, s_sacc, s_stage, s_bar, smem_floats;
    int zrow, ring_hazard, blk;  // blk = C * GB floats: one activation block in the [half][C][4] layout
    float min_temp;
    Layer layers[MAX_LAYERS];
    int stage_lo[MAX_STAGES + 1];
    const float* wpack; const float* hpack; const float* E;
    float* rings; float* mail_h; float* mail_s;
    unsigned* ready; unsigned* ack; long long* avail; unsigned* abort_flag;
    // this run
    long long* seq;
    long long seq_stride, t_begin, t_head, t_end;
    int B, n_groups, teacher_forced, n_temperature;
    const float* temperature; const float* noise;
    long long noise_stride, noise_t0;
    float* logits_out; long long* decisions; unsigned long long* step_ts;
    long long* trace; long long trace_t;   // debug timeline (MMK_WN_TRACE_T): clock64 stamps of one step, rank 0 of each stage
};

// barrier slots in shared memory (8 bytes each)
enum { BAR_H0 = 0, BAR_H1, BAR_Y0, BAR_Y1, BAR_R0, BAR_R1, BAR_HI, BAR_HD, BAR_Z0, BAR_Z1, BAR_COUNT };

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
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
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
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long WAIT_LIMIT_NS = 4000000000ull;   // watchdog: 4 s on one wait means a lost signal

// Every thread that consumes a buffer waits on its mbarrier itself (local shared memory, HW-assisted sleep).
// Returns false when the watchdog fired or another CTA aborted the launch.
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity, unsigned* abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    const unsigned long long t0 = globaltimer();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 15u) == 0u) {
            if (ld_acquire_u32(abort_flag) != 0u) return false;
            if (globaltimer() - t0 > WAIT_LIMIT_NS) { atomicExch(abort_flag, 1u); return false; }
        }
    }
    return true;
}
// Thread 0 spins on a global flag, then the CTA syncs; returns true when the launch must be abandoned (watchdog,
// another CTA's abort, or `dead` set by any thread of this CTA after a failed mbarrier wait).
template <typename T, typename LD>
__device__ __forceinline__ bool wait_flag(const T* flag, T target, unsigned* abort_flag, bool dead, LD ld) {
    if (threadIdx.x == 0 && ld(flag) < target) {
        const unsigned long long t0 = globaltimer();
        unsigned spins = 0;
        while (ld(flag) < target) {
            if ((++spins & 31u) == 0u) {
                if (ld_acquire_u32(abort_flag) != 0u) { dead = true; break; }
                if (globaltimer() - t0 > WAIT_LIMIT_NS) { atomicExch(abort_flag, 1u); dead = true; break; }
            }
        }
    }
    return __syncthreads_or(dead ? 1 : 0) != 0;
}
__device__ __forceinline__ bool wait_u32(const unsigned* flag, unsigned target, unsigned* abort_flag, bool dead) {
    return wait_flag(flag, target, abort_flag, dead, [](const unsigned* p) { return ld_acquire_u32(p); });
}
__device__ __forceinline__ bool wait_s64(const long long* flag, long long target, unsigned* abort_flag, bool dead) {
    return wait_flag(flag, target, abort_flag, dead, [](const long long* p) { return ld_acquire_s64(p); });
}

// ------------------------------------------------------------------------------------------------------------
// Warp contraction: 4 columns x 8 prompts, K split over the lanes; returns the (c = lane / 8, p = lane % 8) output.
//   Wq : [KJ][32] float4 — lane's 4 column weights at k = lane + 32 j
//   x  : activation block in the [half][K][4] layout (half = prompt / 4)
// ------------------------------------------------------------------------------------------------------------
template <int OFF, int N>
__device__ __forceinline__ void tr_step(const float (&a)[N], float (&out)[N / 2], bool up) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = up ? a[i] : a[i + N / 2];
        const float keep = up ? a[i + N / 2] : a[i];
        out[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
}

__device__ __forceinline__ float quad_dot(const float4* __restrict__ Wq, const float* __restrict__ x, int K, int KJ) {
    const int lane = threadIdx.x & 31;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
    const float4* xlo = reinterpret_cast<const float4*>(x);
    const float4* xhi = xlo + K;
#pragma unroll 2
    for (int j = 0; j < KJ; ++j) {
        const int k = lane + 32 * j;
        const float4 w = Wq[j * 32 + lane];
        const float4 a = xlo[k], b = xhi[k];
        const float wv[4] = {w.x, w.y, w.z, w.w};
        const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[c * 8 + p] = fmaf(wv[c], xv[p], acc[c * 8 + p]);
    }
    float a16[16], a8[8], a4[4], a2[2], a1[1];
    tr_step<16>(acc, a16, (lane & 16) != 0);
    tr_step<8>(a16, a8, (lane & 8) != 0);
    tr_step<4>(a8, a4, (lane & 4) != 0);
    tr_step<2>(a4, a2, (lane & 2) != 0);
    tr_step<1>(a2, a1, (lane & 1) != 0);
    return a1[0];
}

// offset (floats) of element (channel k, prompt p) in a [half][K][4] block
__device__ __forceinline__ int blk_off(int K, int k, int p) { return ((p >> 2) * K + k) * 4 + (p & 3); }

// ------------------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) wavenet_chain_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: keeps the shuffles convergent
    const int CS = P.CS;
    const int rank = (int)cluster_ctarank();
    const int stage = blockIdx.x / CS;
    const int l_lo = P.stage_lo[stage], l_hi = P.stage_lo[stage + 1];
    const bool first_stage = stage == 0, last_stage = stage == P.NST - 1;
    const int C = P.C, S = P.S, nf = P.nf, ns = P.ns, blk = P.blk;
    const unsigned sbase = smem_u32(smem);
    unsigned* abort_flag = P.abort_flag;

    float* w_s = smem + P.s_w;
    float* head_s = smem + P.s_head;
    float* x1 = smem + P.s_x1;        // [2][blk]  layer input h_l(t), double buffered
    float* x0 = smem + P.s_x0;        // [2][blk]  ring reads h_l(t-d), double buffered
    float* ybuf = smem + P.s_y;       // [2][blk]  gated outputs, double buffered
    float* hin = smem + P.s_hin;      // [Kh*GB]   head input
    float* hid = smem + P.s_hid;      // [Hh*GB]   head hidden
    float* zbuf = smem + P.s_z;       // [2][(Q+1)][GB] raw logits, logit-major (rank 0 receives)
    float* zrows = smem + P.s_zrows;  // [GB][zrow]   per-prompt rows for the sampler
    float* gatebuf = smem + P.s_gate; // [2][GQ][32]  tap-0 (+bias) and tap-1 partial pre-activations
    float* sacc = smem + P.s_sacc;    // [SQ][32]     this CTA's slice of the running skip sum, task-lane layout
    float* stg = smem + P.s_stage + warp * 32;   // warp-private staging (8 chunks of 16 bytes)
    const unsigned bar0 = sbase + P.s_bar * 4;
    auto bar = [&](int i) { return bar0 + 8u * (unsigned)i; };

    // ---- resident weights
    {
        const int n_own = l_hi - l_lo;
        for (int l = 0; l < n_own; ++l) {
            const float4* s4 = reinterpret_cast<const float4*>(P.wpack + ((size_t)(l_lo + l) * CS + rank) * P.layer_block);
            float4* d4 = reinterpret_cast<float4*>(w_s + (size_t)l * P.layer_block);
            for (int i = tid; i < P.layer_block / 4; i += NT) d4[i] = __ldg(s4 + i);
        }
        if (last_stage) {
            const float4* s4 = reinterpret_cast<const float4*>(P.hpack + (size_t)rank * P.head_block);
            float4* d4 = reinterpret_cast<float4*>(head_s);
            for (int i = tid; i < P.head_block / 4; i += NT) d4[i] = __ldg(s4 + i);
        }
    }
    const unsigned xbytes = (unsigned)blk * 4u;
    if (tid == 0) {
        for (int i = 0; i < BAR_COUNT; ++i) mbar_init(bar(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // arm the first phase of every exchange barrier (tx bytes may land before or after the arm)
        mbar_expect_tx(bar(BAR_H0), xbytes); mbar_expect_tx(bar(BAR_H1), xbytes);
        mbar_expect_tx(bar(BAR_Y0), xbytes); mbar_expect_tx(bar(BAR_Y1), xbytes);
        mbar_expect_tx(bar(BAR_HI), (unsigned)P.Kh * GB * 4u);
        mbar_expect_tx(bar(BAR_HD), (unsigned)P.Hh * GB * 4u);
        mbar_expect_tx(bar(BAR_Z0), (unsigned)(P.Q + 1) * GB * 4u);
        mbar_expect_tx(bar(BAR_Z1), (unsigned)(P.Q + 1) * GB * 4u);
    }
    __syncthreads();
    cluster_sync_all();

    // phase bits of the barriers (every thread tracks them; they advance identically in all CTAs of a cluster)
    unsigned ph_h = 0, ph_y = 0, ph_r = 0, ph_hi = 0, ph_hd = 0, ph_z = 0;   // bit b = parity of buffer b
    int xb = 0, yb = 0, rb = 0, zb = 0;
    const bool has_skip = S > 0;
    bool dead = false;   // a wait of this thread gave up: leave at the next CTA-wide check
    long long* trace_row = nullptr;
    int trace_n = 0;
    auto stamp = [&]() {
        if (trace_row && trace_n < TRACE_EV) trace_row[trace_n++] = clock64();
    };

    // all-gather of a staged quad: chunk i (16 B = 4 prompts of one channel) -> byte offset dst_off[i] in every peer
    auto send_quad = [&](int nchunk, int lanes_per_peer_shift, unsigned dst_off, unsigned bar_id) {
        // lane -> chunk = lane % nchunk, first peer = lane / nchunk, peer stride = 32 / nchunk
        const int chunk = lane & (nchunk - 1);
        const float4 v = reinterpret_cast<const float4*>(stg)[chunk];
        for (int peer = lane >> lanes_per_peer_shift; peer < CS; peer += (32 >> lanes_per_peer_shift))
            st_async_v4(mapa(dst_off, (unsigned)peer), v, mapa(bar(bar_id), (unsigned)peer));
    };

    // one TMA bulk copy brings the ring slot h_l(t - d) of this group into x0[buf]
    auto issue_ring_load = [&](int l, long long t, int g, int buf) {
        const Layer& ly = P.layers[l];
        const float* src = P.rings + ly.ring_off + ((size_t)(t % ly.dilation) * P.G + g) * blk;
        mbar_expect_tx(bar(BAR_R0 + buf), xbytes);
        bulk_g2s(sbase + (unsigned)(P.s_x0 + buf * blk) * 4u, src, xbytes, bar(BAR_R0 + buf));
    };

    for (long long t = P.t_begin; t < P.t_end; ++t) {
        const unsigned delivery = (unsigned)(t - P.t_begin);
        const bool head_on = t >= P.t_head;
        for (int g = 0; g < P.n_groups; ++g) {
            trace_row = (P.trace && t == P.trace_t && rank == 0 && tid == 0)
                            ? P.trace + ((size_t)stage * P.G + g) * TRACE_EV : nullptr;
            trace_n = 0;
            if (trace_row) trace_row[trace_n++] = (long long)globaltimer();
            stamp();
            // ---------------- stage input -> x1[xb] ----------------
            if (tid == 0) issue_ring_load(l_lo, t, g, rb);
            if (first_stage) {
                if (wait_s64(P.avail + g, t + 1, abort_flag, dead)) goto done;
                // embedding gather: x1[k][p] = E[q_{b,t}][k]   (EmbeddingIO, modules/io.py:148-154)
                {
                    const int p = warp, b = g * GB + p;
                    long long q = 0;
                    if (b < P.B) q = __ldcg(P.seq + (size_t)b * P.seq_stride + t);
                    q = q < 0 ? 0 : (q >= P.Q ? P.Q - 1 : q);
                    const float* row = P.E + (size_t)q * C;
                    float* dst = x1 + xb * blk;
                    for (int k = lane; k < C; k += 32) dst[blk_off(C, k, p)] = (b < P.B) ? __ldg(row + k) : 0.0f;
                }
            } else {
                if (wait_u32(P.ready + stage * P.G + g, (delivery + 1) * CS, abort_flag, dead)) goto done;
                const float4* mh = reinterpret_cast<const float4*>(P.mail_h + (((size_t)stage * P.G + g) * 2 + (delivery & 1)) * blk);
                float4* dst = reinterpret_cast<float4*>(x1 + xb * blk);
                for (int i = tid; i < blk / 4; i += NT) dst[i] = __ldcg(mh + i);
                if (has_skip) {
                    const float* ms = P.mail_s + ((((size_t)stage * P.G + g) * 2 + (delivery & 1)) * CS + rank) * (size_t)(P.SQ * 32);
                    for (int i = tid; i < P.SQ * 32; i += NT) sacc[i] = __ldcg(ms + i);
                }
                __syncthreads();
                if (tid == 0) red_release_add(P.ack + stage * P.G + g, 1u);
            }
            __syncthreads();
            stamp();
            bool h_local = true;   // x1[xb] was filled by this CTA itself: no mbarrier to wait on

            // ---------------- owned layers ----------------
            for (int l = l_lo; l < l_hi; ++l) {
                const Layer& ly = P.layers[l];
                const float* W = w_s + (size_t)(l - l_lo) * P.layer_block;
                const bool last_owned = (l == l_hi - 1), last_layer = (l == P.L - 1);
                const bool has_res = ly.has_res != 0;
                const bool use_y = has_res || has_skip;
                float* x1c = x1 + xb * blk;
                const float* x0c = x0 + rb * blk;
                const float* yc = ybuf + yb * blk;

                // ---- phase A: the two taps of the dilated conv, 2*GQ warp tasks
                for (int task = warp; task < 2 * P.GQ; task += NW) {
                    const bool tap1 = task < P.GQ;
                    const int q = tap1 ? task : task - P.GQ;
                    if (tap1) {
                        if (!h_local) dead |= !mbar_wait(bar(BAR_H0 + xb), (ph_h >> xb) & 1u, abort_flag);
                        const float v = quad_dot(reinterpret_cast<const float4*>(W + P.o_g1) + (size_t)q * P.KJ * 32, x1c, C, P.KJ);
                        gatebuf[(P.GQ + q) * 32 + lane] = v;
                    } else {
                        dead |= !mbar_wait(bar(BAR_R0 + rb), (ph_r >> rb) & 1u, abort_flag);
                        const float v = quad_dot(reinterpret_cast<const float4*>(W + P.o_g0) + (size_t)q * P.KJ * 32, x0c, C, P.KJ);
                        gatebuf[q * 32 + lane] = v + W[P.o_gb + q * 4 + (lane >> 3)];
                    }
                }
                stamp();
                __syncthreads();   // #1: both halves of the pre-activation are in gatebuf; x1c and x0c are fully consumed
                stamp();
                if (!h_local) ph_h ^= 1u << xb;
                ph_r ^= 1u << rb;
                if (!h_local && tid == 0) mbar_expect_tx(bar(BAR_H0 + xb), xbytes);   // re-arm for its next use
                // prefetch the next unit's ring slot (fast mode; hazard mode loads after the cluster barrier below)
                if (tid == 0 && !last_owned && !P.ring_hazard) issue_ring_load(l + 1, t, g, rb ^ 1);

                // ---- gate: y = tanh(a_f) * sigmoid(a_g)   (wavenet_v2.py:151), then all-gather y
                for (int q = warp; q < P.GQ; q += NW) {
                    const float a = gatebuf[q * 32 + lane] + gatebuf[(P.GQ + q) * 32 + lane];
                    const float v = lane < 16 ? tanhf(a) : sigmoid_acc(a);
                    const float y = v * __shfl_xor_sync(0xffffffffu, v, 16);
                    // lanes < 16: channel c = lane / 8 of the quad's two channels, prompt p = lane % 8
                    if (lane < 16) stg[((lane & 7) >> 2) * 8 + (lane >> 3) * 4 + (lane & 3)] = y;   // chunk = half*2 + c
                    __syncwarp();
                    // chunk (half, c) -> block offset ((half * C + rank*nf + 2q + c) * 4) floats
                    const int chunk = lane & 3;
                    const unsigned coff = (unsigned)(((chunk >> 1) * C + rank * nf + 2 * q + (chunk & 1)) * 4) * 4u;
                    if (use_y) send_quad(4, 2, sbase + (unsigned)(P.s_y + yb * blk) * 4u + coff, BAR_Y0 + yb);
                    if (!has_res && !last_layer) {
                        if (!last_owned) send_quad(4, 2, sbase + (unsigned)(P.s_x1 + (xb ^ 1) * blk) * 4u + coff, BAR_H0 + (xb ^ 1));
                    }
                    if (!has_skip && last_layer && head_on)
                        send_quad(4, 2, sbase + (unsigned)P.s_hin * 4u + coff, BAR_HI);
                    __syncwarp();
                }

                // ---- phase B: residual / skip 1x1 convs on the gathered y
                const int nB = (has_res ? P.RQ : 0) + (has_skip ? P.SQ : 0);
                const bool to_mail = last_owned && !last_layer;
                stamp();
                if (to_mail && wait_u32(P.ack + (stage + 1) * P.G + g, delivery >= 1 ? (delivery - 1) * CS : 0u, abort_flag, dead))
                    goto done;
                float* mh_out = nullptr;
                if (to_mail) mh_out = P.mail_h + (((size_t)(stage + 1) * P.G + g) * 2 + (delivery & 1)) * blk;
                if (use_y) {
                    stamp();
                    dead |= !mbar_wait(bar(BAR_Y0 + yb), (ph_y >> yb) & 1u, abort_flag);
                    stamp();
                    // ring write of this layer's input (own channel slice): every peer has finished reading the slot
                    if (warp == NW - 1) {
                        // nf*GB/4 chunks of 16 B: chunk i -> half = i / nf, channel = rank*nf + i % nf
                        for (int i = lane; i < nf * GB / 4; i += 32) {
                            const int hf = i / nf, ch = rank * nf + i % nf;
                            const float4 v = *reinterpret_cast<const float4*>(x1c + (hf * C + ch) * 4);
                            float* dst = P.rings + ly.ring_off + ((size_t)(t % ly.dilation) * P.G + g) * blk + (hf * C + ch) * 4;
                            __stcg(reinterpret_cast<float4*>(dst), v);
                        }
                    }
                    for (int task = warp; task < nB; task += NW) {
                        const bool is_res = has_res && task < P.RQ;
                        const int c = lane >> 3, p = lane & 7;
                        if (is_res) {
                            const int q = task, ch = rank * nf + 4 * q + c;
                            float v = quad_dot(reinterpret_cast<const float4*>(W + P.o_r) + (size_t)q * P.KJ * 32, yc, C, P.KJ);
                            v = x1c[blk_off(C, ch, p)] + (v + W[P.o_rb + q * 4 + c]);   // h_{l+1} = h_l + conv_res(y)
                            stg[(p >> 2) * 16 + c * 4 + (p & 3)] = v;                   // chunk = half*4 + c
                            __syncwarp();
                            const int chunk = lane & 7;
                            const unsigned coff = (unsigned)(((chunk >> 2) * C + rank * nf + 4 * q + (chunk & 3)) * 4) * 4u;
                            if (!last_owned) {
                                send_quad(8, 3, sbase + (unsigned)(P.s_x1 + (xb ^ 1) * blk) * 4u + coff, BAR_H0 + (xb ^ 1));
                            } else if (to_mail) {
                                if (lane < 8) __stcg(reinterpret_cast<float4*>(mh_out + coff / 4), reinterpret_cast<const float4*>(stg)[lane]);
                            }
                            __syncwarp();
                        } else {
                            const int q = task - (has_res ? P.RQ : 0);
                            float v = quad_dot(reinterpret_cast<const float4*>(W + P.o_s) + (size_t)q * P.KJ * 32, yc, C, P.KJ);
                            v += W[P.o_sb + q * 4 + c];
                            v = (l == 0) ? v : (v + sacc[q * 32 + lane]);               // skips = conv_skip(y) + skips
                            sacc[q * 32 + lane] = v;
                            if (last_layer && head_on) {                                 // all-gather the head input
                                stg[(p >> 2) * 16 + c * 4 + (p & 3)] = v;
                                __syncwarp();
                                const int chunk = lane & 7;
                                const unsigned coff = (unsigned)(((chunk >> 2) * S + rank * ns + 4 * q + (chunk & 3)) * 4) * 4u;
                                send_quad(8, 3, sbase + (unsigned)P.s_hin * 4u + coff, BAR_HI);
                                __syncwarp();
                            }
                        }
                    }
                    if (!has_res && to_mail) {
                        // h_{l+1} = y: own channel slice of the gathered y goes to the mailbox
                        if (warp == NW - 2) {
                            for (int i = lane; i < nf * GB / 4; i += 32) {
                                const int hf = i / nf, ch = rank * nf + i % nf;
                                __stcg(reinterpret_cast<float4*>(mh_out + (hf * C + ch) * 4),
                                       *reinterpret_cast<const float4*>(yc + (hf * C + ch) * 4));
                            }
                        }
                    }
                    ph_y ^= 1u << yb;
                }
                if (!use_y) {
                    // no residual and no skip convs (e.g. the reference's default config): h_{l+1} = y went straight
                    // to the peers' next input buffer.  Nothing here proves that every peer's TMA read of the ring slot
                    // has landed (there is no y gather to wait on), so a cluster barrier separates the reads from the
                    // write; hazard mode (forced for such nets) adds the barrier that orders the write before later reads
                    cluster_sync_all();
                    if (warp == NW - 1) {
                        for (int i = lane; i < nf * GB / 4; i += 32) {
                            const int hf = i / nf, ch = rank * nf + i % nf;
                            const float4 v = *reinterpret_cast<const float4*>(x1c + (hf * C + ch) * 4);
                            float* dst = P.rings + ly.ring_off + ((size_t)(t % ly.dilation) * P.G + g) * blk + (hf * C + ch) * 4;
                            __stcg(reinterpret_cast<float4*>(dst), v);
                        }
                    }
                    if (to_mail) {
                        // own y slice: still in this warp's gate registers only -> re-read from gatebuf is not possible;
                        // the finalize loop staged nothing for the mailbox, so recompute from gatebuf (cheap)
                        for (int q = warp; q < P.GQ; q += NW) {
                            const float a = gatebuf[q * 32 + lane] + gatebuf[(P.GQ + q) * 32 + lane];
                            const float v = lane < 16 ? tanhf(a) : sigmoid_acc(a);
                            const float y = v * __shfl_xor_sync(0xffffffffu, v, 16);
                            if (lane < 16) {
                                const int c = lane >> 3, p = lane & 7;
                                __stcg(mh_out + blk_off(C, rank * nf + 2 * q + c, p), y);
                            }
                        }
                    }
                }
                if (P.ring_hazard) {
                    // small pipelines: order the ring write before any later ring read of the cluster
                    __threadfence();
                    asm volatile("fence.proxy.async.global;" ::: "memory");   // later ring reads are TMA (async proxy)
                    cluster_sync_all();
                    if (tid == 0 && !last_owned) issue_ring_load(l + 1, t, g, rb ^ 1);
                } else if (warp == NW - 1) {
                    __threadfence();   // the ring write is performed before this warp joins the next CTA barrier
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                if (to_mail) {
                    if (has_skip) {
                        float* ms = P.mail_s + ((((size_t)(stage + 1) * P.G + g) * 2 + (delivery & 1)) * CS + rank) * (size_t)(P.SQ * 32);
                        __syncthreads();
                        for (int i = tid; i < P.SQ * 32; i += NT) __stcg(ms + i, sacc[i]);
                    }
                    __syncthreads();
                    if (tid == 0) { __threadfence(); red_release_add(P.ready + (stage + 1) * P.G + g, 1u); }
                }
                if (use_y) {
                    if (tid == 0) mbar_expect_tx(bar(BAR_Y0 + yb), xbytes);   // re-arm
                    yb ^= 1;
                }
                if (!last_owned) { xb ^= 1; h_local = false; }
                rb ^= 1;
                stamp();
            }  // layers
            xb ^= 1;   // the next group's stage input goes to the other buffer

            // ---------------- head + sampler (last stage) ----------------
            if (last_stage && head_on) {
                const float* H1 = head_s + P.o_h1; const float* b1 = head_s + P.o_h1b;
                const float* H2 = head_s + P.o_h2; const float* b2 = head_s + P.o_h2b;
                dead |= !mbar_wait(bar(BAR_HI), ph_hi & 1u, abort_flag);
                stamp();
                for (int q = warp; q < P.HQ; q += NW) {        // hidden = mish(W1 x + b1), all-gathered
                    const int c = lane >> 3, p = lane & 7;
                    float v = quad_dot(reinterpret_cast<const float4*>(H1) + (size_t)q * P.KJh * 32, hin, P.Kh, P.KJh);
                    v = mish_acc(v + b1[q * 4 + c]);
                    stg[(p >> 2) * 16 + c * 4 + (p & 3)] = v;
                    __syncwarp();
                    const int chunk = lane & 7;
                    const unsigned coff = (unsigned)(((chunk >> 2) * P.Hh + rank * P.nh + 4 * q + (chunk & 3)) * 4) * 4u;
                    send_quad(8, 3, sbase + (unsigned)P.s_hid * 4u + coff, BAR_HD);
                    __syncwarp();
                }
                stamp();
                dead |= !mbar_wait(bar(BAR_HD), ph_hd & 1u, abort_flag);
                stamp();
                for (int q = warp; q < P.ZQ; q += NW) {        // z = W2 hidden + b2 -> rank 0, logit-major [o][8]
                    const int c = lane >> 3, p = lane & 7;
                    float v = quad_dot(reinterpret_cast<const float4*>(H2) + (size_t)q * P.KJ2 * 32, hid, P.Hh, P.KJ2);
                    v += b2[q * 4 + c];
                    stg[c * 8 + p] = v;                          // chunk = c*2 + half
                    __syncwarp();
                    if (lane < 8) {
                        const int o = rank * P.nz + 4 * q + (lane >> 1);
                        if (o <= P.Q) {
                            const unsigned doff = sbase + (unsigned)(P.s_z + zb * (P.Q + 1) * GB + o * GB + (lane & 1) * 4) * 4u;
                            st_async_v4(mapa(doff, 0u), reinterpret_cast<const float4*>(stg)[lane], mapa(bar(BAR_Z0 + zb), 0u));
                        }
                    }
                    __syncwarp();
                }
                __syncthreads();   // hin / hid fully consumed by this CTA before they are re-armed
                ph_hi ^= 1u; ph_hd ^= 1u;
                if (tid == 0) {
                    mbar_expect_tx(bar(BAR_HI), (unsigned)P.Kh * GB * 4u);
                    mbar_expect_tx(bar(BAR_HD), (unsigned)P.Hh * GB * 4u);
                }
                if (rank == 0) {
                    stamp();
                    dead |= !mbar_wait(bar(BAR_Z0 + zb), (ph_z >> zb) & 1u, abort_flag);
                    stamp();
                    const float* zsrc = zbuf + zb * (P.Q + 1) * GB;
                    for (int i = tid; i < (P.Q + 1) * GB; i += NT) zrows[(i & 7) * P.zrow + (i >> 3)] = zsrc[i];
                    __syncthreads();
                    if (tid == 0) mbar_expect_tx(bar(BAR_Z0 + zb), (unsigned)(P.Q + 1) * GB * 4u);
                    {
                        const int p = warp, b = g * GB + p;
                        if (b < P.B) {
                            const long long hstep = t - P.t_head, n_head = P.t_end - P.t_head;
                            float* lout = P.logits_out ? P.logits_out + ((size_t)b * n_head + hstep) * P.Q : nullptr;
                            const bool sample = P.temperature != nullptr;
                            float Tt = 1.0f, u = 0.0f;
                            if (sample) {
                                Tt = P.temperature[P.n_temperature == 1 ? 0 : b];
                                u = P.noise[(size_t)b * P.noise_stride + (t + 1 - P.noise_t0)];
                            }
                            const int choice = mmk::decide_warp(zrows + p * P.zrow, P.Q, P.min_temp, lout, sample, Tt, u);
                            if (lane == 0) {
                                if (P.decisions) P.decisions[(size_t)b * n_head + hstep] = choice;
                                if (!P.teacher_forced) __stcg(P.seq + (size_t)b * P.seq_stride + t + 1, (long long)choice);
                            }
                        }
                    }
                    __syncthreads();
                    stamp();
                    if (tid == 0 && !P.teacher_forced) { __threadfence(); st_release_s64(P.avail + g, t + 2); }
                }
                ph_z ^= 1u << zb;
                zb ^= 1;
            }
            if (last_stage && rank == 0 && tid == 0 && g == P.n_groups - 1 && P.step_ts)
                P.step_ts[t - P.t_begin] = globaltimer();
        }  // groups
    }      // time
done:
    // no CTA may exit while peers can still store into its shared memory
    __syncthreads();
    cluster_sync_all();
}

static int pad4(int v) { return (v + 3) / 4 * 4; }

}  // namespace mmk2

using namespace mmk2;

struct wn2_handle {
    Params p{};
    int device = 0, max_batch = 0, rf = 0;
    size_t smem_bytes = 0;
    std::vector<void*> allocs;
    size_t flags_bytes = 0;
    long long* d_trace = nullptr;   // debug timeline, only with MMK_WN_TRACE_T set
    long long trace_t = -1;
};

// fills geometry + weight block offsets + smem carve-up; returns dynamic smem bytes
static size_t wn2_plan(Params& p, int CS, int layers_per_stage) {
    p.CS = CS;
    p.nf = p.C / CS; p.ns = p.S / CS; p.nh = p.Hh / CS;
    p.GQ = p.nf / 2; p.RQ = p.nf / 4; p.SQ = p.ns / 4; p.HQ = p.nh / 4;
    p.ZQ = ((p.Q + 1 + CS - 1) / CS + 3) / 4; p.nz = 4 * p.ZQ;
    p.KJ = p.C / 32; p.KJh = p.Kh / 32; p.KJ2 = p.Hh / 32;
    p.blk = p.C * GB;
    int o = 0;
    auto take = [&](int floats) { int r = o; o += pad4(floats); return r; };
    p.o_g1 = take(p.GQ * p.KJ * 32 * 4); p.o_g0 = take(p.GQ * p.KJ * 32 * 4); p.o_gb = take(p.GQ * 4);
    p.o_r = take(p.RQ * p.KJ * 32 * 4); p.o_rb = take(std::max(4, p.RQ * 4));
    p.o_s = take(p.SQ * p.KJ * 32 * 4); p.o_sb = take(std::max(4, p.SQ * 4));
    p.layer_block = o;
    o = 0;
    p.o_h1 = take(p.HQ * p.KJh * 32 * 4); p.o_h1b = take(p.HQ * 4);
    p.o_h2 = take(p.ZQ * p.KJ2 * 32 * 4); p.o_h2b = take(p.ZQ * 4);
    p.head_block = o;
    o = 0;
    p.zrow = pad4(p.Q + 1) + 4;
    if ((p.zrow % 32) == 0) p.zrow += 4;
    p.s_w = take(layers_per_stage * p.layer_block);
    p.s_head = take(p.head_block);
    p.s_x1 = take(2 * p.blk); p.s_x0 = take(2 * p.blk); p.s_y = take(2 * p.blk);
    p.s_hin = take(p.Kh * GB); p.s_hid = take(p.Hh * GB);
    p.s_z = take(2 * (p.Q + 1) * GB); p.s_zrows = take(GB * p.zrow);
    p.s_gate = take(2 * p.GQ * 32); p.s_sacc = take(std::max(4, p.SQ * 32));
    p.s_stage = take(NW * 32);
    p.s_bar = take(BAR_COUNT * 2);
    p.smem_floats = o;
    return (size_t)o * sizeof(float);
}

static int wn2_query_clusters(int CS, size_t smem, int* out) {
    const void* k = (const void*)wavenet_chain_kernel;
    MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS * 8);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
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

int wn2_destroy(wn2_handle* h) {
    if (!h) return 0;
    for (void* a : h->allocs) cudaFree(a);
    delete h;
    return 0;
}

int wn2_create(const mmk_wavenet_desc* d, int max_batch, wn2_handle** out, int* unsupported) {
    *unsupported = 1;
    if (d->dilated_dim % 32 || d->head_hidden % 32 || (d->skips_dim % 32) || d->n_layers > MAX_LAYERS) return 1;
    auto* h = new wn2_handle();
    Params& p = h->p;
    cudaGetDevice(&h->device);
    p.L = d->n_layers; p.C = d->dilated_dim; p.S = d->skips_dim; p.Hh = d->head_hidden; p.Q = d->q_levels;
    p.Kh = p.S > 0 ? p.S : p.C;
    p.min_temp = d->min_temperature;
    h->max_batch = max_batch;
    p.G = (max_batch + GB - 1) / GB;
    int rf = 1;
    bool any_no_y = false;
    for (int l = 0; l < p.L; ++l) {
        rf += d->dilations[l];
        if (!d->conv_res_w[l] && p.S == 0) any_no_y = true;
    }
    h->rf = rf;
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);

    int best_cs = 0, best_nst = 0;
    const char* force_cs = getenv("MMK_WN_CLUSTER");
    const char* force_nst = getenv("MMK_WN_STAGES");
    for (int CS : {16, 8, 4, 2, 1}) {
        if (force_cs && atoi(force_cs) != CS) continue;
        if (p.C % CS || p.S % CS || p.Hh % CS) continue;
        const int nf = p.C / CS, ns = p.S / CS, nh = p.Hh / CS;
        if (nf % 4 || ns % 4 || nh % 4) continue;            // whole quads per CTA
        int nst_min = 0;
        for (int nst = 1; nst <= std::min(p.L, MAX_STAGES); ++nst) {
            Params q = p;
            if (wn2_plan(q, CS, (p.L + nst - 1) / nst) <= (size_t)max_optin) { nst_min = nst; break; }
        }
        if (!nst_min) continue;
        Params q = p;
        const size_t smem_min = wn2_plan(q, CS, (p.L + nst_min - 1) / nst_min);
        int max_clusters = 0;
        if (CS == 1) {
            int per_sm = 0, sms = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
            cudaFuncSetAttribute((const void*)wavenet_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_min);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)wavenet_chain_kernel, NT, smem_min);
            max_clusters = per_sm * sms;
        } else if (wn2_query_clusters(CS, smem_min, &max_clusters)) { wn2_destroy(h); *unsupported = 0; return 1; }
        if (max_clusters < nst_min) continue;
        int nst = std::min(std::min(max_clusters, p.L), MAX_STAGES);
        if (force_nst) nst = std::max(nst_min, std::min(nst, atoi(force_nst)));
        else nst = std::min(nst, std::max(nst_min, std::max(1, p.G)));
        best_cs = CS; best_nst = nst;
        break;
    }
    if (!best_cs) { wn2_destroy(h); return 1; }
    const int CS = best_cs, NST = best_nst;
    p.NST = NST;
    const int per = (p.L + NST - 1) / NST;
    h->smem_bytes = wn2_plan(p, CS, per);
    {
        int base = p.L / NST, extra = p.L % NST, lo = 0;
        for (int s = 0; s < NST; ++s) { p.stage_lo[s] = lo; lo += base + (s < extra ? 1 : 0); }
        p.stage_lo[NST] = p.L;
    }
    // a ring slot written in one unit is read again G * (layers of the stage) * dilation units later; tiny
    // pipelines (and nets without any cluster-wide exchange between a layer's read and write) take the barrier path
    const int min_layers = p.L / NST;
    p.ring_hazard = (any_no_y || p.G * min_layers < 4) ? 1 : 0;
    if (const char* e = getenv("MMK_WN_RING_HAZARD")) p.ring_hazard = ((atoi(e) != 0) || any_no_y) ? 1 : 0;
    *unsupported = 0;
    const void* kern = (const void*)wavenet_chain_kernel;
    MMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    if (CS > 8) MMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    // ---- pack weights: lane-major quads, Wq[j][lane] = 4 column weights at k = lane + 32 j
    const int C = p.C, S = p.S, nf = p.nf, ns = p.ns, KJ = p.KJ;
    std::vector<float> wpack((size_t)p.L * CS * p.layer_block, 0.0f);
    long long ring_off = 0;
    auto put_quad = [&](float* dst, int q, int KJq, auto weight_of /* (col c, k) -> float */) {
        for (int j = 0; j < KJq; ++j)
            for (int ln = 0; ln < 32; ++ln)
                for (int c = 0; c < 4; ++c) dst[(((size_t)q * KJq + j) * 32 + ln) * 4 + c] = weight_of(c, ln + 32 * j);
    };
    for (int l = 0; l < p.L; ++l) {
        const bool has_res = d->conv_res_w[l] != nullptr;
        p.layers[l].dilation = d->dilations[l];
        p.layers[l].has_res = has_res ? 1 : 0;
        p.layers[l].ring_off = ring_off;
        ring_off += (long long)d->dilations[l] * p.G * p.blk;
        const float* wd = d->conv_dil_w[l];   // (2C, C, 2): [o][c][tap], tap 0 = older sample
        const float* bd = d->conv_dil_b[l];
        for (int r = 0; r < CS; ++r) {
            float* blkp = wpack.data() + ((size_t)l * CS + r) * p.layer_block;
            for (int q = 0; q < p.GQ; ++q) {
                // quad columns: f(2q), f(2q+1), g(2q), g(2q+1) of this rank's channels
                auto orow = [&](int c) { return (c < 2 ? 0 : C) + r * nf + 2 * q + (c & 1); };
                put_quad(blkp + p.o_g1, q, KJ, [&](int c, int k) { return wd[((size_t)orow(c) * C + k) * 2 + 1]; });
                put_quad(blkp + p.o_g0, q, KJ, [&](int c, int k) { return wd[((size_t)orow(c) * C + k) * 2 + 0]; });
                for (int c = 0; c < 4; ++c) blkp[p.o_gb + q * 4 + c] = bd[orow(c)];
            }
            if (has_res)
                for (int q = 0; q < p.RQ; ++q) {
                    put_quad(blkp + p.o_r, q, KJ, [&](int c, int k) { return d->conv_res_w[l][(size_t)(r * nf + 4 * q + c) * C + k]; });
                    for (int c = 0; c < 4; ++c) blkp[p.o_rb + q * 4 + c] = d->conv_res_b[l][r * nf + 4 * q + c];
                }
            if (S > 0)
                for (int q = 0; q < p.SQ; ++q) {
                    put_quad(blkp + p.o_s, q, KJ, [&](int c, int k) { return d->conv_skip_w[l][(size_t)(r * ns + 4 * q + c) * C + k]; });
                    for (int c = 0; c < 4; ++c) blkp[p.o_sb + q * 4 + c] = d->conv_skip_b[l][r * ns + 4 * q + c];
                }
        }
    }
    std::vector<float> hpack((size_t)CS * p.head_block, 0.0f);
    for (int r = 0; r < CS; ++r) {
        float* blkp = hpack.data() + (size_t)r * p.head_block;
        for (int q = 0; q < p.HQ; ++q) {
            put_quad(blkp + p.o_h1, q, p.KJh, [&](int c, int k) { return d->head_w1[(size_t)(r * p.nh + 4 * q + c) * p.Kh + k]; });
            for (int c = 0; c < 4; ++c) blkp[p.o_h1b + q * 4 + c] = d->head_b1[r * p.nh + 4 * q + c];
        }
        for (int q = 0; q < p.ZQ; ++q) {
            put_quad(blkp + p.o_h2, q, p.KJ2, [&](int c, int k) {
                const int o = r * p.nz + 4 * q + c;
                return o <= p.Q ? d->head_w2[(size_t)o * p.Hh + k] : 0.0f;
            });
            for (int c = 0; c < 4; ++c) {
                const int o = r * p.nz + 4 * q + c;
                blkp[p.o_h2b + q * 4 + c] = o <= p.Q ? d->head_b2[o] : 0.0f;
            }
        }
    }
    bool ok = true;
    auto dev_alloc = [&](size_t bytes, const void* src) -> void* {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, bytes) != cudaSuccess) { ok = false; return nullptr; }
        h->allocs.push_back(ptr);
        if (src) cudaMemcpy(ptr, src, bytes, cudaMemcpyHostToDevice); else cudaMemset(ptr, 0, bytes);
        return ptr;
    };
    p.wpack = (const float*)dev_alloc(wpack.size() * sizeof(float), wpack.data());
    p.hpack = (const float*)dev_alloc(hpack.size() * sizeof(float), hpack.data());
    p.E = (const float*)dev_alloc((size_t)p.Q * C * sizeof(float), d->embedding);
    p.rings = (float*)dev_alloc((size_t)ring_off * sizeof(float), nullptr);
    p.mail_h = (float*)dev_alloc((size_t)(NST + 1) * p.G * 2 * p.blk * sizeof(float), nullptr);
    p.mail_s = (float*)dev_alloc((size_t)(NST + 1) * p.G * 2 * CS * std::max(1, p.SQ * 32) * sizeof(float), nullptr);
    h->flags_bytes = sizeof(long long) * (size_t)p.G + sizeof(unsigned) * (2 * (size_t)(NST + 1) * p.G + 16);
    void* flags = dev_alloc(h->flags_bytes, nullptr);
    if (!ok) { wn2_destroy(h); MMK_FAIL("cudaMalloc failed while creating the WaveNet handle"); }
    p.avail = (long long*)flags;
    p.ready = (unsigned*)(p.avail + p.G);
    p.ack = p.ready + (size_t)(NST + 1) * p.G;
    p.abort_flag = p.ack + (size_t)(NST + 1) * p.G;
    if (const char* e = getenv("MMK_WN_TRACE_T")) {
        h->trace_t = atoll(e);
        h->d_trace = (long long*)dev_alloc((size_t)NST * p.G * TRACE_EV * sizeof(long long), nullptr);
    }
    MMK_CUDA(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

int wn2_launch_info(wn2_handle* h, mmk_launch_info* out) {
    out->cluster_size = h->p.CS; out->n_stages = h->p.NST; out->group_size = GB; out->threads = NT;
    out->smem_bytes = (int)h->smem_bytes; out->sm_used = h->p.CS * h->p.NST;
    return 0;
}

int wn2_sync_check(wn2_handle* h, void* stream) {
    unsigned aborted = 0;
    MMK_CUDA(cudaMemcpyAsync(&aborted, h->p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MMK_CHECK(aborted == 0, "WaveNet kernel watchdog fired: an inter-stage wait timed out (results invalid)");
    if (h->d_trace) {   // debug: dump the timeline of step MMK_WN_TRACE_T as text (stage group stamps...)
        const size_t n = (size_t)h->p.NST * h->p.G * TRACE_EV;
        std::vector<long long> tr(n);
        MMK_CUDA(cudaMemcpy(tr.data(), h->d_trace, n * sizeof(long long), cudaMemcpyDeviceToHost));
        const char* path = getenv("MMK_WN_TRACE_FILE");
        if (FILE* f = fopen(path ? path : "wn_trace.txt", "w")) {
            for (int s = 0; s < h->p.NST; ++s)
                for (int g = 0; g < h->p.G; ++g) {
                    fprintf(f, "%d %d", s, g);
                    for (int e = 0; e < TRACE_EV; ++e) fprintf(f, " %lld", tr[((size_t)s * h->p.G + g) * TRACE_EV + e]);
                    fprintf(f, "\n");
                }
            fclose(f);
        }
        MMK_CUDA(cudaMemset(h->d_trace, 0, n * sizeof(long long)));
    }
    return 0;
}

__global__ void wn2_init_flags_kernel(long long* avail, int G, long long avail0, unsigned* u32s, int n_u32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G) avail[i] = avail0;
    if (i < n_u32) u32s[i] = 0u;
}

int wn2_run(wn2_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
            int64_t t_end, int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Params p = h->p;
    p.seq = reinterpret_cast<long long*>(d_seq) - seq_t0;
    p.seq_stride = seq_stride; p.t_begin = t_begin; p.t_head = t_head; p.t_end = t_end;
    p.B = B; p.n_groups = (B + GB - 1) / GB; p.teacher_forced = teacher_forced ? 1 : 0;
    p.temperature = d_temperature; p.n_temperature = n_temperature;
    p.noise = d_noise; p.noise_stride = noise_stride; p.noise_t0 = noise_t0;
    p.logits_out = d_logits_out; p.decisions = reinterpret_cast<long long*>(d_decisions); p.step_ts = d_step_ts;
    p.trace = h->d_trace; p.trace_t = h->d_trace ? t_begin + h->trace_t : -1;
    const int n_u32 = 2 * (p.NST + 1) * p.G + 1;
    const long long avail0 = teacher_forced ? (t_end + 1) : (t_head + 1);
    wn2_init_flags_kernel<<<(std::max(p.G, n_u32) + 255) / 256, 256, 0, st>>>(p.avail, p.G, avail0, p.ready, n_u32);
    MMK_CUDA(cudaGetLastError());
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.CS * p.NST);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[] = {&p};
    MMK_CUDA(cudaLaunchKernelExC(&cfg, (const void*)wavenet_chain_kernel, args));
    return 0;
}
