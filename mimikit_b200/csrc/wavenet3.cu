// Warp-autonomous persistent WaveNet generation kernel (sm_100a) — the fast path.  Same function as wavenet.cu (the
// general kernel; see that file for the reference lines it follows: wavenet_v2.py:131-176, 276-293, 447-452;
// modules/io.py:148-154; networks/mlp.py:44-63; modules/targets.py:40-52).
//
// A generated sample is a dependency chain (L layers x 2 dependent contractions, head, sampler), so a step costs
// L x (latency of one layer for one prompt group) whatever the batch.  Everything here shortens that chain:
//   * stage = cluster of CS CTAs that keeps a contiguous range of layers resident in shared memory (fp32); the
//     output rows of every contraction are split over the CTAs, groups of 8 prompts flow through the stages.
//   * the unit of work is a warp task: 2 (or 3) output rows x 8 prompts, K split over the 32 lanes (16-24 register
//     accumulators per lane), reduced by a 16-shuffle transpose tree that leaves output (row, prompt) in lane
//     16*row + prompt.  A gate channel's filter and gate rows are one task, so tanh * sigmoid needs one more shuffle.
//   * warps never meet at a CTA barrier inside a layer: all-gathers are st.async stores into every peer's shared
//     memory that complete_tx on the peer's mbarrier; each warp waits on its local mbarrier and goes on.
//   * the older tap of the dilated conv is applied when a layer input is PRODUCED (in the bubble while the gated
//     output is in flight) and parked, as a pre-activation, in a lane-private ring in global memory; d steps later the
//     same lane fetches it one layer ahead.  No shared ring, no TMA, no cross-CTA ring hazard.
//   * running skip sums live in registers of the lanes that produce them, for all layers of a stage.
//   * stage-to-stage hand-off and the sampled index travel as flagged 8-byte words (data, tag): the consumer polls
//     the payload itself; no separate flag, no fence on the path.
#include "common.cuh"
#include "sampler.cuh"
#include "wavenet_impl.h"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace mmk3 {

using mmk::mish_acc;

constexpr int NT = 256;   // threads per CTA
constexpr int NW = NT / 32;
constexpr int GB = 8;     // prompts per pipeline group
constexpr int MAX_LAYERS = 96;
constexpr int MAX_STAGES = 32;
constexpr int MAX_TS = 4;      // warp tasks of one kind per warp (template instantiations: 1, 2, 4)
constexpr int TRACE_EV = 48;

struct Layer {
    int dilation, has_res;
    long long ring_off;   // float offset of this layer's pre-activation ring
};
struct LayerS {           // shared-memory copy of the owned layers (dynamic indexing into kernel params is slow)
    unsigned dilation, mask;   // mask = dilation - 1 when the dilation is a power of two, else 0xffffffff
    int has_res, pad;
    long long ring_off;
};
constexpr int MAX_OWN = 16;    // layers per stage held in the shared copy

struct Params {
    int L, C, S, Hh, Q, Kh, CS, NST, G;
    int nf, ns, nr, nh, nz;          // per-CTA: gate channels, skip rows, residual rows, head hidden rows, logit rows
    int KJ, KJh, KJ2;                // contraction depth / 32 of the three kinds (C, Kh, Hh)
    int o_ta, o_b0, o_tb, o_bb, layer_block;      // float offsets inside a (layer, rank) weight block
    int o_h1, o_h1b, o_h2, o_h2b, o_h2x, o_h2xb, head_block;
    int s_w, s_head, s_x1, s_y, s_hin, s_hid, s_z, s_bar, s_ly, smem_floats;
    int zrow, blk;                   // blk = C * GB floats: one activation block in the [half][K][4] layout
    float min_temp;
    Layer layers[MAX_LAYERS];
    int stage_lo[MAX_STAGES + 1];
    const float* wpack; const float* hpack; const float* E;
    float* rings;
    uint4* mail_h;                   // [(stage, g, parity)][blk / 2] lines {f0, tag, f1, tag}
    uint2* mail_s;                   // [(stage, g, parity)][CS][skip tasks][16] words {f, tag}
    unsigned long long* samples;     // [G][GB] words {index, tag}
    unsigned* ack; unsigned* abort_flag;
    // this run
    long long* seq;
    long long seq_stride, t_begin, t_head, t_end;
    int B, n_groups, teacher_forced, n_temperature;
    const float* temperature; const float* noise;
    long long noise_stride, noise_t0;
    float* logits_out; long long* decisions; unsigned long long* step_ts;
    long long* trace; long long trace_t;
};

enum { BAR_X0 = 0, BAR_X1, BAR_Y0, BAR_Y1, BAR_HI, BAR_HD, BAR_Z, BAR_COUNT };

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
__device__ __forceinline__ void st_async_f32(unsigned raddr, float v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
                 ::"r"(raddr), "f"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_poll_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_poll_v2(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flagged_v4(uint4* p, float a, float b, unsigned tag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag),
                 "r"(__float_as_uint(b)), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_flagged_v2(uint2* p, unsigned a, unsigned tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(tag) : "memory");
}
__device__ __forceinline__ void red_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long WAIT_LIMIT_NS = 4000000000ull;   // watchdog: 4 s on one wait means a lost signal

// A spin that gives up when the watchdog fires or another CTA aborted the launch.  `ready()` is re-evaluated.
template <typename F>
__device__ __forceinline__ bool spin_until(F ready, unsigned* abort_flag) {
    if (ready()) return true;
    const unsigned long long t0 = globaltimer();
    unsigned spins = 0;
    while (!ready()) {
        if ((++spins & 63u) == 0u) {
            if (ld_relaxed_u32(abort_flag) != 0u) return false;
            if (globaltimer() - t0 > WAIT_LIMIT_NS) { atomicExch(abort_flag, 1u); return false; }
        }
    }
    return true;
}
// mbarrier wait: the first probe is inline (the hardware suspends the warp inside try_wait), the retry loop with the
// watchdog is out of line so that the hot path stays small in the instruction cache
__device__ __noinline__ bool mbar_wait_slow(unsigned bar, unsigned parity, unsigned* abort_flag) {
    return spin_until([&] { return mbar_try_wait(bar, parity); }, abort_flag);
}
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity, unsigned* abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    return mbar_wait_slow(bar, parity, abort_flag);
}
__device__ __noinline__ bool ack_wait(const unsigned* p, unsigned target, unsigned* abort_flag) {
    return spin_until([&] { return ld_relaxed_u32(p) >= target; }, abort_flag);
}
__device__ __noinline__ bool poll_line(const uint4* p, unsigned tag, uint4& v, unsigned* abort_flag) {
    return spin_until([&] { v = ld_poll_v4(p); return v.y == tag && v.w == tag; }, abort_flag);
}
__device__ __noinline__ bool poll_word(const uint2* p, unsigned tag, uint2& v, unsigned* abort_flag) {
    return spin_until([&] { v = ld_poll_v2(p); return v.y == tag; }, abort_flag);
}

// ------------------------------------------------------------------------------------------------------------
// Warp task: ROWS output rows x 8 prompts, K split over the lanes (k = lane + 32 j).
//   W  : [KJ][32] float2 — (row A, row B) weights of this lane's k;  Wx : [KJ][32] float, the optional third row
//   x  : activation block, [half][K][4] layout (half = prompt / 4)
// ------------------------------------------------------------------------------------------------------------
template <int ROWS>
__device__ __forceinline__ void dot_rows(const float* __restrict__ W, const float* __restrict__ Wx,
                                         const float* __restrict__ x, int K, int KJ, float (&acc)[ROWS][8]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[r][p] = 0.0f;
    const float2* W2 = reinterpret_cast<const float2*>(W);
    const float4* xlo = reinterpret_cast<const float4*>(x);
    const float4* xhi = xlo + K;
#pragma unroll 4
    for (int j = 0; j < KJ; ++j) {
        const int k = lane + 32 * j;
        const float2 w = W2[k];
        const float4 a = xlo[k], b = xhi[k];
        const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            acc[0][p] = fmaf(w.x, xv[p], acc[0][p]);
            acc[1][p] = fmaf(w.y, xv[p], acc[1][p]);
        }
        if (ROWS == 3) {
            const float wx = Wx[k];
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[ROWS - 1][p] = fmaf(wx, xv[p], acc[ROWS - 1][p]);
        }
    }
}

// Transposing reduction over the 32 lanes: lane L ends with output (row = L >> 4, prompt = L & 7), i.e. every
// output is held twice (lanes L and L ^ 8).  16 shuffles, fixed summation order.
__device__ __forceinline__ float tree2(const float (&a0)[8], const float (&a1)[8]) {
    const int lane = threadIdx.x & 31;
    const bool b4 = lane & 16, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    float v8[8], v4[4], v2[2];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const float keep = b4 ? a1[p] : a0[p], send = b4 ? a0[p] : a1[p];
        v8[p] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b2 ? v8[4 + i] : v8[i], send = b2 ? v8[i] : v8[4 + i];
        v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b1 ? v4[2 + i] : v4[i], send = b1 ? v4[i] : v4[2 + i];
        v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float keep = b0 ? v2[1] : v2[0], send = b0 ? v2[0] : v2[1];
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    return v;
}
// One row: every lane ends with output (prompt = L & 7).
__device__ __forceinline__ float tree1(const float (&a)[8]) {
    const int lane = threadIdx.x & 31;
    const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    float v4[4], v2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b2 ? a[4 + i] : a[i], send = b2 ? a[i] : a[4 + i];
        v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b1 ? v4[2 + i] : v4[i], send = b1 ? v4[i] : v4[2 + i];
        v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float keep = b0 ? v2[1] : v2[0], send = b0 ? v2[0] : v2[1];
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// The 4 prompts [4*h4, 4*h4+4) of row `row` after tree2, gathered into one lane.
__device__ __forceinline__ float4 gather4(float v, int row, int h4) {
    const int src = row * 16 + h4 * 4;
    float4 r;
    r.x = __shfl_sync(0xffffffffu, v, src);
    r.y = __shfl_sync(0xffffffffu, v, src + 1);
    r.z = __shfl_sync(0xffffffffu, v, src + 2);
    r.w = __shfl_sync(0xffffffffu, v, src + 3);
    return r;
}

// offset (floats) of element (channel k, prompt p) in a [half][K][4] block
__device__ __forceinline__ int blk_off(int K, int k, int p) { return ((p >> 2) * K + k) * 4 + (p & 3); }

// tanh for the filter half (m = 2), sigmoid for the gate half (m = 1), one instruction stream:
// tanh(a) = 2 / (1 + exp(-2a)) - 1,  sigmoid(a) = 1 / (1 + exp(-a))
__device__ __forceinline__ float gate_act(float a, bool gate_half) {
    const float m = gate_half ? 1.0f : 2.0f;
    const float e = expf(-m * a);
    return __fdiv_rn(m, 1.0f + e) - (m - 1.0f);
}

// ------------------------------------------------------------------------------------------------------------
// The kernel.  TS = max warp tasks of one kind per warp (ceil(rows-per-CTA / 2 / 8)).
// ------------------------------------------------------------------------------------------------------------
// Gate task: one pass over the layer input gives the newer-tap rows (f, g) for NOW and the older-tap rows (f, g) that
// are parked for t + dilation.  W : [K][4] float4 = (f newer, g newer, f older, g older) at contraction index k.
__device__ __forceinline__ void dot_gate(const float* __restrict__ W, const float* __restrict__ x, int K, int KJ,
                                         float (&acc)[4][8]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[r][p] = 0.0f;
    const float4* W4 = reinterpret_cast<const float4*>(W);
    const float4* xlo = reinterpret_cast<const float4*>(x);
    const float4* xhi = xlo + K;
#pragma unroll 4
    for (int j = 0; j < KJ; ++j) {
        const int k = lane + 32 * j;
        const float4 w = W4[k];
        const float4 a = xlo[k], b = xhi[k];
        const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            acc[0][p] = fmaf(w.x, xv[p], acc[0][p]);
            acc[1][p] = fmaf(w.y, xv[p], acc[1][p]);
            acc[2][p] = fmaf(w.z, xv[p], acc[2][p]);
            acc[3][p] = fmaf(w.w, xv[p], acc[3][p]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The kernel.  TS = max warp tasks of one kind per warp (ceil(rows-per-CTA / 2 / 8)).
// ------------------------------------------------------------------------------------------------------------
template <int TS, bool TRACE>
__global__ void __launch_bounds__(NT, 1) wavenet_warp_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int CS = P.CS;
    const int rank = (int)cluster_ctarank();
    const int stage = blockIdx.x / CS;
    const int l_lo = P.stage_lo[stage], l_hi = P.stage_lo[stage + 1];
    const bool first_stage = stage == 0, last_stage = stage == P.NST - 1;
    const int C = P.C, S = P.S, nf = P.nf, ns = P.ns, nr = P.nr, blk = P.blk, KJ = P.KJ;
    const unsigned sbase = smem_u32(smem);
    unsigned* abort_flag = P.abort_flag;
    const bool has_skip = S > 0;
    const int n_skip_tasks = ns / 2;

    const float* w_s = smem + P.s_w;
    const float* head_s = smem + P.s_head;
    float* x1 = smem + P.s_x1;        // [2][blk]  layer input h_l(t)
    float* ybuf = smem + P.s_y;       // [2][blk]  gated outputs
    float* hin = smem + P.s_hin;      // [Kh*GB]   head input (skip sums)
    float* hid = smem + P.s_hid;      // [Hh*GB]   head hidden
    float* zbuf = smem + P.s_z;       // [GB][zrow] raw logits per prompt (rank 0 receives)
    const unsigned off_bar0 = (unsigned)P.s_bar * 4u;
    auto bar = [&](int i) { return sbase + off_bar0 + 8u * (unsigned)i; };
    const unsigned xbytes = (unsigned)blk * 4u;

    // lane roles in an all-gather: one row (2 chunks) -> peer = lane >> 1; two rows (4 chunks) -> peers lane >> 2 (+8)
    const int h4 = lane & 1;            // which 4 prompts, one-row gathers
    const int ch4 = lane & 3;           // chunk of a two-row gather: row = ch4 >> 1, prompts half = ch4 & 1
    const int lane16 = (lane & 7) | ((lane >> 4) << 3);   // index of this lane's output among the 16 of a task
    const int my_row = lane >> 4, my_p = lane & 7;
    // remote shared-memory windows of this lane's peers (a CTA's window is linear: remote = window + local offset)
    const int peer_y = lane >> 1, peer_b0 = lane >> 2, peer_b1 = (lane >> 2) + 8;
    const unsigned win_y = mapa(sbase, (unsigned)min(peer_y, CS - 1)) - sbase;
    const unsigned win_b0 = mapa(sbase, (unsigned)min(peer_b0, CS - 1)) - sbase;
    const unsigned win_b1 = mapa(sbase, (unsigned)min(peer_b1, CS - 1)) - sbase;
    const unsigned win_0 = mapa(sbase, 0u) - sbase;
    // two-row all-gather of (c4 = this lane's chunk) into byte offset `dst` of every peer, signalling barrier `b`
    auto send2 = [&](unsigned dst, const float4& c4, int b) {
        if (peer_b0 < CS) st_async_v4(win_b0 + dst, c4, win_b0 + bar(b));
        if (peer_b1 < CS) st_async_v4(win_b1 + dst, c4, win_b1 + bar(b));
    };

    // ---- resident weights
    {
        const int n_own = l_hi - l_lo;
        for (int l = 0; l < n_own; ++l) {
            const float4* s4 = reinterpret_cast<const float4*>(P.wpack + ((size_t)(l_lo + l) * CS + rank) * P.layer_block);
            float4* d4 = reinterpret_cast<float4*>(smem + P.s_w + (size_t)l * P.layer_block);
            for (int i = tid; i < P.layer_block / 4; i += NT) d4[i] = __ldg(s4 + i);
        }
        if (last_stage) {
            const float4* s4 = reinterpret_cast<const float4*>(P.hpack + (size_t)rank * P.head_block);
            float4* d4 = reinterpret_cast<float4*>(smem + P.s_head);
            for (int i = tid; i < P.head_block / 4; i += NT) d4[i] = __ldg(s4 + i);
        }
    }
    LayerS* ly_s = reinterpret_cast<LayerS*>(smem + P.s_ly);
    if (tid < l_hi - l_lo) {
        const Layer& ly = P.layers[l_lo + tid];
        const unsigned d = (unsigned)ly.dilation;
        ly_s[tid] = LayerS{d, (d & (d - 1u)) == 0u ? d - 1u : 0xffffffffu, ly.has_res, 0, ly.ring_off};
    }
    if (tid == 0) {
        for (int i = 0; i < BAR_COUNT; ++i) mbar_init(bar(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // arm the first phase of every exchange barrier (tx bytes may land before or after the arm)
        mbar_expect_tx(bar(BAR_X0), xbytes); mbar_expect_tx(bar(BAR_X1), xbytes);
        mbar_expect_tx(bar(BAR_Y0), xbytes); mbar_expect_tx(bar(BAR_Y1), xbytes);
        mbar_expect_tx(bar(BAR_HI), (unsigned)P.Kh * GB * 4u);
        mbar_expect_tx(bar(BAR_HD), (unsigned)P.Hh * GB * 4u);
        mbar_expect_tx(bar(BAR_Z), (unsigned)(P.Q + 1) * GB * 4u);
    }
    __syncthreads();
    cluster_sync_all();

    // phase parities of the barriers (warp-uniform, identical in every CTA of a cluster)
    unsigned par_x = 0u, par_y = 0u, cnt_hi = 0u, cnt_hd = 0u, cnt_z = 0u;   // bit b of par_* = parity of buffer b
    unsigned n = 0;                 // running layer-unit counter: buffers alternate with it
    bool dead = false;
    float sacc[TS];                 // running skip sums of this lane's outputs (skip tasks come first in the task list)
    float a0[TS];                   // prefetched older-tap pre-activations (bias included) for the next layer-unit
#pragma unroll
    for (int s = 0; s < TS; ++s) { sacc[s] = 0.0f; a0[s] = 0.0f; }

    // what the next unit's fill needs, fetched while the current unit's last layer is in flight
    uint4 pf_line0 = make_uint4(0, 0, 0, 0), pf_line1 = make_uint4(0, 0, 0, 0);
    uint2 pf_skip[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) pf_skip[s] = make_uint2(0, 0);
    float pf_e[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    bool pf_ok = false;             // pf_line* / pf_skip were fetched for the unit about to start (tags still to be checked);
                                    // first stage: pf_e holds the embedding row of this warp's prompt

    long long* trace_row = nullptr;
    int trace_n = 0;
    auto stamp = [&]() { if (TRACE && trace_row && trace_n < TRACE_EV) trace_row[trace_n++] = clock64(); };

    // lane-private ring slot of (layer l, time t, group g, gate task)
    auto ring_ptr = [&](int l, long long t, int g, int task) -> float* {
        const LayerS ly = ly_s[l - l_lo];
        const unsigned slot = ly.mask != 0xffffffffu ? ((unsigned)t & ly.mask) : ((unsigned)t % ly.dilation);   // t < 2^32
        return P.rings + ly.ring_off + ((((size_t)slot * P.G + g) * CS + rank) * nf + task) * 16 + lane16;
    };
    __syncthreads();
#pragma unroll
    for (int s = 0; s < TS; ++s) {
        const int task = warp + NW * s;
        if (task < nf) a0[s] = __ldcg(ring_ptr(l_lo, P.t_begin, 0, task));
    }

    for (long long t = P.t_begin; t < P.t_end; ++t) {
        const unsigned delivery = (unsigned)(t - P.t_begin);
        const unsigned tag = delivery + 1u;
        const bool head_on = t >= P.t_head;
        const bool flow = P.teacher_forced || t <= P.t_head;   // nothing upstream paces the stages: use the acks
        for (int g = 0; g < P.n_groups; ++g) {
            if (TRACE) {
                trace_row = (P.trace && t == P.trace_t && rank == 0 && tid == 0)
                                ? P.trace + ((size_t)stage * P.G + g) * TRACE_EV : nullptr;
                trace_n = 0;
                if (trace_row) trace_row[trace_n++] = (long long)globaltimer();
            }
            stamp();
            // the unit after this one, in processing order (for the prefetches issued during the last layer)
            int ng = g + 1;
            long long nt = t;
            if (ng == P.n_groups) { ng = 0; ++nt; }
            const bool more = nt < P.t_end;

            // ---------------- stage input -> x1[n & 1] ----------------
            float* xin = x1 + (n & 1u) * blk;
            if (first_stage) {
                // embedding gather: x1[k][p] = E[q_{b,t}][k]   (EmbeddingIO, modules/io.py:148-154)
                const int p = warp, b = g * GB + p;
                if (!pf_ok) {
                    long long q = 0;
                    if (b < P.B) {
                        if (!P.teacher_forced && t > P.t_head) {   // produced by the sampler one step ago: poll the word
                            uint2 v;
                            dead |= !poll_word(reinterpret_cast<const uint2*>(P.samples + g * GB + p), tag, v, abort_flag);
                            q = (long long)v.x;
                        } else {
                            q = __ldcg(P.seq + (size_t)b * P.seq_stride + t);
                        }
                    }
                    q = q < 0 ? 0 : (q >= P.Q ? P.Q - 1 : q);
                    const float* row = P.E + (size_t)q * C;
                    for (int k = lane; k < C; k += 32) xin[blk_off(C, k, p)] = (b < P.B) ? __ldg(row + k) : 0.0f;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < KJ) xin[blk_off(C, lane + 32 * j, p)] = pf_e[j];
                }
#pragma unroll
                for (int s = 0; s < TS; ++s) sacc[s] = 0.0f;
            } else {
                const size_t box = ((size_t)stage * P.G + g) * 2 + (delivery & 1u);
                const uint4* mh = P.mail_h + box * (size_t)(blk / 2);
                const int n_lines = blk / 2;
                {
                    uint4 v0 = pf_line0, v1 = pf_line1;
                    if (!pf_ok) {
                        if (tid < n_lines) v0 = ld_poll_v4(mh + tid);
                        if (tid + NT < n_lines) v1 = ld_poll_v4(mh + tid + NT);
                    }
                    if (tid < n_lines) {
                        if (!(v0.y == tag && v0.w == tag)) dead |= !poll_line(mh + tid, tag, v0, abort_flag);
                        *reinterpret_cast<float2*>(xin + 2 * tid) = make_float2(__uint_as_float(v0.x), __uint_as_float(v0.z));
                    }
                    if (tid + NT < n_lines) {
                        if (!(v1.y == tag && v1.w == tag)) dead |= !poll_line(mh + tid + NT, tag, v1, abort_flag);
                        *reinterpret_cast<float2*>(xin + 2 * (tid + NT)) = make_float2(__uint_as_float(v1.x), __uint_as_float(v1.z));
                    }
                    for (int i = tid + 2 * NT; i < n_lines; i += NT) {
                        uint4 v = ld_poll_v4(mh + i);
                        if (!(v.y == tag && v.w == tag)) dead |= !poll_line(mh + i, tag, v, abort_flag);
                        *reinterpret_cast<float2*>(xin + 2 * i) = make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
                    }
                }
                if (has_skip) {
                    const uint2* ms = P.mail_s + (box * CS + rank) * (size_t)(n_skip_tasks * 16);
#pragma unroll
                    for (int s = 0; s < TS; ++s) {
                        const int task = warp + NW * s;
                        if (task < n_skip_tasks) {
                            uint2 v = pf_skip[s];
                            if (!pf_ok) v = ld_poll_v2(ms + task * 16 + lane16);
                            if (v.y != tag) dead |= !poll_word(ms + task * 16 + lane16, tag, v, abort_flag);
                            sacc[s] = __uint_as_float(v.x);
                        }
                    }
                }
            }
            pf_ok = false;
            if (__syncthreads_or(dead ? 1 : 0)) goto done;
            if (!first_stage && flow && tid == 0) red_add_u32(P.ack + stage * P.G + g, 1u);
            stamp();
            bool x_local = true;   // x1[n & 1] was filled by this CTA itself: no mbarrier to wait on

            // ---------------- owned layers ----------------
            for (int l = l_lo; l < l_hi; ++l, ++n) {
                const bool has_res = ly_s[l - l_lo].has_res != 0;
                const float* W = w_s + (size_t)(l - l_lo) * P.layer_block;
                const bool last_owned = (l == l_hi - 1), last_layer = (l == P.L - 1);
                const bool to_mail = last_owned && !last_layer;
                const unsigned nb = n & 1u, nb1 = nb ^ 1u;
                const float* xc = x1 + nb * blk;
                const float* yc = ybuf + nb * blk;
                const unsigned off_xn = (unsigned)(P.s_x1 + nb1 * blk) * 4u;   // next layer's input buffer
                const unsigned off_y = (unsigned)(P.s_y + nb * blk) * 4u;
                const size_t box_out = ((size_t)(stage + 1) * P.G + g) * 2 + (delivery & 1u);
                uint4* mh_out = P.mail_h + box_out * (size_t)(blk / 2);

                // ---- phase A: both taps on h_l(t); newer tap + parked older tap -> gate -> all-gather y
                if (!x_local) {
                    dead |= !mbar_wait(bar(BAR_X0 + nb), (par_x >> nb) & 1u, abort_flag);
                    par_x ^= 1u << nb;
                    if (tid == 0) mbar_expect_tx(bar(BAR_X0 + nb), xbytes);   // arm the buffer's next use
                }
                stamp();
                if (to_mail && flow && !has_res) {
                    // y goes straight to the next stage's mailbox: its previous delivery there must have been consumed
                    if (delivery >= 2u && lane == 0)
                        dead |= !ack_wait(P.ack + (stage + 1) * P.G + g, (delivery - 1u) * CS, abort_flag);
                    __syncwarp();
                }
                float park[TS];
#pragma unroll
                for (int s = 0; s < TS; ++s) {
                    park[s] = 0.0f;
                    const int task = warp + NW * s;
                    if (task < nf) {
                        float acc[4][8];
                        dot_gate(W + P.o_ta + (size_t)task * KJ * 128, xc, C, KJ, acc);
                        const float a = tree2(acc[0], acc[1]) + a0[s];
                        const float v = gate_act(a, my_row != 0);
                        const float y = v * __shfl_xor_sync(0xffffffffu, v, 16);   // tanh(f) * sigmoid(g), wavenet_v2.py:151
                        const float4 c4 = gather4(y, 0, h4);
                        const int ch = rank * nf + task;
                        const unsigned coff = (unsigned)((h4 * C + ch) * 4) * 4u;
                        if (peer_y < CS) {
                            if (has_res || has_skip || last_layer)   // last layer without skips: y is the head input
                                st_async_v4(win_y + sbase + off_y + coff, c4, win_y + bar(BAR_Y0 + nb));
                            if (!has_res && !last_layer && !last_owned)        // h_{l+1} = y
                                st_async_v4(win_y + sbase + off_xn + coff, c4, win_y + bar(BAR_X0 + nb1));
                        }
                        if (!has_res && to_mail && lane < 2) {
                            const int m = h4 * C + ch;
                            st_flagged_v4(mh_out + 2 * m, c4.x, c4.y, tag);
                            st_flagged_v4(mh_out + 2 * m + 1, c4.z, c4.w, tag);
                        }
                        // older tap of this input, consumed at t + dilation
                        park[s] = tree2(acc[2], acc[3]) + W[P.o_b0 + task * 2 + my_row];
                    }
                }
                stamp();

                // ---- bubble (y is in flight): residual operands, park the older tap, fetch the next unit's operands
                float hreg[TS];
                const int nB = (has_skip ? ns : 0) / 2 + (has_res ? nr : 0) / 2;
#pragma unroll
                for (int s = 0; s < TS; ++s) {
                    hreg[s] = 0.0f;
                    const int task = warp + NW * s;
                    if (has_res && task >= (has_skip ? n_skip_tasks : 0) && task < nB) {
                        const int ch = rank * nf + 2 * (task - (has_skip ? n_skip_tasks : 0)) + my_row;
                        hreg[s] = xc[blk_off(C, ch, my_p)];
                    }
                }
                {
#pragma unroll
                    for (int s = 0; s < TS; ++s) {
                        const int task = warp + NW * s;
                        if (task < nf) __stcg(ring_ptr(l, t, g, task), park[s]);   // read back by this same lane at t + dilation
                    }
                    // next layer-unit's parked values (issued after the store: with one layer, one group and
                    // dilation 1 it is the very word just written)
                    const int nl = last_owned ? l_lo : l + 1;
                    const int g2 = last_owned ? ng : g;
                    const long long t2 = last_owned ? nt : t;
#pragma unroll
                    for (int s = 0; s < TS; ++s) {
                        const int task = warp + NW * s;
                        if (task < nf && t2 < P.t_end) a0[s] = __ldcg(ring_ptr(nl, t2, g2, task));
                    }
                    if (last_owned && more) {
                        // next unit's stage input: mailbox lines / skip words, or (first stage) the embedding row
                        if (!first_stage) {
                            const size_t boxn = ((size_t)stage * P.G + ng) * 2 + ((unsigned)(nt - P.t_begin) & 1u);
                            const uint4* mhn = P.mail_h + boxn * (size_t)(blk / 2);
                            if (tid < blk / 2) pf_line0 = ld_poll_v4(mhn + tid);
                            if (tid + NT < blk / 2) pf_line1 = ld_poll_v4(mhn + tid + NT);
                            if (has_skip) {
                                const uint2* msn = P.mail_s + (boxn * CS + rank) * (size_t)(n_skip_tasks * 16);
#pragma unroll
                                for (int s = 0; s < TS; ++s) {
                                    const int task = warp + NW * s;
                                    if (task < n_skip_tasks) pf_skip[s] = ld_poll_v2(msn + task * 16 + lane16);
                                }
                            }
                            pf_ok = true;
                        } else if (KJ <= 4) {
                            const int p = warp, b = ng * GB + p;
                            long long q = 0;
                            bool known = true;
                            if (b < P.B) {
                                if (!P.teacher_forced && nt > P.t_head) {
                                    const uint2 v = ld_poll_v2(reinterpret_cast<const uint2*>(P.samples + ng * GB + p));
                                    known = v.y == (unsigned)(nt - P.t_begin) + 1u;
                                    q = (long long)v.x;
                                } else {
                                    q = __ldcg(P.seq + (size_t)b * P.seq_stride + nt);
                                }
                            }
                            if (known) {
                                q = q < 0 ? 0 : (q >= P.Q ? P.Q - 1 : q);
                                const float* row = P.E + (size_t)q * C;
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (j < KJ) pf_e[j] = (b < P.B) ? __ldg(row + lane + 32 * j) : 0.0f;
                            }
                            pf_ok = known;
                        }
                    }
                }
                stamp();

                // ---- phase B: skip and residual 1x1 convs on the gathered y
                if (nB > 0 || (!has_skip && last_layer)) {
                    dead |= !mbar_wait(bar(BAR_Y0 + nb), (par_y >> nb) & 1u, abort_flag);
                    par_y ^= 1u << nb;
                    if (tid == 0) mbar_expect_tx(bar(BAR_Y0 + nb), xbytes);
                }
                stamp();
                if (to_mail && flow && has_res) {
                    if (delivery >= 2u && lane == 0)
                        dead |= !ack_wait(P.ack + (stage + 1) * P.G + g, (delivery - 1u) * CS, abort_flag);
                    __syncwarp();
                }
#pragma unroll
                for (int s = 0; s < TS; ++s) {
                    const int task = warp + NW * s;
                    if (task < nB) {
                        float acc[2][8];
                        dot_rows<2>(W + P.o_tb + (size_t)task * KJ * 64, nullptr, yc, C, KJ, acc);
                        float v = tree2(acc[0], acc[1]) + W[P.o_bb + task * 2 + my_row];
                        const bool is_skip = has_skip && task < n_skip_tasks;
                        if (is_skip) {
                            v = (l == 0) ? v : (v + sacc[s]);               // skips = conv_skip(y) + skips (wavenet_v2.py:165-171)
                            sacc[s] = v;
                            if (last_layer && head_on) {                      // all-gather the head input
                                const float4 c4 = gather4(v, ch4 >> 1, ch4 & 1);
                                const int ch = rank * ns + 2 * task + (ch4 >> 1);
                                send2(sbase + (unsigned)P.s_hin * 4u + (unsigned)(((ch4 & 1) * S + ch) * 4) * 4u, c4, BAR_HI);
                            }
                            if (to_mail && (lane & 8) == 0) {
                                uint2* ms = P.mail_s + (box_out * CS + rank) * (size_t)(n_skip_tasks * 16);
                                st_flagged_v2(ms + task * 16 + lane16, __float_as_uint(v), tag);
                            }
                        } else {
                            v = hreg[s] + v;                                  // h_{l+1} = h_l + conv_res(y) (wavenet_v2.py:172-175)
                            const float4 c4 = gather4(v, ch4 >> 1, ch4 & 1);
                            const int ch = rank * nf + 2 * (task - (has_skip ? n_skip_tasks : 0)) + (ch4 >> 1);
                            const int m = (ch4 & 1) * C + ch;
                            if (!last_owned) {
                                send2(sbase + off_xn + (unsigned)m * 16u, c4, BAR_X0 + nb1);
                            } else if (to_mail && lane < 4) {
                                st_flagged_v4(mh_out + 2 * m, c4.x, c4.y, tag);
                                st_flagged_v4(mh_out + 2 * m + 1, c4.z, c4.w, tag);
                            }
                        }
                    }
                }
                stamp();
                x_local = false;
            }  // layers

            // ---------------- head + sampler (last stage) ----------------
            if (last_stage && head_on) {
                const float* H = head_s;
                const float* head_x = hin;
                if (has_skip) {
                    dead |= !mbar_wait(bar(BAR_HI), cnt_hi & 1u, abort_flag);
                    ++cnt_hi;
                    if (tid == 0) mbar_expect_tx(bar(BAR_HI), (unsigned)P.Kh * GB * 4u);
                } else {
                    head_x = ybuf + ((n - 1u) & 1u) * blk;   // gathered y of the last layer (its barrier was waited above)
                }
                stamp();
#pragma unroll
                for (int s = 0; s < TS; ++s) {                 // hidden = mish(W1 x + b1), all-gathered (mlp.py:44-53)
                    const int task = warp + NW * s;
                    if (task < P.nh / 2) {
                        float acc[2][8];
                        dot_rows<2>(H + P.o_h1 + (size_t)task * P.KJh * 64, nullptr, head_x, P.Kh, P.KJh, acc);
                        float v = tree2(acc[0], acc[1]) + H[P.o_h1b + task * 2 + my_row];
                        v = mish_acc(v);
                        const float4 c4 = gather4(v, ch4 >> 1, ch4 & 1);
                        const int ch = rank * P.nh + 2 * task + (ch4 >> 1);
                        send2(sbase + (unsigned)P.s_hid * 4u + (unsigned)(((ch4 & 1) * P.Hh + ch) * 4) * 4u, c4, BAR_HD);
                    }
                }
                dead |= !mbar_wait(bar(BAR_HD), cnt_hd & 1u, abort_flag);
                ++cnt_hd;
                if (tid == 0) mbar_expect_tx(bar(BAR_HD), (unsigned)P.Hh * GB * 4u);
                stamp();
                {
                    const unsigned z0 = win_0 + sbase + (unsigned)P.s_z * 4u, zb = win_0 + bar(BAR_Z);
#pragma unroll
                    for (int s = 0; s < TS; ++s) {             // z = W2 hidden + b2 -> rank 0, [prompt][logit]
                        const int task = warp + NW * s;
                        if (task < P.nz / 2) {
                            const bool extra = (rank == CS - 1) && (task == P.nz / 2 - 1);   // + the temperature row Q
                            float v, vx = 0.0f;
                            if (extra) {
                                float acc[3][8];
                                dot_rows<3>(H + P.o_h2 + (size_t)task * P.KJ2 * 64, H + P.o_h2x, hid, P.Hh, P.KJ2, acc);
                                v = tree2(acc[0], acc[1]);
                                vx = tree1(acc[2]) + H[P.o_h2xb];
                            } else {
                                float acc[2][8];
                                dot_rows<2>(H + P.o_h2 + (size_t)task * P.KJ2 * 64, nullptr, hid, P.Hh, P.KJ2, acc);
                                v = tree2(acc[0], acc[1]);
                            }
                            v += H[P.o_h2b + task * 2 + my_row];
                            const int o = rank * P.nz + 2 * task + my_row;
                            if ((lane & 8) == 0) st_async_f32(z0 + (unsigned)(my_p * P.zrow + o) * 4u, v, zb);
                            else if (extra && lane < 16) st_async_f32(z0 + (unsigned)(my_p * P.zrow + P.Q) * 4u, vx, zb);
                        }
                    }
                }
                stamp();
                if (rank == 0) {
                    dead |= !mbar_wait(bar(BAR_Z), cnt_z & 1u, abort_flag);
                    ++cnt_z;
                    if (tid == 0) mbar_expect_tx(bar(BAR_Z), (unsigned)(P.Q + 1) * GB * 4u);
                    stamp();
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
                        const int choice = mmk::decide_warp(zbuf + p * P.zrow, P.Q, P.min_temp, lout, sample, Tt, u);
                        if (lane == 0) {
                            if (P.decisions) P.decisions[(size_t)b * n_head + hstep] = choice;
                            if (!P.teacher_forced) {
                                st_flagged_v2(reinterpret_cast<uint2*>(P.samples + g * GB + p), (unsigned)choice, tag + 1u);
                                __stcg(P.seq + (size_t)b * P.seq_stride + t + 1, (long long)choice);
                            }
                        }
                    }
                    stamp();
                }
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

}  // namespace mmk3

using namespace mmk3;

struct wn3_handle {
    Params p{};
    int device = 0, max_batch = 0, rf = 0, ts = 1;
    size_t smem_bytes = 0;
    std::vector<void*> allocs;
    void* d_flags = nullptr;        // mailboxes + sample words + acks + abort flag: cleared before every launch
    size_t flags_bytes = 0;
    long long* d_trace = nullptr;   // debug timeline, only with MMK_WN_TRACE_T set
    long long trace_t = -1;
};

static const void* wn3_kernel(int ts, bool trace = false) {
    if (trace) {   // debug timeline build of the same kernel (MMK_WN_TRACE_T)
        switch (ts) {
            case 1: return (const void*)wavenet_warp_kernel<1, true>;
            case 2: return (const void*)wavenet_warp_kernel<2, true>;
            default: return (const void*)wavenet_warp_kernel<4, true>;
        }
    }
    switch (ts) {
        case 1: return (const void*)wavenet_warp_kernel<1, false>;
        case 2: return (const void*)wavenet_warp_kernel<2, false>;
        default: return (const void*)wavenet_warp_kernel<4, false>;
    }
}

// fills geometry + weight block offsets + smem carve-up; returns dynamic smem bytes
static size_t wn3_plan(Params& p, int CS, int layers_per_stage) {
    p.CS = CS;
    p.nf = p.C / CS; p.ns = p.S / CS; p.nr = p.nf; p.nh = p.Hh / CS; p.nz = p.Q / CS;
    p.KJ = p.C / 32; p.KJh = p.Kh / 32; p.KJ2 = p.Hh / 32;
    p.blk = p.C * GB;
    int o = 0;
    auto take = [&](int floats) { int r = o; o += pad4(floats); return r; };
    p.o_ta = take(p.nf * p.KJ * 128); p.o_b0 = take(p.nf * 2);
    const int nB = (p.ns + p.nr) / 2;
    p.o_tb = take(nB * p.KJ * 64); p.o_bb = take(std::max(2, nB * 2));
    p.layer_block = o;
    o = 0;
    p.o_h1 = take((p.nh / 2) * p.KJh * 64); p.o_h1b = take(p.nh);
    p.o_h2 = take((p.nz / 2) * p.KJ2 * 64); p.o_h2b = take(p.nz);
    p.o_h2x = take(p.KJ2 * 32); p.o_h2xb = take(1);
    p.head_block = o;
    o = 0;
    p.zrow = pad4(p.Q + 1) + 4;
    if ((p.zrow % 32) == 0) p.zrow += 4;
    p.s_w = take(layers_per_stage * p.layer_block);
    p.s_head = take(p.head_block);
    p.s_x1 = take(2 * p.blk); p.s_y = take(2 * p.blk);
    p.s_hin = take(p.Kh * GB); p.s_hid = take(p.Hh * GB);
    p.s_z = take(GB * p.zrow);
    p.s_bar = take(BAR_COUNT * 2);
    p.s_ly = take(MAX_OWN * (int)(sizeof(LayerS) / sizeof(float)));
    p.smem_floats = o;
    return (size_t)o * sizeof(float);
}

static int wn3_query_clusters(const void* k, int CS, size_t smem, int* out) {
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

int wn3_destroy(wn3_handle* h) {
    if (!h) return 0;
    for (void* a : h->allocs) cudaFree(a);
    delete h;
    return 0;
}

int wn3_create(const mmk_wavenet_desc* d, int max_batch, wn3_handle** out, int* unsupported) {
    *unsupported = 1;
    if (d->dilated_dim % 32 || d->head_hidden % 32 || (d->skips_dim % 32) || d->n_layers > MAX_LAYERS) return 1;
    // residual convs: on every layer but the last, or on none (wavenet_v2.py:78,216)
    bool any_res = false, all_res = true;
    for (int l = 0; l < d->n_layers - 1; ++l) { any_res |= d->conv_res_w[l] != nullptr; all_res &= d->conv_res_w[l] != nullptr; }
    if (any_res && !all_res) return 1;
    auto* h = new wn3_handle();
    Params& p = h->p;
    cudaGetDevice(&h->device);
    p.L = d->n_layers; p.C = d->dilated_dim; p.S = d->skips_dim; p.Hh = d->head_hidden; p.Q = d->q_levels;
    p.Kh = p.S > 0 ? p.S : p.C;
    p.min_temp = d->min_temperature;
    h->max_batch = max_batch;
    p.G = (max_batch + GB - 1) / GB;
    int rf = 1;
    for (int l = 0; l < p.L; ++l) rf += d->dilations[l];
    h->rf = rf;
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);

    int best_cs = 0, best_nst = 0, best_ts = 0;
    const char* force_cs = getenv("MMK_WN_CLUSTER");
    const char* force_nst = getenv("MMK_WN_STAGES");
    for (int CS : {16, 8, 4, 2, 1}) {
        if (force_cs && atoi(force_cs) != CS) continue;
        if (p.C % (2 * CS) || p.S % (2 * CS) || p.Hh % (2 * CS) || p.Q % (2 * CS)) continue;   // whole row pairs per CTA
        const int nf = p.C / CS, ns = p.S / CS, nh = p.Hh / CS, nz = p.Q / CS;
        const int most = std::max(std::max(nf, (ns + nf) / 2), std::max(nh / 2, nz / 2));
        const int ts = (most + NW - 1) / NW;
        if (ts > MAX_TS) continue;
        const int tsi = ts <= 1 ? 1 : (ts <= 2 ? 2 : 4);
        const void* kern = wn3_kernel(tsi);
        int nst_min = 0;
        for (int nst = 1; nst <= std::min(p.L, MAX_STAGES); ++nst) {
            Params q = p;
            if ((p.L + nst - 1) / nst <= MAX_OWN && wn3_plan(q, CS, (p.L + nst - 1) / nst) <= (size_t)max_optin) { nst_min = nst; break; }
        }
        if (!nst_min) continue;
        Params q = p;
        const size_t smem_min = wn3_plan(q, CS, (p.L + nst_min - 1) / nst_min);
        int max_clusters = 0;
        if (CS == 1) {
            int per_sm = 0, sms = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_min);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem_min);
            max_clusters = per_sm * sms;
        } else if (wn3_query_clusters(kern, CS, smem_min, &max_clusters)) { wn3_destroy(h); *unsupported = 0; return 1; }
        if (max_clusters < nst_min) continue;
        int nst = std::min(std::min(max_clusters, p.L), MAX_STAGES);
        if (force_nst) nst = std::max(nst_min, std::min(nst, atoi(force_nst)));
        else nst = std::min(nst, std::max(nst_min, std::max(1, p.G)));
        best_cs = CS; best_nst = nst; best_ts = tsi;
        break;
    }
    if (!best_cs) { wn3_destroy(h); return 1; }
    const int CS = best_cs, NST = best_nst;
    h->ts = best_ts;
    p.NST = NST;
    const int per = (p.L + NST - 1) / NST;
    h->smem_bytes = wn3_plan(p, CS, per);
    {
        int base = p.L / NST, extra = p.L % NST, lo = 0;
        for (int s = 0; s < NST; ++s) { p.stage_lo[s] = lo; lo += base + (s < extra ? 1 : 0); }
        p.stage_lo[NST] = p.L;
    }
    *unsupported = 0;
    for (int tr = 0; tr < 2; ++tr) {
        const void* kern = wn3_kernel(h->ts, tr != 0);
        MMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        if (CS > 8) MMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }

    // ---- pack weights: task-major, W[task][k][2] = (row A, row B) at contraction index k
    const int C = p.C, S = p.S, nf = p.nf, ns = p.ns, KJ = p.KJ;
    std::vector<float> wpack((size_t)p.L * CS * p.layer_block, 0.0f);
    auto put_pair = [&](float* dst, int task, int K, auto weight_of /* (row r in {0,1}, k) -> float */) {
        for (int k = 0; k < K; ++k)
            for (int r = 0; r < 2; ++r) dst[((size_t)task * K + k) * 2 + r] = weight_of(r, k);
    };
    long long ring_off = 0;
    for (int l = 0; l < p.L; ++l) {
        const bool has_res = d->conv_res_w[l] != nullptr;
        p.layers[l].dilation = d->dilations[l];
        p.layers[l].has_res = has_res ? 1 : 0;
        p.layers[l].ring_off = ring_off;
        ring_off += (long long)d->dilations[l] * p.G * CS * nf * 16;
        const float* wd = d->conv_dil_w[l];   // (2C, C, 2): [o][c][tap], tap 0 = older sample
        const float* bd = d->conv_dil_b[l];
        for (int r = 0; r < CS; ++r) {
            float* blkp = wpack.data() + ((size_t)l * CS + r) * p.layer_block;
            for (int i = 0; i < nf; ++i) {
                // task rows: filter row (channel), gate row (C + channel)
                auto orow = [&](int rr) { return (rr ? C : 0) + r * nf + i; };
                for (int k = 0; k < C; ++k)       // (f newer, g newer, f older, g older) at contraction index k
                    for (int q = 0; q < 4; ++q)
                        blkp[p.o_ta + ((size_t)i * C + k) * 4 + q] = wd[((size_t)orow(q & 1) * C + k) * 2 + (q < 2 ? 1 : 0)];
                for (int rr = 0; rr < 2; ++rr) blkp[p.o_b0 + i * 2 + rr] = bd[orow(rr)];
            }
            int task = 0;
            if (S > 0)
                for (int i = 0; i < ns / 2; ++i, ++task) {
                    put_pair(blkp + p.o_tb, task, C, [&](int rr, int k) { return d->conv_skip_w[l][(size_t)(r * ns + 2 * i + rr) * C + k]; });
                    for (int rr = 0; rr < 2; ++rr) blkp[p.o_bb + task * 2 + rr] = d->conv_skip_b[l][r * ns + 2 * i + rr];
                }
            if (has_res)
                for (int i = 0; i < nf / 2; ++i, ++task) {
                    put_pair(blkp + p.o_tb, task, C, [&](int rr, int k) { return d->conv_res_w[l][(size_t)(r * nf + 2 * i + rr) * C + k]; });
                    for (int rr = 0; rr < 2; ++rr) blkp[p.o_bb + task * 2 + rr] = d->conv_res_b[l][r * nf + 2 * i + rr];
                }
        }
    }
    std::vector<float> hpack((size_t)CS * p.head_block, 0.0f);
    for (int r = 0; r < CS; ++r) {
        float* blkp = hpack.data() + (size_t)r * p.head_block;
        for (int i = 0; i < p.nh / 2; ++i) {
            put_pair(blkp + p.o_h1, i, p.Kh, [&](int rr, int k) { return d->head_w1[(size_t)(r * p.nh + 2 * i + rr) * p.Kh + k]; });
            for (int rr = 0; rr < 2; ++rr) blkp[p.o_h1b + i * 2 + rr] = d->head_b1[r * p.nh + 2 * i + rr];
        }
        for (int i = 0; i < p.nz / 2; ++i) {
            put_pair(blkp + p.o_h2, i, p.Hh, [&](int rr, int k) { return d->head_w2[(size_t)(r * p.nz + 2 * i + rr) * p.Hh + k]; });
            for (int rr = 0; rr < 2; ++rr) blkp[p.o_h2b + i * 2 + rr] = d->head_b2[r * p.nz + 2 * i + rr];
        }
        for (int k = 0; k < p.Hh; ++k) blkp[p.o_h2x + k] = d->head_w2[(size_t)p.Q * p.Hh + k];   // learned-temperature row
        blkp[p.o_h2xb] = d->head_b2[p.Q];
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
    const size_t boxes = (size_t)(NST + 1) * p.G * 2;
    const size_t mh_bytes = boxes * (p.blk / 2) * sizeof(uint4);
    const size_t ms_bytes = boxes * CS * std::max(1, p.ns / 2) * 16 * sizeof(uint2);
    const size_t sm_bytes = (size_t)p.G * GB * sizeof(unsigned long long);
    const size_t ack_bytes = ((size_t)(NST + 1) * p.G + 4) * sizeof(unsigned);
    h->flags_bytes = mh_bytes + ms_bytes + sm_bytes + ack_bytes;
    h->d_flags = dev_alloc(h->flags_bytes, nullptr);
    if (const char* e = getenv("MMK_WN_TRACE_T")) {
        h->trace_t = atoll(e);
        h->d_trace = (long long*)dev_alloc((size_t)NST * p.G * TRACE_EV * sizeof(long long), nullptr);
    }
    if (!ok) { wn3_destroy(h); MMK_FAIL("cudaMalloc failed while creating the WaveNet handle"); }
    char* f = (char*)h->d_flags;
    p.mail_h = (uint4*)f; f += mh_bytes;
    p.mail_s = (uint2*)f; f += ms_bytes;
    p.samples = (unsigned long long*)f; f += sm_bytes;
    p.ack = (unsigned*)f;
    p.abort_flag = p.ack + (size_t)(NST + 1) * p.G;
    MMK_CUDA(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

int wn3_launch_info(wn3_handle* h, mmk_launch_info* out) {
    out->cluster_size = h->p.CS; out->n_stages = h->p.NST; out->group_size = GB; out->threads = NT;
    out->smem_bytes = (int)h->smem_bytes; out->sm_used = h->p.CS * h->p.NST;
    return 0;
}

int wn3_sync_check(wn3_handle* h, void* stream) {
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

int wn3_run(wn3_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
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
    MMK_CUDA(cudaMemsetAsync(h->d_flags, 0, h->flags_bytes, st));   // tags of an earlier launch must not match
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
    MMK_CUDA(cudaLaunchKernelExC(&cfg, wn3_kernel(h->ts, h->d_trace != nullptr), args));
    return 0;
}
