// Layer-pipelined persistent fp32 WaveNet generation kernel (sm_100a) — the default fp32 path.
// Function computed: WaveNet.generate_step / forward (mimikit/networks/wavenet_v2.py:447-452, 276-293), WNLayer.forward
// (131-176), EmbeddingIO (modules/io.py:148-154), MLP head + learned temperature (networks/mlp.py:44-63),
// CategoricalSampler (modules/targets.py:40-52), driven as GenerateLoopV2.run does (loops/generate.py:184-229).
//
// A generated sample is a chain of L layers x 2 dependent contractions + head + sampler; a step costs L x (latency of one
// layer for one prompt group) whatever the batch.  ncu on the previous design (every contraction split by row over a
// 16-CTA cluster, two 16-way all-gathers per layer) showed 2.6 us per layer of which 0.1 us was FFMA issue: the chain was
// made of exchange latency and per-phase fixed costs.  This kernel gives every layer its OWN pair of CTAs instead:
//   * CTA (l, h) owns half h of layer l for the whole launch: gate channels [h C/2, (h+1) C/2) (filter and gate rows) and
//     the residual / skip rows of the same range.  Groups of 4 prompts flow down the pipeline of L pairs + the head CTAs;
//     with B = 64 there are 16 groups in a 31-deep pipeline, so no group ever queues and the step time is the chain.
//   * the weights of the two contractions on the critical path (newer conv tap, residual 1x1) live in REGISTERS for the
//     whole launch (48 per thread at C = 128: the register file is the largest on-chip memory); the two off the
//     critical path (older conv tap for t + dilation, skip 1x1) live in shared memory and run after the hand-off.
//   * thread = (channel quad q, k-slice s of 16), 2 C threads per CTA: 8 gate rows (or 4 residual / skip rows) x 4 prompts
//     x C/16 contraction steps, the layer input read as one LDS.128 per step.  (A warp-wide LDS.128 costs four
//     shared-memory wavefronts whatever the lanes share, so the activation reads — not the FFMAs — set the pace: the
//     first version, 8 slices x 2 rows per thread and 4 C threads, spent 4x the wavefronts on them.)  A 16-lane
//     transposing shuffle tree leaves (filter, gate) of one (channel, prompt) in each lane: tanh * sigmoid is local and
//     every lane sends exactly one value.
//   * per layer there are two point-to-point DSMEM hops (y halves swapped inside the pair, the new layer input sent to
//     both CTAs of the next pair) as st.async stores that complete_tx on the receiver's mbarrier; buffers are double
//     buffered and returned with remote mbarrier arrivals (credits).  Pairs in different clusters talk through
//     self-flagged 8-byte words {value, tag} in L2.  No CTA-wide or cluster-wide barrier inside a unit.
//   * the older tap is applied when a layer input is produced, bias included, and parked in a thread-private L2 ring
//     slot for t + dilation; the running skip sum travels down the pipeline with its group.
#include "common.cuh"
#include "sampler.cuh"
#include "wavenet_impl.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace mmk6 {

using mmk::mish_acc;

constexpr int GB = 4;             // prompts per pipeline group
constexpr int KS = 16;            // k-slices (lanes that share one output row)
constexpr int MAX_LAYERS = 96;
constexpr int MAX_HC = 8;         // head CTAs
constexpr int TRACE_EV = 24;

struct Layer {
    int dilation, has_res;
    long long ring_off;           // float offset of this layer's parked-tap ring
};

struct Params {
    int L, C, S, Hh, Q, Kh, CS, NPC, NCL, NHC, G;
    int head_slot;                // first pair slot of the head CTAs
    int nh1, nz, nzp;             // head: hidden rows / logit rows (padded to 4) per head CTA
    int skip_passes;              // ceil((S/2) / (C/2))
    int o_wold, o_wsk, o_bias, layer_block;      // float offsets inside a (layer, half) shared-memory weight block
    int o_w1, o_w2, o_b1, o_b2, o_wt, head_block;      // ... inside a head CTA's block (o_wt: temperature row + its bias)
    int s_in, inblk, s_y, yblk, s_sk, skblk, s_bar, smem_floats, zrow;
    float min_temp;
    Layer layers[MAX_LAYERS];
    const float* wreg;            // [L][2][3 * C/16][2C] float4: register-resident weights, thread-major
    const float* wsm;             // [L][2][layer_block]
    const float* hpack;           // [NHC][head_block]
    const float* E;
    float* rings;
    uint2* mail_x;                // [(slot, g, parity)][max(C, Kh) * GB] words {value, tag}: layer inputs across clusters
    uint2* mail_s;                // [(slot, g, parity)][S * GB] words: running skip sums across clusters
    unsigned long long* samples;  // [G][GB] words {index, tag}
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

// barrier roles (layer CTA | head CTA)
enum { BAR_IN = 0 /* +buf: layer input | head input */, BAR_Y = 2 /* gated output | head hidden */,
       BAR_SK = 4 /* incoming skip sum | logits */, BAR_FREE = 6 /* credits from the consumers */,
       BAR_ZDONE = 8 /* head: the deciding warps are done with a logits buffer */, BAR_COUNT = 10 };

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
// Credit return.  Relaxed on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU (0.6 stalled warps per issue in ncu) and
// nothing needs publishing — the credit only says that this warp's shared-memory READS of the buffer are over, and they are
// (their values were consumed by the arithmetic whose results were sent before this instruction issues).
__device__ __forceinline__ void mbar_arrive_remote(unsigned raddr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void st_async_f32(unsigned raddr, float v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
                 ::"r"(raddr), "f"(v), "r"(rbar) : "memory");
}
// CTA barrier `id` over `nthreads` threads with an OR of a predicate
__device__ __forceinline__ bool named_sync_or(int id, int nthreads, bool pred) {
    unsigned r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"((unsigned)(pred ? 1u : 0u)), "r"(id), "r"(nthreads) : "memory");
    return r != 0u;
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
__device__ __forceinline__ uint2 ld_poll_v2(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
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
__device__ __noinline__ bool poll_word(const uint2* p, unsigned tag, uint2& v, unsigned* abort_flag) {
    return spin_until([&] { v = ld_poll_v2(p); return v.y == tag; }, abort_flag);
}

// ------------------------------------------------------------------------------------------------------------
// Transposing reductions, fixed summation order.  fold<N>(a, bit, mask): lanes with `bit` clear keep the lower half of the
// N partial sums, lanes with it set the upper half; each adds the partner's partial sums of the half it keeps.
// ------------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void fold(const float (&a)[N], float (&o)[N / 2], bool bit, int mask) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float keep = bit ? a[N / 2 + i] : a[i], send = bit ? a[i] : a[N / 2 + i];
        o[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
}
// The same without the selects, for levels whose halves were swapped ahead of time: the weights are per-lane data, so a
// lane whose bit is set simply holds its rows in the swapped order (row slot j = row j ^ lane bits, see the packing).
template <int N>
__device__ __forceinline__ void fold_swapped(const float (&a)[N], float (&o)[N / 2], int mask) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) o[i] = a[i] + __shfl_xor_sync(0xffffffffu, a[N / 2 + i], mask);
}
// 32 partial sums per lane over the 16 lanes of a k-slice group -> lane (b3 b2 b1 b0) keeps values 2 * lane16 + {0, 1}
__device__ __forceinline__ float2 tree32x16(const float (&a)[32]) {
    const int lane = threadIdx.x & 31;
    float v16[16], v8[8], v4[4], v2[2];
    fold_swapped<32>(a, v16, 8);
    fold_swapped<16>(v16, v8, 4);
    fold<8>(v8, v4, lane & 2, 2);
    fold<4>(v4, v2, lane & 1, 1);
    return make_float2(v2[0], v2[1]);
}
// 16 partial sums per lane over 16 lanes -> lane16 keeps value lane16
__device__ __forceinline__ float tree16x16(const float (&a)[16]) {
    const int lane = threadIdx.x & 31;
    float v8[8], v4[4], v2[2], v1[1];
    fold_swapped<16>(a, v8, 8);
    fold_swapped<8>(v8, v4, 4);
    fold<4>(v4, v2, lane & 2, 2);
    fold<2>(v2, v1, lane & 1, 1);
    return v1[0];
}
// 16 partial sums per lane over the 32 lanes of a warp -> lanes 2 i and 2 i + 1 keep value i
__device__ __forceinline__ float tree16x32(const float (&a)[16]) {
    const int lane = threadIdx.x & 31;
    float v8[8], v4[4], v2[2], v1[1];
    fold_swapped<16>(a, v8, 16);
    fold_swapped<8>(v8, v4, 8);
    fold<4>(v4, v2, lane & 4, 4);
    fold<2>(v2, v1, lane & 2, 2);
    return v1[0] + __shfl_xor_sync(0xffffffffu, v1[0], 1);
}

// tanh for the filter row (m = 2), sigmoid for the gate row (m = 1):
// tanh(a) = 2 / (1 + exp(-2a)) - 1,  sigmoid(a) = 1 / (1 + exp(-a))
__device__ __forceinline__ float gate_act(float a, bool gate_half) {
    const float m = gate_half ? 1.0f : 2.0f;
    const float e = expf(-m * a);
    return __fdiv_rn(m, 1.0f + e) - (m - 1.0f);
}

// tanh(f) * sigmoid(g) with one division: t = exp(-2f), u = exp(-g): (1 - t) / ((1 + t) (1 + u)).  f is clamped at -30
// (tanh is -1 to the last bit long before) so that t stays finite; u = inf gives 0, the limit of the sigmoid.
__device__ __forceinline__ float gated(float f, float g) {
    const float t = expf(-2.0f * fmaxf(f, -30.0f)), u = expf(-g);
    return __fdiv_rn(1.0f - t, (1.0f + t) * (1.0f + u));
}

// acc[(i * 4 + p)] += w[i] * x[p] for 4 rows i, 4 prompts p
__device__ __forceinline__ void fma16(float* acc, const float4& w, const float4& x) {
    acc[0] = fmaf(w.x, x.x, acc[0]); acc[1] = fmaf(w.x, x.y, acc[1]); acc[2] = fmaf(w.x, x.z, acc[2]); acc[3] = fmaf(w.x, x.w, acc[3]);
    acc[4] = fmaf(w.y, x.x, acc[4]); acc[5] = fmaf(w.y, x.y, acc[5]); acc[6] = fmaf(w.y, x.z, acc[6]); acc[7] = fmaf(w.y, x.w, acc[7]);
    acc[8] = fmaf(w.z, x.x, acc[8]); acc[9] = fmaf(w.z, x.y, acc[9]); acc[10] = fmaf(w.z, x.z, acc[10]); acc[11] = fmaf(w.z, x.w, acc[11]);
    acc[12] = fmaf(w.w, x.x, acc[12]); acc[13] = fmaf(w.w, x.y, acc[13]); acc[14] = fmaf(w.w, x.z, acc[14]); acc[15] = fmaf(w.w, x.w, acc[15]);
}
// gate rows: acc[(i * 4 + p) * 2 + r] += w(i, r) * x[p]; wa = (f0, g0, f1, g1) of channels 0, 1; wb of channels 2, 3
__device__ __forceinline__ void fma_gate(float (&acc)[32], const float4& wa, const float4& wb, const float4& x) {
    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            acc[(i * 4 + p) * 2 + 0] = fmaf(w[i * 2 + 0], xv[p], acc[(i * 4 + p) * 2 + 0]);
            acc[(i * 4 + p) * 2 + 1] = fmaf(w[i * 2 + 1], xv[p], acc[(i * 4 + p) * 2 + 1]);
        }
}

// Head contraction: NCH chunks of 4 output rows x 4 prompts at once (independent accumulators hide the latencies of a lone
// warp), K split over the 32 lanes (k = lane + 32 j).  W4: chunk c at W4 + c * cstride, [K/32][32] float4 = the 4 rows at
// this lane's k; x4: [K] float4.  out[c]: lanes 2 i and 2 i + 1 hold output (row = i >> 2, prompt = i & 3) of chunk c.
template <int NCH>
__device__ __forceinline__ void head_rows(const float4* __restrict__ W4, int cstride, const float4* __restrict__ x4, int K,
                                          float (&out)[NCH]) {
    const int lane = threadIdx.x & 31;
    float acc[NCH][16];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[c][i] = 0.0f;
#pragma unroll 2
    for (int j = 0; j < K / 32; ++j) {
        const float4 x = x4[j * 32 + lane];
#pragma unroll
        for (int c = 0; c < NCH; ++c) fma16(acc[c], W4[(size_t)c * cstride + j * 32 + lane], x);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) out[c] = tree16x32(acc[c]);
}
// One row x 4 prompts, K over the lanes: every lane returns output (prompt = (lane >> 1) & 3 ... see tree): used for the
// learned-temperature row.  w: [K] floats of the row; returns in lanes 8 p .. 8 p + 7 the output of prompt p.
__device__ __forceinline__ float head_row1(const float* __restrict__ w, const float4* __restrict__ x4, int K) {
    const int lane = threadIdx.x & 31;
    float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int j = 0; j < K / 32; ++j) {
        const float4 x = x4[j * 32 + lane];
        const float wk = w[j * 32 + lane];
        a[0] = fmaf(wk, x.x, a[0]); a[1] = fmaf(wk, x.y, a[1]); a[2] = fmaf(wk, x.z, a[2]); a[3] = fmaf(wk, x.w, a[3]);
    }
    float v2[2], v1[1];
    fold<4>(a, v2, lane & 16, 16);
    fold<2>(v2, v1, lane & 8, 8);
    float v = v1[0];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// The kernel.  One CTA = half a layer (or a slice of the head) for the whole launch; 2 C threads.
// thread = (k-slice s of 16, channel quad q): 4 channels (8 gate rows / 4 residual rows / 4 skip rows) x C/16 steps.
// ------------------------------------------------------------------------------------------------------------
template <int C, bool TRACE>
__global__ void __maxnreg__(255) wavenet6_kernel(const __grid_constant__ Params P) {
    constexpr int NT = 2 * C, NW = NT / 32, KJ = C / KS, CH = C / 2, NQ = C / 8;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int s = tid & 15, q = tid >> 4;         // k-slice; channel quad inside this CTA's half
    const int lane16 = lane & 15;
    const unsigned rank = cluster_ctarank();
    const int cluster = blockIdx.x / P.CS;
    const int slot = cluster * P.NPC + (int)(rank >> 1);
    const int h = (int)(rank & 1u);
    const unsigned sbase = smem_u32(smem);
    unsigned* abort_flag = P.abort_flag;
    const int G = P.G, S = P.S, SH = S / 2, L = P.L;
    const bool has_skip = S > 0;
    auto bar = [&](int i) { return sbase + (unsigned)P.s_bar * 4u + 8u * (unsigned)i; };
    auto window = [&](int r) { return mapa(sbase, (unsigned)r) - sbase; };   // remote = window + local address
    auto slot_cluster = [&](int sl) { return sl / P.NPC; };
    auto slot_rank0 = [&](int sl) { return (sl % P.NPC) * 2; };
    const bool is_layer = slot < L;
    const int hr = (slot - P.head_slot) * 2 + h;                             // head CTA index
    const bool is_head = slot >= P.head_slot && hr < P.NHC;
    bool dead = false;

    float* inb = smem + P.s_in;       // [2][inblk]  layer input h_l(t)           | head input (skip sums)
    float* yb = smem + P.s_y;         // [2][yblk]   gated outputs of the pair    | head hidden
    float* skb = smem + P.s_sk;       // [2][skblk]  incoming running skip sums   | raw logits [GB][zrow]
    const unsigned in_bytes_layer = (unsigned)C * GB * 4u, y_bytes = (unsigned)C * GB * 4u;
    const unsigned sk_bytes = (unsigned)SH * GB * 4u;
    const unsigned hin_bytes = (unsigned)P.Kh * GB * 4u, hid_bytes = (unsigned)P.Hh * GB * 4u;
    const unsigned z_bytes = (unsigned)(P.Q + 1) * GB * 4u;

    // ---- barriers
    if (tid == 0) {
        const bool last_layer = is_layer && slot == L - 1;
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(BAR_IN + b), 1); mbar_init(bar(BAR_Y + b), 1); mbar_init(bar(BAR_SK + b), 1);
            // credits: every warp of every consumer CTA arrives once per consumed delivery
            mbar_init(bar(BAR_FREE + b), (unsigned)(last_layer ? P.NHC * (NW > GB ? NW - GB : NW) : 2 * NW));
            mbar_init(bar(BAR_ZDONE + b), GB);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int b = 0; b < 2; ++b) {   // arm the first phase of every exchange barrier (tx bytes may land before or after)
            if (is_layer) {
                mbar_expect_tx(bar(BAR_IN + b), in_bytes_layer);
                mbar_expect_tx(bar(BAR_Y + b), y_bytes);
                mbar_expect_tx(bar(BAR_SK + b), sk_bytes);
            } else if (is_head) {
                mbar_expect_tx(bar(BAR_IN + b), hin_bytes);
                mbar_expect_tx(bar(BAR_Y + b), hid_bytes);
                mbar_expect_tx(bar(BAR_SK + b), z_bytes);
            }
        }
    }

    long long* trace_row = nullptr;
    int trace_n = 0;
    auto stamp = [&]() { if (TRACE && trace_row && trace_n < TRACE_EV) trace_row[trace_n++] = clock64(); };

    if (is_layer) {
        // =====================================================================================================
        // layer role
        // =====================================================================================================
        const int l = slot;
        const bool first = l == 0, last_layer = l == L - 1;
        const bool has_res = P.layers[l].has_res != 0;
        const unsigned dil = (unsigned)P.layers[l].dilation;
        const unsigned dmask = (dil & (dil - 1u)) == 0u ? dil - 1u : 0xffffffffu;
        float2* ring = reinterpret_cast<float2*>(P.rings + P.layers[l].ring_off) + (size_t)h * NT + tid;
        auto ring_ptr = [&](long long t, int g) -> float2* {
            const unsigned sl = dmask != 0xffffffffu ? ((unsigned)t & dmask) : ((unsigned)t % dil);   // t < 2^32
            return ring + ((size_t)sl * G + g) * 2 * NT;
        };
        const bool up_local = !first && slot_cluster(slot - 1) == cluster;
        const int down_slot = last_layer ? P.head_slot : slot + 1;
        const bool down_local = slot_cluster(down_slot) == cluster;
        const bool need_y = has_res || has_skip;            // the pair itself consumes y
        const bool y_is_out = !has_res;                      // y is the next layer's input (or the head's, last layer w/o skips)
        const int n_down = last_layer ? P.NHC : 2;

        // remote windows
        const unsigned win_own = window((int)rank), win_peer = window((int)(rank ^ 1u));
        const unsigned win_up0 = up_local ? window(slot_rank0(slot - 1)) : 0u;
        const unsigned win_up1 = up_local ? window(slot_rank0(slot - 1) + 1) : 0u;
        const int drank0 = slot_rank0(down_slot);
        const unsigned win_dn0 = down_local ? window(drank0) : 0u;
        const unsigned win_dn1 = down_local ? window(drank0 + 1) : 0u;
        const unsigned win_dnh = h ? win_dn1 : win_dn0;      // same-half CTA of the next pair

        // ---- register-resident weights: newer conv tap (filter, gate rows of 4 channels) and 4 residual rows
        float4 wga[KJ], wgb[KJ], wr[KJ];
        {
            const float4* wb = reinterpret_cast<const float4*>(P.wreg) + ((size_t)(l * 2 + h) * 3 * KJ) * NT + tid;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                wga[j] = __ldg(wb + (size_t)(3 * j + 0) * NT);
                wgb[j] = __ldg(wb + (size_t)(3 * j + 1) * NT);
                wr[j] = __ldg(wb + (size_t)(3 * j + 2) * NT);
            }
        }
        // ---- shared-memory weights: older conv tap, skip rows; per-thread biases of the outputs it ends up holding
        const float* blk_g = P.wsm + (size_t)(l * 2 + h) * P.layer_block;
        {
            const float4* s4 = reinterpret_cast<const float4*>(blk_g);
            float4* d4 = reinterpret_cast<float4*>(smem);
            for (int i = tid; i < P.o_bias / 4; i += NT) d4[i] = __ldg(s4 + i);
        }
        const float b_gf = __ldg(blk_g + P.o_bias + tid), b_gg = __ldg(blk_g + P.o_bias + NT + tid);
        const float b_res = __ldg(blk_g + P.o_bias + 2 * NT + tid);
        float b_skip[2];
        b_skip[0] = __ldg(blk_g + P.o_bias + 3 * NT + tid);
        b_skip[1] = __ldg(blk_g + P.o_bias + 4 * NT + tid);
        const float4* Wold4 = reinterpret_cast<const float4*>(smem + P.o_wold);
        const float4* Wsk4 = reinterpret_cast<const float4*>(smem + P.o_wsk);
        __syncthreads();
        cluster_sync_all();

        // after a reduction lane16 holds (channel 4 q + (lane16 >> 2), prompt lane16 & 3)
        const int po = lane16 & 3;
        const int cl = 4 * q + (lane16 >> 2);           // channel / row inside this CTA's half
        const int ch = h * CH + cl;                      // global channel / residual row
        const unsigned out_off = (unsigned)(ch * GB + po) * 4u;
        unsigned n = 0, nh = 0;                          // units done; units done with the head on
        float2 a0 = __ldcg(ring_ptr(P.t_begin, 0));     // parked older-tap pre-activations (f, g) of the unit about to start
        float pf_e[2] = {0.0f, 0.0f};                    // first layer: prefetched embedding values of the next unit
        bool pf_ok = false;
        uint2 pf_x[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};   // mailbox-fed layer: the next unit's input words, fetched
        bool pf_x_ok = false;                                        // a unit ahead (an L2 round trip per unit otherwise)

        for (long long t = P.t_begin; t < P.t_end && !dead; ++t) {
            const unsigned delivery = (unsigned)(t - P.t_begin);
            const unsigned tag = delivery + 1u;
            const bool head_on = t >= P.t_head;
            const bool flow = P.teacher_forced || t <= P.t_head;   // nothing upstream paces the pipeline: honour the acks
            for (int g = 0; g < P.n_groups; ++g, ++n) {
                if (TRACE) {
                    trace_row = (P.trace && t == P.trace_t && h == 0 && tid == 0) ? P.trace + ((size_t)slot * G + g) * TRACE_EV : nullptr;
                    trace_n = 0;
                    if (trace_row) trace_row[trace_n++] = (long long)globaltimer();
                }
                stamp();
                const unsigned nb = n & 1u;
                int ng = g + 1;
                long long nt = t;
                if (ng == P.n_groups) { ng = 0; ++nt; }
                const bool sends_down = !last_layer || head_on;
                const unsigned cnt_down = last_layer ? nh : n;      // deliveries made to the downstream buffers so far
                const unsigned db = cnt_down & 1u;                   // downstream buffer of this unit
                const size_t box_dn = ((size_t)down_slot * G + g) * 2 + (delivery & 1u);

                // ---- credit: the downstream buffers [db] were last used two deliveries ago
                if (sends_down) {
                    if (down_local) {
                        if (cnt_down >= 2u) dead |= !mbar_wait(bar(BAR_FREE + db), ((cnt_down >> 1) - 1u) & 1u, abort_flag);
                    } else if (flow) {
                        const unsigned dlv = last_layer ? (unsigned)(t - P.t_head) : delivery;   // same (group, parity) box two steps ago
                        if (dlv >= 2u) {
                            const unsigned per_delivery = last_layer ? (unsigned)(P.NHC * (NW > GB ? NW - GB : NW)) : 2u * NW;   // consumer warps
                            if (lane == 0) dead |= !ack_wait(P.ack + (size_t)down_slot * G + g, (dlv - 1u) * per_delivery, abort_flag);
                            dead = __any_sync(0xffffffffu, dead);
                        }
                    }
                }

                // ---- first layer, free-running generation: peek at the next unit's sampled index now (the word and then the
                //      embedding row are two dependent L2 round trips; both fit under this unit's work when the sampler is ahead)
                uint2 peek = make_uint2(0u, 0u);
                const bool peeking = first && !P.teacher_forced && nt > P.t_head && nt < P.t_end;
                if (peeking && ng * GB + (tid & 3) < P.B)
                    peek = ld_poll_v2(reinterpret_cast<const uint2*>(P.samples + ng * GB + (tid & 3)));

                // ---- layer input -> inb[nb]
                float* xin = inb + nb * P.inblk;
                if (first) {
                    // embedding gather: x[k][p] = E[q_{b,t}][k]   (EmbeddingIO, modules/io.py:148-154); value i = (k = i >> 2, p = i & 3)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = tid + u * NT;
                        float e = pf_e[u];
                        if (!pf_ok) {
                            const int p = i & 3, k = i >> 2, b = g * GB + p;
                            e = 0.0f;
                            if (b < P.B) {
                                long long qi;
                                if (!P.teacher_forced && t > P.t_head) {   // produced by the sampler one step ago: poll the word
                                    uint2 v;
                                    dead |= !poll_word(reinterpret_cast<const uint2*>(P.samples + g * GB + p), tag, v, abort_flag);
                                    qi = (long long)v.x;
                                } else {
                                    qi = __ldcg(P.seq + (size_t)b * P.seq_stride + t);
                                }
                                qi = qi < 0 ? 0 : (qi >= P.Q ? P.Q - 1 : qi);
                                e = __ldg(P.E + (size_t)qi * C + k);
                            }
                        }
                        xin[i] = e;
                    }
                    pf_ok = false;
                    if (__syncthreads_or(dead ? 1 : 0)) { dead = true; break; }
                } else if (!up_local) {
                    const size_t box = ((size_t)slot * G + g) * 2 + (delivery & 1u);
                    const uint2* mx = P.mail_x + box * (size_t)P.inblk;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = tid + u * NT;
                        uint2 v = pf_x[u];
                        if (!pf_x_ok) v = ld_poll_v2(mx + i);
                        if (v.y != tag) dead |= !poll_word(mx + i, tag, v, abort_flag);
                        xin[i] = __uint_as_float(v.x);
                    }
                    pf_x_ok = false;
                    if (__syncthreads_or(dead ? 1 : 0)) { dead = true; break; }
                    if (nt < P.t_end) {   // the next unit's words: valid already when this stage is the slower one
                        const size_t boxn = ((size_t)slot * G + ng) * 2 + ((unsigned)(nt - P.t_begin) & 1u);
                        const uint2* mxn = P.mail_x + boxn * (size_t)P.inblk;
                        pf_x[0] = ld_poll_v2(mxn + tid);
                        pf_x[1] = ld_poll_v2(mxn + tid + NT);
                        pf_x_ok = true;
                    }
                } else {
                    dead |= !mbar_wait(bar(BAR_IN + nb), (n >> 1) & 1u, abort_flag);
                    if (tid == 0) mbar_expect_tx(bar(BAR_IN + nb), in_bytes_layer);   // arm the buffer's next use
                }
                stamp();

                // ---- newer tap on h_l(t) + the parked older tap -> gate -> y   (wavenet_v2.py:143-151)
                const float4* X4 = reinterpret_cast<const float4*>(xin);
                float y;
                {
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
#pragma unroll
                    for (int j = 0; j < KJ; ++j) fma_gate(acc, wga[j], wgb[j], X4[s + 16 * j]);
                    stamp();
                    const float2 fg = tree32x16(acc);
                    stamp();
                    y = gated(fg.x + a0.x, fg.y + a0.y);      // tanh(f) * sigmoid(g)
                    stamp();
                }
                {
                    if (need_y) {
                        const unsigned dst = sbase + (unsigned)(P.s_y + nb * P.yblk) * 4u + out_off;
                        st_async_f32(win_own + dst, y, win_own + bar(BAR_Y + nb));
                        st_async_f32(win_peer + dst, y, win_peer + bar(BAR_Y + nb));
                    }
                    if (y_is_out && sends_down && (!last_layer || !has_skip)) {   // h_{l+1} = y (no residual conv) or the head input (no skips)
                        if (down_local) {
                            const unsigned dst = sbase + (unsigned)(P.s_in + db * P.inblk) * 4u + out_off;
                            for (int dd = 0; dd < n_down; ++dd) {
                                const unsigned w = window(drank0 + dd);
                                st_async_f32(w + dst, y, w + bar(BAR_IN + db));
                            }
                        } else {
                            st_flagged_v2(P.mail_x + box_dn * (size_t)P.inblk + ch * GB + po, __float_as_uint(y), tag);
                        }
                    }
                }
                stamp();

                const float4* Y4 = reinterpret_cast<const float4*>(yb + nb * P.yblk);
                bool y_waited = false;
                // mailbox-fed layer: the incoming skip sums of this unit are asked for now and looked at in the skip phase
                uint2 pf_s[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
                if (has_skip && !first && !up_local) {
                    const size_t box = ((size_t)slot * G + g) * 2 + (delivery & 1u);
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass)
                        if (pass < P.skip_passes && (pass * CH + 8 * warp) < SH)
                            pf_s[pass] = ld_poll_v2(P.mail_s + box * (size_t)(S * GB) + (h * SH + pass * CH + cl) * GB + po);
                }
                // ---- residual 1x1 conv -> h_{l+1} = h_l + conv_res(y)   (wavenet_v2.py:172-175)
                if (has_res) {
                    dead |= !mbar_wait(bar(BAR_Y + nb), (n >> 1) & 1u, abort_flag);
                    if (tid == 0) mbar_expect_tx(bar(BAR_Y + nb), y_bytes);
                    y_waited = true;
                    stamp();
                    float acc[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
#pragma unroll
                    for (int j = 0; j < KJ; ++j) fma16(acc, wr[j], Y4[s + 16 * j]);
                    stamp();
                    float v = tree16x16(acc) + b_res;
                    v = xin[ch * GB + po] + v;
                    if (down_local) {
                        const unsigned dst = sbase + (unsigned)(P.s_in + db * P.inblk) * 4u + out_off;
                        st_async_f32(win_dn0 + dst, v, win_dn0 + bar(BAR_IN + db));
                        st_async_f32(win_dn1 + dst, v, win_dn1 + bar(BAR_IN + db));
                    } else {
                        st_flagged_v2(P.mail_x + box_dn * (size_t)P.inblk + ch * GB + po, __float_as_uint(v), tag);
                    }
                    stamp();
                }

                // ---- skip 1x1 conv: skips = conv_skip(y) + skips (wavenet_v2.py:165-171); the sum travels with the group
                auto skip_phase = [&]() {
                    if (!has_skip) return;
                    if (!y_waited) {
                        dead |= !mbar_wait(bar(BAR_Y + nb), (n >> 1) & 1u, abort_flag);
                        if (tid == 0) mbar_expect_tx(bar(BAR_Y + nb), y_bytes);
                        y_waited = true;
                    }
                    bool sk_waited = false;
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const int row = pass * CH + cl;                  // row inside this CTA's half of the skip rows
                        if (pass >= P.skip_passes || (pass * CH + 8 * warp) >= SH) continue;   // warp-uniform (SH is a multiple of 8)
                        float acc[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
                        const float4* Wp = Wsk4 + (size_t)pass * KJ * NT;
#pragma unroll
                        for (int j = 0; j < KJ; ++j) fma16(acc, Wp[j * NT + tid], Y4[s + 16 * j]);
                        float v = tree16x16(acc) + b_skip[pass];
                        const int grow = h * SH + row;                  // global skip row
                        if (!first) {
                            if (up_local) {
                                if (!sk_waited) {
                                    dead |= !mbar_wait(bar(BAR_SK + nb), (n >> 1) & 1u, abort_flag);
                                    if (tid == 0) mbar_expect_tx(bar(BAR_SK + nb), sk_bytes);
                                    sk_waited = true;
                                }
                                v = v + skb[nb * P.skblk + row * GB + po];
                            } else {
                                const size_t box = ((size_t)slot * G + g) * 2 + (delivery & 1u);
                                const uint2* ms = P.mail_s + box * (size_t)(S * GB) + grow * GB + po;
                                uint2 w = pf_s[pass];
                                if (w.y != tag) dead |= !poll_word(ms, tag, w, abort_flag);
                                v = v + __uint_as_float(w.x);
                            }
                        }
                        if (!last_layer) {
                            if (down_local)
                                st_async_f32(win_dnh + sbase + (unsigned)(P.s_sk + db * P.skblk + row * GB + po) * 4u, v, win_dnh + bar(BAR_SK + db));
                            else
                                st_flagged_v2(P.mail_s + box_dn * (size_t)(S * GB) + grow * GB + po, __float_as_uint(v), tag);
                        } else if (head_on) {                            // head input: every head CTA gets the full skip vector
                            if (down_local) {
                                const unsigned dst = sbase + (unsigned)(P.s_in + db * P.inblk + grow * GB + po) * 4u;
                                for (int dd = 0; dd < n_down; ++dd) {
                                    const unsigned w = window(drank0 + dd);
                                    st_async_f32(w + dst, v, w + bar(BAR_IN + db));
                                }
                            } else {
                                st_flagged_v2(P.mail_x + box_dn * (size_t)P.inblk + grow * GB + po, __float_as_uint(v), tag);
                            }
                        }
                    }
                };
                if (last_layer) skip_phase();          // the head waits for it: before the older tap

                // ---- older tap of this input, consumed at t + dilation: park it, bias included
                {
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
#pragma unroll
                    for (int j = 0; j < KJ; ++j) fma_gate(acc, Wold4[(2 * j + 0) * NT + tid], Wold4[(2 * j + 1) * NT + tid], X4[s + 16 * j]);
                    stamp();
                    float2 park = tree32x16(acc);
                    stamp();
                    park.x += b_gf; park.y += b_gg;
                    __stcg(ring_ptr(t, g), park);          // read back by this same thread at t + dilation
                    if (nt < P.t_end) a0 = __ldcg(ring_ptr(nt, ng));   // after the store: with one group and dilation 1 it is that word
                }
                if (peeking) {
                    // every thread of a warp looks at the same 4 words: the warp agrees on whether the sampler was ahead
                    const bool known = ng * GB + (tid & 3) >= P.B || peek.y == (unsigned)(nt - P.t_begin) + 1u;
                    if (__all_sync(0xffffffffu, known)) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = tid + u * NT, p = i & 3, k = i >> 2, b = ng * GB + p;
                            pf_e[u] = 0.0f;
                            if (b < P.B) {
                                long long qi = (long long)peek.x;
                                qi = qi >= P.Q ? P.Q - 1 : qi;
                                pf_e[u] = __ldg(P.E + (size_t)qi * C + k);
                            }
                        }
                        pf_ok = true;
                    }
                }
                if (first && nt < P.t_end && (P.teacher_forced || nt <= P.t_head)) {
                    // next unit's embedding values are known already (prompt / teacher forcing)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = tid + u * NT, p = i & 3, k = i >> 2, b = ng * GB + p;
                        pf_e[u] = 0.0f;
                        if (b < P.B) {
                            long long qi = __ldcg(P.seq + (size_t)b * P.seq_stride + nt);
                            qi = qi < 0 ? 0 : (qi >= P.Q ? P.Q - 1 : qi);
                            pf_e[u] = __ldg(P.E + (size_t)qi * C + k);
                        }
                    }
                    pf_ok = true;
                }
                stamp();
                if (!last_layer) skip_phase();
                stamp();

                // ---- return the input buffers of this unit to the producers upstream
                if (!first) {
                    __syncwarp();
                    if (up_local) {
                        if (lane == 0) mbar_arrive_remote(win_up0 + bar(BAR_FREE + nb));
                        if (lane == 1) mbar_arrive_remote(win_up1 + bar(BAR_FREE + nb));
                    } else if (lane == 0) {
                        red_add_u32(P.ack + (size_t)slot * G + g, 1u);
                    }
                }
                if (last_layer && h == 0 && tid == 0 && !head_on && g == P.n_groups - 1 && P.step_ts)
                    P.step_ts[t - P.t_begin] = globaltimer();   // steps before the head starts are stamped by the last layer
                if (head_on) ++nh;
                if (dead) break;
            }
        }
    } else if (is_head) {
        // =====================================================================================================
        // head role: hidden = mish(W1 x + b1) (mlp.py:44-53), z = W2 hidden + b2, learned temperature, sampler.
        // Rows in chunks of 4 per warp, K over the 32 lanes; the head CTAs take turns to gather the logits and decide.
        // =====================================================================================================
        const int NHC = P.NHC, Kh = P.Kh, Hh = P.Hh, Q = P.Q;
        {
            const float4* s4 = reinterpret_cast<const float4*>(P.hpack + (size_t)hr * P.head_block);
            float4* d4 = reinterpret_cast<float4*>(smem);
            for (int i = tid; i < P.head_block / 4; i += NT) d4[i] = __ldg(s4 + i);
        }
        const float4* W1 = reinterpret_cast<const float4*>(smem + P.o_w1);
        const float4* W2 = reinterpret_cast<const float4*>(smem + P.o_w2);
        const float* B1 = smem + P.o_b1;
        const float* B2 = smem + P.o_b2;
        __syncthreads();
        cluster_sync_all();

        const int up_slot = L - 1;
        const bool up_local = slot_cluster(up_slot) == cluster;
        const unsigned win_up0 = up_local ? window(slot_rank0(up_slot)) : 0u;
        const unsigned win_up1 = up_local ? window(slot_rank0(up_slot) + 1) : 0u;
        const int hrank0 = slot_rank0(P.head_slot);
        const int ro = (lane >> 3) & 3, po = (lane >> 1) & 3;   // after head_rows: row inside the chunk, prompt (lanes 2 i, 2 i + 1 agree)
        const int nc1 = P.nh1 / 4, nc2 = P.nzp / 4;
        const int my_rows = min(P.nz, Q - hr * P.nz);           // logit rows of this head CTA (may be <= 0 for a trailing CTA)
        // warps [0, GB) only decide (one prompt each, when it is this CTA's turn); the rows are shared by the other warps, so
        // that a decision (~1 900 cycles) never delays the hidden rows the other head CTAs are waiting for.  (With 4 warps
        // per CTA every warp does both.)
        constexpr bool SPLIT = NW > GB;
        constexpr int NRW = SPLIT ? NW - GB : NW;               // row warps
        const bool row_warp = !SPLIT || warp >= GB;
        const int rw = SPLIT ? warp - GB : warp, rtid = SPLIT ? tid - GB * 32 : tid;
        const bool row_leader = row_warp && rw == NRW - 1 && lane == 0;   // re-arms the row warps' barriers
        const long long t0 = P.t_begin > P.t_head ? P.t_begin : P.t_head;
        unsigned nh = 0;
        for (long long t = t0; t < P.t_end && !dead; ++t) {
            const unsigned delivery = (unsigned)(t - P.t_begin);
            const unsigned tag = delivery + 1u;
            for (int g = 0; g < P.n_groups; ++g, ++nh) {
                if (TRACE) {
                    trace_row = (P.trace && t == P.trace_t && rtid == 0) ? P.trace + ((size_t)(P.head_slot + hr) * G + g) * TRACE_EV : nullptr;
                    trace_n = 0;
                    if (trace_row) trace_row[trace_n++] = (long long)globaltimer();
                }
                stamp();
                const unsigned nb = nh & 1u;
                const int decider = (int)(nh % (unsigned)NHC);
                const unsigned zcnt = nh / (unsigned)NHC;             // units the decider has decided before this one
                const unsigned zb = zcnt & 1u;
                if (row_warp) {
                    float* hin = inb + nb * P.inblk;
                    if (up_local) {
                        dead |= !mbar_wait(bar(BAR_IN + nb), (nh >> 1) & 1u, abort_flag);
                        if (row_leader) mbar_expect_tx(bar(BAR_IN + nb), hin_bytes);
                    } else {
                        const size_t box = ((size_t)P.head_slot * G + g) * 2 + (delivery & 1u);
                        const uint2* mx = P.mail_x + box * (size_t)P.inblk;
                        for (int i = rtid; i < Kh * GB; i += NRW * 32) {
                            uint2 v = ld_poll_v2(mx + i);
                            if (v.y != tag) dead |= !poll_word(mx + i, tag, v, abort_flag);
                            hin[i] = __uint_as_float(v.x);
                        }
                        if (named_sync_or(2, NRW * 32, dead)) { dead = true; break; }
                    }
                    stamp();
                    // this CTA's logits buffer [zb] is refilled after this unit's hidden gather: the deciding warps must be done
                    // with its previous contents before any of this CTA's hidden rows leaves
                    if (hr == decider && zcnt >= 2u)
                        dead |= !mbar_wait(bar(BAR_ZDONE + zb), ((zcnt >> 1) - 1u) & 1u, abort_flag);
                    // hidden rows of this CTA, all-gathered over the head CTAs
                    auto hidden_out = [&](int ck, float v) {
                        const int row = ck * 4 + ro;
                        v = mish_acc(v + B1[row]);
                        const unsigned dst = sbase + (unsigned)(P.s_y + nb * P.yblk + (hr * P.nh1 + row) * GB + po) * 4u;
                        for (int dd = (lane & 1); dd < NHC; dd += 2) {
                            const unsigned w = window(hrank0 + dd);
                            st_async_f32(w + dst, v, w + bar(BAR_Y + nb));
                        }
                    };
                    {
                        const int cs1 = (Kh / 32) * 32;
                        int ck = rw;
                        for (; ck + NRW < nc1; ck += 2 * NRW) {
                            float v[2];
                            head_rows<2>(W1 + (size_t)ck * cs1, NRW * cs1, reinterpret_cast<const float4*>(hin), Kh, v);
                            hidden_out(ck, v[0]); hidden_out(ck + NRW, v[1]);
                        }
                        if (ck < nc1) {
                            float v[1];
                            head_rows<1>(W1 + (size_t)ck * cs1, 0, reinterpret_cast<const float4*>(hin), Kh, v);
                            hidden_out(ck, v[0]);
                        }
                    }
                    // the head input is consumed: return it to the last layer's CTAs
                    __syncwarp();
                    if (up_local) {
                        if (lane == 0) mbar_arrive_remote(win_up0 + bar(BAR_FREE + nb));
                        if (lane == 1) mbar_arrive_remote(win_up1 + bar(BAR_FREE + nb));
                    } else if (lane == 0) {
                        red_add_u32(P.ack + (size_t)P.head_slot * G + g, 1u);
                    }
                    dead |= !mbar_wait(bar(BAR_Y + nb), (nh >> 1) & 1u, abort_flag);
                    if (row_leader) mbar_expect_tx(bar(BAR_Y + nb), hid_bytes);
                    stamp();
                    // logit rows of this CTA (+ the learned-temperature row Q) -> the head CTA whose turn it is, [prompt][row]
                    const unsigned w = window(hrank0 + decider);
                    const float4* hid4 = reinterpret_cast<const float4*>(yb + nb * P.yblk);
                    auto logit_out = [&](int ck, float v) {
                        const int row = ck * 4 + ro;
                        if ((lane & 1) == 0 && row < my_rows)
                            st_async_f32(w + sbase + (unsigned)(P.s_sk + zb * P.skblk + po * P.zrow + hr * P.nz + row) * 4u, v + B2[row],
                                         w + bar(BAR_SK + zb));
                    };
                    {
                        const int cs2 = (Hh / 32) * 32;
                        int ck = rw;
                        for (; ck + 3 * NRW < nc2; ck += 4 * NRW) {
                            float v[4];
                            head_rows<4>(W2 + (size_t)ck * cs2, NRW * cs2, hid4, Hh, v);
                            logit_out(ck, v[0]); logit_out(ck + NRW, v[1]); logit_out(ck + 2 * NRW, v[2]); logit_out(ck + 3 * NRW, v[3]);
                        }
                        for (; ck + NRW < nc2; ck += 2 * NRW) {
                            float v[2];
                            head_rows<2>(W2 + (size_t)ck * cs2, NRW * cs2, hid4, Hh, v);
                            logit_out(ck, v[0]); logit_out(ck + NRW, v[1]);
                        }
                        if (ck < nc2) {
                            float v[1];
                            head_rows<1>(W2 + (size_t)ck * cs2, 0, hid4, Hh, v);
                            logit_out(ck, v[0]);
                        }
                    }
                    // the learned-temperature row (Q) is one more dot product per prompt: the decider's first row warp takes it
                    if (hr == decider && rw == 0) {
                        const float v = head_row1(smem + P.o_wt, hid4, Hh) + smem[P.o_wt + Hh];
                        if ((lane & 7) == 0)
                            st_async_f32(w + sbase + (unsigned)(P.s_sk + zb * P.skblk + (lane >> 3) * P.zrow + Q) * 4u, v, w + bar(BAR_SK + zb));
                    }
                    stamp();
                }
                if (hr == decider && warp < GB) {
                    dead |= !mbar_wait(bar(BAR_SK + zb), (zcnt >> 1) & 1u, abort_flag);
                    if (tid == 0) mbar_expect_tx(bar(BAR_SK + zb), z_bytes);
                    const int p = warp, b = g * GB + p;
                    if (b < P.B && !dead) {
                        const long long hstep = t - P.t_head, n_head = P.t_end - P.t_head;
                        float* lout = P.logits_out ? P.logits_out + ((size_t)b * n_head + hstep) * Q : nullptr;
                        const bool sample = P.temperature != nullptr;
                        float Tt = 1.0f, u = 0.0f;
                        if (sample) {
                            Tt = P.temperature[P.n_temperature == 1 ? 0 : b];
                            u = P.noise[(size_t)b * P.noise_stride + (t + 1 - P.noise_t0)];
                        }
                        const int choice = mmk::decide_warp(skb + zb * P.skblk + p * P.zrow, Q, P.min_temp, lout, sample, Tt, u);
                        if (lane == 0) {
                            if (P.decisions) P.decisions[(size_t)b * n_head + hstep] = choice;
                            if (!P.teacher_forced) {
                                st_flagged_v2(reinterpret_cast<uint2*>(P.samples + g * GB + p), (unsigned)choice, tag + 1u);
                                __stcg(P.seq + (size_t)b * P.seq_stride + t + 1, (long long)choice);
                            }
                        }
                    }
                    if (tid == 0 && g == P.n_groups - 1 && P.step_ts) P.step_ts[t - P.t_begin] = globaltimer();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar(BAR_ZDONE + zb)) : "memory");
                }
                if (dead) break;
            }
        }
    } else {
        __syncthreads();
        cluster_sync_all();
    }
    // no CTA may exit while peers can still store into its shared memory
    __syncthreads();
    cluster_sync_all();
}

static int pad4(int v) { return (v + 3) / 4 * 4; }

}  // namespace mmk6

using namespace mmk6;

struct wn6_handle {
    Params p{};
    int device = 0, max_batch = 0, rf = 0, nt = 0;
    size_t smem_bytes = 0;
    std::vector<void*> allocs;
    void* d_flags = nullptr;        // mailboxes + sample words + acks + abort flag: cleared before every launch
    size_t flags_bytes = 0;
    long long* d_trace = nullptr;   // debug timeline, only with MMK_WN_TRACE_T set
    long long trace_t = -1;
    int trace_rows = 0;
};

static const void* wn6_kernel(int C, bool trace) {
    switch (C) {
        case 64: return trace ? (const void*)wavenet6_kernel<64, true> : (const void*)wavenet6_kernel<64, false>;
        case 128: return trace ? (const void*)wavenet6_kernel<128, true> : (const void*)wavenet6_kernel<128, false>;
    }
    return nullptr;
}

// geometry + shared-memory carve-up for a cluster size and head width; returns the dynamic shared memory in bytes
static size_t wn6_plan(Params& p, int CS, int NHC) {
    const int C = p.C, S = p.S, NT = 2 * C, CH = C / 2, SH = S / 2;
    p.CS = CS; p.NPC = CS / 2; p.NHC = NHC;
    p.head_slot = p.L;
    if ((p.L % p.NPC) + NHC / 2 > p.NPC) p.head_slot = (p.L + p.NPC - 1) / p.NPC * p.NPC;
    p.NCL = (p.head_slot + NHC / 2 + p.NPC - 1) / p.NPC;
    p.skip_passes = S > 0 ? (SH + CH - 1) / CH : 0;
    int o = 0;
    auto take = [&](int floats) { int r = o; o += pad4(floats); return r; };
    p.o_wold = take(C * C);
    p.o_wsk = take(p.skip_passes * CH * C);
    p.o_bias = take(5 * NT);
    p.layer_block = o;
    o = 0;
    p.nh1 = p.Hh / NHC;
    p.nz = (p.Q + NHC - 1) / NHC;               // the learned-temperature row Q is handled apart
    p.nzp = pad4(p.nz);
    p.o_w1 = take(p.nh1 * p.Kh);
    p.o_w2 = take(p.nzp * p.Hh);
    p.o_b1 = take(p.nh1);
    p.o_b2 = take(p.nzp);
    p.o_wt = take(p.Hh + 1);
    p.head_block = o;
    o = std::max(p.o_bias, p.head_block);       // biases of a layer block go straight from global memory to registers
    p.zrow = pad4(p.Q + 1) + 4;
    if ((p.zrow % 32) == 0) p.zrow += 4;
    p.inblk = std::max(C, p.Kh) * GB;
    p.yblk = std::max(C, p.Hh) * GB;
    p.skblk = std::max(std::max(SH, 1) * GB, GB * p.zrow);
    p.s_in = take(2 * p.inblk);
    p.s_y = take(2 * p.yblk);
    p.s_sk = take(2 * p.skblk);
    p.s_bar = take(BAR_COUNT * 2);
    p.smem_floats = o;
    return (size_t)o * sizeof(float);
}

static int wn6_query_clusters(const void* k, int NT, int CS, size_t smem, int* out) {
    MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS * 4);
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

int wn6_destroy(wn6_handle* h) {
    if (!h) return 0;
    for (void* a : h->allocs) cudaFree(a);
    delete h;
    return 0;
}

int wn6_create(const mmk_wavenet_desc* d, int max_batch, wn6_handle** out, int* unsupported) {
    *unsupported = 1;
    const int C = d->dilated_dim, S = d->skips_dim;
    if (!(C == 64 || C == 128)) return 1;
    if (S % 16 || S > 2 * C || d->head_hidden % 32 || d->n_layers > MAX_LAYERS || d->n_layers < 1) return 1;
    const int Kh = S > 0 ? S : C;
    if (Kh % 32) return 1;
    // residual convs: on every layer but the last, or on none (wavenet_v2.py:78,216)
    bool any_res = false, all_res = true;
    for (int l = 0; l < d->n_layers - 1; ++l) { any_res |= d->conv_res_w[l] != nullptr; all_res &= d->conv_res_w[l] != nullptr; }
    if (any_res && !all_res) return 1;
    auto* h = new wn6_handle();
    Params& p = h->p;
    cudaGetDevice(&h->device);
    p.L = d->n_layers; p.C = C; p.S = S; p.Hh = d->head_hidden; p.Q = d->q_levels; p.Kh = Kh;
    p.min_temp = d->min_temperature;
    h->max_batch = max_batch;
    h->nt = 2 * C;
    p.G = (max_batch + GB - 1) / GB;
    int rf = 1;
    for (int l = 0; l < p.L; ++l) rf += d->dilations[l];
    h->rf = rf;
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);

    const void* kern = wn6_kernel(C, false);
    const char* force_cs = getenv("MMK_WN_CLUSTER");
    const char* force_hc = getenv("MMK_WN_HEAD_CTAS");
    int best_cs = 0, best_nhc = 0;
    for (int CS : {16, 8, 4, 2}) {
        if (force_cs && atoi(force_cs) != CS) continue;
        // head width: as many CTAs as still sit in the last layer's cluster (DSMEM hand-off), else the widest that fits
        const int npc = CS / 2, free_slots = (p.L % npc) ? npc - p.L % npc : 0;
        int nhc = 0;
        for (int pass = 0; pass < 2 && !nhc; ++pass)
            for (int cand : {8, 4, 2}) {
                if (force_hc && atoi(force_hc) != cand) continue;
                if (cand > CS || cand > MAX_HC || p.Hh % (cand * 4)) continue;
                if (pass == 0 && cand / 2 > free_slots) continue;
                Params q = p;
                if (wn6_plan(q, CS, cand) <= (size_t)max_optin) { nhc = cand; break; }
            }
        if (!nhc) continue;
        Params q = p;
        const size_t smem = wn6_plan(q, CS, nhc);
        if (!force_cs && CS > 2 && 2 * p.L + nhc <= CS / 2) continue;   // a smaller cluster holds everything
        int max_clusters = 0;
        if (wn6_query_clusters(kern, h->nt, CS, smem, &max_clusters)) { wn6_destroy(h); *unsupported = 0; return 1; }
        if (max_clusters < q.NCL) continue;
        best_cs = CS; best_nhc = nhc;
        break;
    }
    if (!best_cs) { wn6_destroy(h); return 1; }
    h->smem_bytes = wn6_plan(p, best_cs, best_nhc);
    *unsupported = 0;
    for (int tr = 0; tr < 2; ++tr) {
        const void* k = wn6_kernel(C, tr != 0);
        MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        if (best_cs > 8) MMK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }

    // ---- pack weights (thread-major: every load in the kernel is one coalesced float4 per thread)
    const int NT = 2 * C, KJ = C / KS, CH = C / 2, SH = S / 2;
    std::vector<float> wreg((size_t)p.L * 2 * 3 * KJ * NT * 4, 0.0f);
    std::vector<float> wsm((size_t)p.L * 2 * p.layer_block, 0.0f);
    long long ring_off = 0;
    for (int l = 0; l < p.L; ++l) {
        const bool has_res = d->conv_res_w[l] != nullptr;
        p.layers[l].dilation = d->dilations[l];
        p.layers[l].has_res = has_res ? 1 : 0;
        p.layers[l].ring_off = ring_off;
        ring_off += (long long)d->dilations[l] * p.G * 4 * NT;
        const float* wd = d->conv_dil_w[l];   // (2C, C, 2): [o][c][tap], tap 0 = older sample
        const float* bd = d->conv_dil_b[l];
        for (int hh = 0; hh < 2; ++hh) {
            float* rg = wreg.data() + ((size_t)(l * 2 + hh) * 3 * KJ) * NT * 4;
            float* sm = wsm.data() + (size_t)(l * 2 + hh) * p.layer_block;
            for (int tid = 0; tid < NT; ++tid) {
                const int s = tid & 15, q = tid >> 4;
                for (int j = 0; j < KJ; ++j) {
                    const int k = s + 16 * j;
                    for (int i = 0; i < 4; ++i) {          // slot i = channel 4 q + (i ^ (s >> 2)) of this half (pre-swapped for the folds)
                        const int cq = i ^ (s >> 2);
                        const int chn = hh * CH + 4 * q + cq;
                        for (int r = 0; r < 2; ++r) {      // filter row, gate row
                            const int o = (r ? C : 0) + chn;
                            const size_t unit = (size_t)(i / 2), el = (size_t)((i % 2) * 2 + r);
                            rg[(((size_t)(3 * j) + unit) * NT + tid) * 4 + el] = wd[((size_t)o * C + k) * 2 + 1];
                            sm[p.o_wold + (((size_t)(2 * j) + unit) * NT + tid) * 4 + el] = wd[((size_t)o * C + k) * 2 + 0];
                        }
                        rg[(((size_t)(3 * j) + 2) * NT + tid) * 4 + i] = has_res ? d->conv_res_w[l][(size_t)chn * C + k] : 0.0f;
                        for (int pass = 0; pass < p.skip_passes; ++pass) {
                            const int row = pass * CH + 4 * q + cq;
                            if (row < SH)
                                sm[p.o_wsk + (((size_t)pass * KJ + j) * NT + tid) * 4 + i] = d->conv_skip_w[l][(size_t)(hh * SH + row) * C + k];
                        }
                    }
                }
                // biases of the outputs this thread holds after the reductions: channel 4 q + (s >> 2)
                const int cl = 4 * q + (s >> 2), chn = hh * CH + cl;
                sm[p.o_bias + tid] = bd[chn];
                sm[p.o_bias + NT + tid] = bd[C + chn];
                sm[p.o_bias + 2 * NT + tid] = has_res ? d->conv_res_b[l][chn] : 0.0f;
                for (int pass = 0; pass < 2; ++pass) {
                    const int row = pass * CH + cl;
                    sm[p.o_bias + (3 + pass) * NT + tid] = (S > 0 && pass < p.skip_passes && row < SH) ? d->conv_skip_b[l][hh * SH + row] : 0.0f;
                }
            }
        }
    }
    std::vector<float> hpack((size_t)p.NHC * p.head_block, 0.0f);
    for (int r = 0; r < p.NHC; ++r) {
        float* hb = hpack.data() + (size_t)r * p.head_block;
        // chunks of 4 rows: float4 (chunk, j, lane) = the 4 rows at contraction index lane + 32 j
        for (int row = 0; row < p.nh1; ++row) {
            for (int k = 0; k < p.Kh; ++k)
                hb[p.o_w1 + ((((size_t)(row / 4) * (p.Kh / 32)) + k / 32) * 32 + k % 32) * 4 + ((row % 4) ^ (((k % 32) >> 3) & 3))] = d->head_w1[(size_t)(r * p.nh1 + row) * p.Kh + k];
            hb[p.o_b1 + row] = d->head_b1[r * p.nh1 + row];
        }
        for (int row = 0; row < p.nz; ++row) {
            const int grow = r * p.nz + row;
            if (grow >= p.Q) break;
            for (int k = 0; k < p.Hh; ++k)
                hb[p.o_w2 + ((((size_t)(row / 4) * (p.Hh / 32)) + k / 32) * 32 + k % 32) * 4 + ((row % 4) ^ (((k % 32) >> 3) & 3))] = d->head_w2[(size_t)grow * p.Hh + k];
            hb[p.o_b2 + row] = d->head_b2[grow];
        }
        for (int k = 0; k < p.Hh; ++k) hb[p.o_wt + k] = d->head_w2[(size_t)p.Q * p.Hh + k];
        hb[p.o_wt + p.Hh] = d->head_b2[p.Q];
    }
    bool ok = true;
    auto dev_alloc = [&](size_t bytes, const void* src) -> void* {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, std::max<size_t>(bytes, 16)) != cudaSuccess) { ok = false; return nullptr; }
        h->allocs.push_back(ptr);
        if (src) cudaMemcpy(ptr, src, bytes, cudaMemcpyHostToDevice); else cudaMemset(ptr, 0, std::max<size_t>(bytes, 16));
        return ptr;
    };
    p.wreg = (const float*)dev_alloc(wreg.size() * sizeof(float), wreg.data());
    p.wsm = (const float*)dev_alloc(wsm.size() * sizeof(float), wsm.data());
    p.hpack = (const float*)dev_alloc(hpack.size() * sizeof(float), hpack.data());
    p.E = (const float*)dev_alloc((size_t)p.Q * C * sizeof(float), d->embedding);
    p.rings = (float*)dev_alloc((size_t)ring_off * sizeof(float), nullptr);
    const int n_slots = p.head_slot + 1;
    const size_t boxes = (size_t)n_slots * p.G * 2;
    const size_t mx_bytes = boxes * p.inblk * sizeof(uint2);
    const size_t ms_bytes = boxes * std::max(1, S * GB) * sizeof(uint2);
    const size_t sw_bytes = (size_t)p.G * GB * sizeof(unsigned long long);
    const size_t ack_bytes = ((size_t)n_slots * p.G + 4) * sizeof(unsigned);
    h->flags_bytes = mx_bytes + ms_bytes + sw_bytes + ack_bytes;
    h->d_flags = dev_alloc(h->flags_bytes, nullptr);
    if (const char* e = getenv("MMK_WN_TRACE_T")) {
        h->trace_t = atoll(e);
        h->trace_rows = (p.head_slot + p.NHC) * p.G;
        h->d_trace = (long long*)dev_alloc((size_t)h->trace_rows * TRACE_EV * sizeof(long long), nullptr);
    }
    if (!ok) { wn6_destroy(h); MMK_FAIL("cudaMalloc failed while creating the WaveNet handle"); }
    char* f = (char*)h->d_flags;
    p.mail_x = (uint2*)f; f += mx_bytes;
    p.mail_s = (uint2*)f; f += ms_bytes;
    p.samples = (unsigned long long*)f; f += sw_bytes;
    p.ack = (unsigned*)f;
    p.abort_flag = p.ack + (size_t)n_slots * p.G;
    MMK_CUDA(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

int wn6_launch_info(wn6_handle* h, mmk_launch_info* out) {
    out->cluster_size = h->p.CS; out->n_stages = h->p.L + 1; out->group_size = GB; out->threads = h->nt;
    out->smem_bytes = (int)h->smem_bytes; out->sm_used = 2 * h->p.L + h->p.NHC;
    return 0;
}

int wn6_sync_check(wn6_handle* h, void* stream) {
    unsigned aborted = 0;
    MMK_CUDA(cudaMemcpyAsync(&aborted, h->p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MMK_CHECK(aborted == 0, "WaveNet kernel watchdog fired: an inter-stage wait timed out (results invalid)");
    if (h->d_trace) {   // debug: dump the timeline of step MMK_WN_TRACE_T as text (slot group globaltimer stamps...)
        const size_t n = (size_t)h->trace_rows * TRACE_EV;
        std::vector<long long> tr(n);
        MMK_CUDA(cudaMemcpy(tr.data(), h->d_trace, n * sizeof(long long), cudaMemcpyDeviceToHost));
        const char* path = getenv("MMK_WN_TRACE_FILE");
        if (FILE* f = fopen(path ? path : "wn_trace.txt", "w")) {
            for (int r = 0; r < h->trace_rows; ++r) {
                fprintf(f, "%d %d", r / h->p.G, r % h->p.G);
                for (int e = 0; e < TRACE_EV; ++e) fprintf(f, " %lld", tr[(size_t)r * TRACE_EV + e]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        MMK_CUDA(cudaMemset(h->d_trace, 0, n * sizeof(long long)));
    }
    return 0;
}

int wn6_run(wn6_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
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
    cfg.gridDim = dim3(p.CS * p.NCL);
    cfg.blockDim = dim3(h->nt);
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[] = {&p};
    MMK_CUDA(cudaLaunchKernelExC(&cfg, wn6_kernel(p.C, h->d_trace != nullptr), args));
    return 0;
}
