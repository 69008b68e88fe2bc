// Tensor-core WaveNet generation kernel (sm_100a: tcgen05.mma + TMEM + cp.async.bulk) — the bf16 mode of
// mmk_wavenet_* (compute_mode = 1).  Same function as wavenet3.cu / wavenet.cu (reference: wavenet_v2.py:131-176,
// 276-293, 447-452; modules/io.py:148-154; networks/mlp.py:44-63; modules/targets.py:40-52), operands rounded to bf16,
// fp32 accumulation.  Parity bar: teacher-forced logits within 5e-2 relative (BASELINE.json north_star, bf16).
//
// Design: the PROMPT BATCH is the MMA M dimension.  One CTA owns a group of up to 128 prompts for the whole launch
// (prefill + every generated sample) and never talks to another CTA: no collective, no grid or cluster barrier.
//   * every 1x1 / 2-tap contraction of a layer is a tcgen05.mma with M = 128 (prompts), fp32 accumulators in TMEM:
//       D1[128 x 2C]  = [h_l(t) | h_l(t-d)] . W1^T      (gate pre-activations, in 64-channel chunks: columns 128 j + [0, 64)
//                                                        hold the filter rows of chunk j, 128 j + 64 + [0, 64) its gate rows)
//       [H | SK][128 x (C + S)] += y . [Wres | Wskip]^T (ONE instruction of N = C + S per K-step: the residual stream H and
//                                                        the running skip sum SK stay in TMEM across all layers)
//     TMEM columns: D1 [0, 256) | H [256, 256 + C) | SK right after; the head reuses them (hidden -> D1, logits -> H | SK).
//   * weights (bf16, pre-packed on the host in the UMMA canonical K-major SWIZZLE_128B layout) are streamed from L2 through
//     a 4-slot x 32 KB shared-memory ring with cp.async.bulk + mbarrier complete_tx (two requests per slot) by a producer
//     warp; slots are released by tcgen05.commit.  Nothing is resident: 5.8 MB per step per group, all L2 hits.
//   * the tensor core fetches its shared-memory operands once per instruction, so narrow N multiplies the traffic and the
//     MMAs become fetch-bound (measured: N = 64 in a no-swizzle layout ran 3x over the tensor floor): instructions are
//     N = 128 (a gate chunk) or N = C + S, and the gated output y is handed to the res/skip MMAs as a TMEM A operand — the
//     epilogue packs it over filter columns of D1 it has just drained; y never touches shared memory.
//   * issue order (one elected lane of a warp whose control flow is warp-uniform, so descriptors sit in uniform registers):
//     newer-tap chunks of layer l; then per chunk j: res/skip atom j (waits for y_j, which also proves D1 half j drained)
//     followed by the OLDER-tap MMAs of layer l + 1, chunk j — they do not depend on this step's activations and run
//     under the epilogues.
//   * the older conv tap h_l(t-d) comes from a per-layer ring of d + 1 slots in global memory (L2): a producer warp copies
//     every layer-input tile shared -> global with one cp.async.bulk (the epilogue threads never store to global) and
//     fetches it d steps later with one bulk copy into the tap tile.
//   * 16 epilogue warps (4 threads per prompt row = TMEM lane) move TMEM -> registers: bias, tanh * sigmoid (2 MUFU.TANH per
//     gate), bf16 pack; the next layer's input goes straight into its swizzled A tile.  The residual-conv biases never
//     enter the stream (folded into the gate biases on the host); the skip sum is zeroed by the epilogue at the start of
//     a step and only ever accumulated.
//   * head: skip sum -> bf16 -> MMA (W1) -> Mish -> MMA (W2, + learned-temperature row) -> logits staged in shared
//     memory -> the same decide_warp sampler as the fp32 kernels (argmax or inverse-CDF on external noise).
//   * every wait is an mbarrier wait with a watchdog (2 s): a lost signal aborts the launch instead of hanging the GPU.
// Work per step per group: 30 x (16 + 8) K-steps of 128 x 256 x 16: ~3 070 tensor cycles per layer (the floor of this
// design); algorithmic FLOPs 2 x 2 982 016 per sample per prompt.  MMK_TC_TRACE_T=<t> dumps an in-kernel timeline of step t.
#include "common.cuh"
#include "sampler.cuh"
#include "wavenet_impl.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mmk_tc {

constexpr int NEW = 16;            // epilogue warps: 4 threads per prompt row
constexpr int QT = NEW / 4;        // column quarters
constexpr int NT = 32 * (NEW + 3); // + MMA warp, weight producer, tap producer
constexpr int NEPI = 32 * NEW;
constexpr int W_MMA = NEW, W_WP = NEW + 1, W_XP = NEW + 2;
constexpr int MROWS = 128;         // prompts per group = MMA M
constexpr int NSLOT = 4;
constexpr int SLOT_BYTES = 32768;      // one 64-deep K atom of up to 256 rows (loaded as two 16 KB bulk copies)
constexpr int MAXL = 96;
constexpr int TM_D1 = 0, TM_H = 256, TM_TEMP = 128, TM_LOGIT = 256;   // the skip sum sits right after H: TM_H + C

// shared-memory map (bytes)
constexpr int SM_XN = 0;                         // A tile: h_l(t) bf16 [128 x C]; head: hidden
constexpr int SM_Y = 32768;                      // A tile: gated output y; head: skip sum
// 65536 .. 67584: 2 KB that, with XN|Y, make the logits staging area (64 rows x 260 floats)
constexpr int SM_XO = 67584;                     // A tile: h_l(t-d) from the ring
constexpr int SM_W = 100352;                     // weight ring
constexpr int SM_BAR = SM_W + NSLOT * SLOT_BYTES;
constexpr int SM_TOTAL = SM_BAR + 1024;
constexpr int ZROW = 260;

enum {
    B_WFULL = 0, B_WEMPTY = NSLOT, B_XOFULL = 2 * NSLOT, B_XOFREE, B_XNFULL, B_XNSAVED, B_D1FULL, B_YFULL = B_D1FULL + 2,
    B_D2FULL = B_YFULL + 2, B_HAFULL, B_H1DFULL, B_H2AFULL, B_H2DFULL, B_HEADDONE, B_EPISYNC, B_COUNT
};

constexpr int MAX_HS = 12;                    // weight stages of the head

struct Params {
    int L, C, S, Hh, Q, n_h2;
    // weight stages (<= 32 KB each).  Layer l's block of wpack (layer_bytes each) holds, per 64-channel chunk j < n_ch:
    //   older-tap tile  [f rows | g rows of chunk j] x C   at  j * cb              (cb bytes)
    //   newer-tap tile                                      at  (n_ch + j) * cb
    //   res|skip K atom [C res rows | S skip rows] x 64     at  2 n_ch cb + j * wb  (wb bytes)
    // streamed per step as: older(0) ; then per layer l: newer(l), { res|skip(l, j), older(l + 1, j) } for every j.
    // The head's n_hs stages start at head_off + hs_off[i].
    int n_ch, n_hs;
    unsigned cb, wb, hs_off[MAX_HS], hs_bytes[MAX_HS];
    unsigned long long layer_bytes, head_off;
    float min_temp;
    int dil[MAXL];
    unsigned char has_res[MAXL];
    long long ring_off[MAXL];      // byte offset of layer l's ring inside a group's block
    long long ring_group_bytes;
    const unsigned char* wpack;
    const float* E;                // (Q, C)
    const float* b1;               // [L][2C] gate biases (filter rows then gate rows)
    const float* cbs;              // [S] sum of the skip-conv biases
    const float* hb1; const float* hb2;
    unsigned char* rings;
    unsigned* abort_flag;
    // this run
    long long* seq;
    long long seq_stride, t_begin, t_head, t_end;
    int B, teacher_forced, n_temperature;
    const float* temperature; const float* noise;
    long long noise_stride, noise_t0;
    float* logits_out; long long* decisions; unsigned long long* step_ts;
    int exp;                               // timing experiments (bit 0: no weight copies, bit 1: no tap copies; results invalid): always 0 in a build
    long long* trace; long long trace_t;   // MMK_TC_TRACE_T: clock64 stamps of group 0 at that step, [role][layer][16]
};
constexpr int TRACE_EV = 16;

// ------------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long WAIT_LIMIT_NS = 2000000000ull;   // watchdog: 2 s on one wait means a lost signal
__device__ __noinline__ bool mbar_wait_slow(unsigned bar, unsigned parity, unsigned* abort_flag) {
    const unsigned long long t0 = globaltimer();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0u) {
            if (*reinterpret_cast<volatile unsigned*>(abort_flag) != 0u) return false;
            if (globaltimer() - t0 > WAIT_LIMIT_NS) { atomicExch(abort_flag, 1u); return false; }
        }
    }
    return true;
}
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity, unsigned* abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    return mbar_wait_slow(bar, parity, abort_flag);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// generic-proxy st.shared -> async-proxy reads (tcgen05.mma operands): the cheap CTA-local form
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned dst_smem, unsigned cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 operands, fp32 accumulate, M = 128, K = 16
__device__ __forceinline__ void umma_bf16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                          unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in TMEM: row i in lane i, 16 bf16 K-elements packed two per 32-bit column (8 columns)
__device__ __forceinline__ void umma_bf16_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc, unsigned idesc,
                                             unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
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
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
                   "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                   "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor, version 1, layout_type 2): a tile of
// R rows x K bf16 is a sequence of K/64 "atoms"; an atom holds 64 K-elements of every row as 128-byte lines, 8 rows =
// 1024 bytes (SBO), line r of a group stores its 16-byte chunk c at position c ^ (r & 7).  Atoms are R * 128 bytes apart.
// A K-step of 16 elements advances the start address by 32 bytes inside an atom.  Tile bases are 1024-byte aligned.
__host__ __device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3ffffu) >> 4);
    d |= (unsigned long long)1u << 16;                 // LBO: unused for swizzled K-major layouts (CUTLASS writes 1)
    d |= (unsigned long long)(1024u >> 4) << 32;       // SBO
    d |= 1ull << 46;                                   // descriptor version (sm_100)
    d |= 2ull << 61;                                   // SWIZZLE_128B
    return d;
}
// start-address advance (16-byte units) of K-step kk in a tile of R rows
__host__ __device__ __forceinline__ unsigned kstep16(int kk, int R) { return (unsigned)((kk >> 2) * (R * 8) + (kk & 3) * 2); }
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128
__host__ __device__ __forceinline__ unsigned umma_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(MROWS >> 4) << 24);
}
// byte offset of (row m, 8-element chunk kc) in a swizzled tile of R rows
__host__ __device__ __forceinline__ unsigned tile_off(int m, int kc, int R) {
    return (unsigned)((kc >> 3) * (R * 128) + (m >> 3) * 1024 + (m & 7) * 128 + (((kc & 7) ^ (m & 7)) << 4));
}
__device__ __forceinline__ unsigned pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<unsigned*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ float tanh_fast(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// tanh(f) * sigmoid(g), sigmoid(g) = 0.5 tanh(g / 2) + 0.5: two MUFU.TANH and four FP32 ops per gate.  (A packed
// tanh.approx.f16x2 does not help: it is issued as two MUFU.TANH.F16, plus the conversions.)
__device__ __forceinline__ float gate_fast(float f, float g) {
    return tanh_fast(f) * fmaf(0.5f, tanh_fast(0.5f * g), 0.5f);
}
__device__ __forceinline__ float mish_fast(float x) {   // x * tanh(softplus(x)), softplus threshold 20 (F.softplus)
    const float sp = x > 20.0f ? x : __logf(1.0f + __expf(x));
    return x * tanh_fast(sp);
}

__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------
// The kernel: one CTA per group of 128 prompts.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) wavenet_tc_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = blockIdx.x;
    const unsigned sb = smem_u32(smem);
    auto bar = [&](int i) { return sb + (unsigned)SM_BAR + 8u * (unsigned)i; };
    unsigned* s_tmem = reinterpret_cast<unsigned*>(smem + SM_BAR + 8 * B_COUNT);
    int* s_idx = reinterpret_cast<int*>(smem + SM_BAR + 8 * B_COUNT + 16);
    unsigned* abort_flag = P.abort_flag;
    const int L = P.L, C = P.C, S = P.S, Hh = P.Hh, Q = P.Q;
    const unsigned tile_bytes = (unsigned)(MROWS * C * 2);
    const unsigned TM_SK = TM_H + C;

    if (tid == 0) {
        if (sb & 1023u) atomicExch(abort_flag, 2u);       // the swizzled tiles need a 1024-byte aligned window
        for (int i = 0; i < NSLOT; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
        mbar_init(bar(B_XOFULL), 1); mbar_init(bar(B_XOFREE), 1);
        mbar_init(bar(B_XNFULL), NEPI); mbar_init(bar(B_XNSAVED), 1);
        for (int j = 0; j < 2; ++j) { mbar_init(bar(B_D1FULL + j), 1); mbar_init(bar(B_YFULL + j), NEPI); }
        mbar_init(bar(B_D2FULL), 1);
        mbar_init(bar(B_HAFULL), NEPI); mbar_init(bar(B_H1DFULL), 1);
        mbar_init(bar(B_H2AFULL), NEPI); mbar_init(bar(B_H2DFULL), 1);
        mbar_init(bar(B_HEADDONE), NEPI);
        mbar_init(bar(B_EPISYNC), NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<volatile unsigned*>(s_tmem);

    unsigned char* ring_g = P.rings + (size_t)grp * P.ring_group_bytes;

    if (warp == W_WP) {
        // ===================== weight producer: the stage list of every step, in consumption order =====================
        if (lane == 0) {
            unsigned cnt = 0;
            bool dead = false;
            auto load = [&](const unsigned char* src, unsigned bytes) -> bool {
                const unsigned slot = cnt % NSLOT, use = cnt / NSLOT;
                if (!mbar_wait(bar(B_WEMPTY + slot), (use & 1u) ^ 1u, abort_flag)) return false;
                if ((P.exp & 1) && cnt >= NSLOT) { mbar_arrive(bar(B_WFULL + slot)); ++cnt; return true; }
                mbar_expect_tx(bar(B_WFULL + slot), bytes);
                const unsigned dst = sb + SM_W + slot * SLOT_BYTES, half = bytes > 16384u ? 16384u : bytes;
                bulk_g2s(dst, src, half, bar(B_WFULL + slot));            // two requests in flight per slot
                if (bytes > half) bulk_g2s(dst + half, src + half, bytes - half, bar(B_WFULL + slot));
                ++cnt;
                return true;
            };
            for (long long t = P.t_begin; t < P.t_end && !dead; ++t) {
                const int n_ch = P.n_ch;
                for (int j = 0; j < n_ch && !dead; ++j)                       // older tap of layer 0
                    if (!load(P.wpack + (size_t)j * P.cb, P.cb)) dead = true;
                for (int l = 0; l < L && !dead; ++l) {
                    const unsigned char* wl = P.wpack + (size_t)l * P.layer_bytes;
                    for (int j = 0; j < n_ch && !dead; ++j)                   // newer tap
                        if (!load(wl + (size_t)(n_ch + j) * P.cb, P.cb)) dead = true;
                    for (int j = 0; j < n_ch && !dead; ++j) {                 // res|skip atom j, then the next layer's older tap
                        if (!load(wl + (size_t)2 * n_ch * P.cb + (size_t)j * P.wb, P.wb)) { dead = true; break; }
                        if (l + 1 < L && !load(wl + P.layer_bytes + (size_t)j * P.cb, P.cb)) dead = true;
                    }
                }
                if (t >= P.t_head)
                    for (int i = 0; i < P.n_hs && !dead; ++i)
                        if (!load(P.wpack + P.head_off + P.hs_off[i], P.hs_bytes[i])) dead = true;
            }
        }
    } else if (warp == W_XP) {
        // ===================== tap producer: ring traffic of the older conv tap =====================
        // per layer: fetch h_l(t - d_l) into the tap tile; then copy the layer's input tile h_l(t) to its ring slot.
        // The ring of layer l has d_l + 1 slots, so the slot written at step t is never the one read at step t.
        if (lane == 0) {
            unsigned n = 0;
            bool dead = false;
            for (long long t = P.t_begin; t < P.t_end && !dead; ++t) {
                for (int l = 0; l < L; ++l, ++n) {
                    if (!mbar_wait(bar(B_XOFREE), (n & 1u) ^ 1u, abort_flag)) { dead = true; break; }
                    if (l == 0) bulk_wait_all();          // every ring store of the previous steps has landed
                    const unsigned slots = (unsigned)P.dil[l] + 1u;
                    if ((P.exp & 2) && n >= 1) {
                        mbar_arrive(bar(B_XOFULL));
                    } else {
                        mbar_expect_tx(bar(B_XOFULL), tile_bytes);
                        bulk_g2s(sb + SM_XO, ring_g + P.ring_off[l] + (size_t)((t + 1) % slots) * tile_bytes, tile_bytes,
                                 bar(B_XOFULL));
                    }
                    if (!mbar_wait(bar(B_XNFULL), n & 1u, abort_flag)) { dead = true; break; }
                    bulk_s2g(ring_g + P.ring_off[l] + (size_t)(t % slots) * tile_bytes, sb + SM_XN, tile_bytes);
                    bulk_commit();
                    bulk_wait_read();                     // the tile has been read: the epilogue may overwrite it
                    mbar_arrive(bar(B_XNSAVED));
                }
            }
            bulk_wait_all();
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        // The whole warp runs the control flow, so that every descriptor is a warp-uniform value (uniform registers,
        // no per-instruction R2UR waterfall); one elected lane issues the tcgen05.mma / tcgen05.commit instructions.
        const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        auto wait_u = [&](unsigned b, unsigned parity) -> bool {    // warp-uniform result
            return __all_sync(0xffffffffu, mbar_wait(b, parity, abort_flag) ? 1 : 0) != 0;
        };
        auto elect = [&]() -> bool {
            unsigned pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            return pred != 0;
        };
        unsigned wcnt = 0, n_lay = 0, n_head = 0;
        bool dead = false;
        long long* tr = nullptr;
        int tn = 0;
        auto stamp = [&]() { if (tr && lane == 0 && tn < TRACE_EV) tr[tn] = clock64(); ++tn; };
        const unsigned id128 = umma_idesc(128), idRS = umma_idesc(C + S), idS = umma_idesc(S);
        const int n_ch = P.n_ch, KC = C / 16;
        const unsigned long long dXO = umma_desc(sb + SM_XO), dXN = umma_desc(sb + SM_XN), dY = umma_desc(sb + SM_Y);
        const unsigned long long dW0 = umma_desc(sb + SM_W);
        constexpr unsigned SLOT16 = SLOT_BYTES >> 4;
        auto wslot_wait = [&](unsigned& slot) -> bool {
            slot = wcnt % NSLOT;
            const bool ok = wait_u(bar(B_WFULL + slot), (wcnt / NSLOT) & 1u);
            tc_fence_after();
            return ok;
        };
        // older tap of global layer n (chunk j): D1[:, 128 j ..] = XO . W1o_j^T.  Independent of this step's activations,
        // it is issued early — right after the previous layer's res/skip MMAs of the same chunk, whose wait on the gated
        // output also guarantees that the gate epilogue has drained that half of D1.
        auto issue_old = [&](unsigned n, int j) -> bool {
            if (j == 0) {
                if (!wait_u(bar(B_XOFULL), n & 1u)) return false;
                tc_fence_after();
            }
            unsigned slot;
            if (!wslot_wait(slot)) return false;
            const unsigned long long dW = dW0 + (unsigned long long)(slot * SLOT16);
            if (elect()) {
                for (int kk = 0; kk < KC; ++kk)
                    umma_bf16(tmem_u + TM_D1 + 128 * j, dXO + kstep16(kk, MROWS), dW + kstep16(kk, 128), id128, kk > 0);
                umma_commit(bar(B_WEMPTY + slot));
                if (j == n_ch - 1) umma_commit(bar(B_XOFREE));
            }
            __syncwarp();
            ++wcnt;
            return true;
        };
        for (long long t = P.t_begin; t < P.t_end && !dead; ++t) {
            for (int j = 0; j < n_ch && !dead; ++j) dead = !issue_old(n_lay, j);           // layer 0 of this step
            for (int l = 0; l < L && !dead; ++l, ++n_lay) {
                tr = (P.trace && grp == 0 && t == P.trace_t) ? P.trace + (size_t)l * TRACE_EV : nullptr;
                tn = 0;
                stamp();                                                    // 0: layer start
                // ---- newer tap: D1 += XN . W1n^T, 64-channel chunk by chunk (the gate epilogue follows chunk by chunk)
                if (!wait_u(bar(B_XNFULL), n_lay & 1u)) { dead = true; break; }
                tc_fence_after();
                stamp();                                                    // 1: layer input tile ready
                for (int j = 0; j < n_ch && !dead; ++j, ++wcnt) {
                    unsigned slot;
                    if (!wslot_wait(slot)) { dead = true; break; }
                    const unsigned long long dW = dW0 + (unsigned long long)(slot * SLOT16);
                    if (elect()) {
                        for (int kk = 0; kk < KC; ++kk)
                            umma_bf16(tmem_u + TM_D1 + 128 * j, dXN + kstep16(kk, MROWS), dW + kstep16(kk, 128), id128, 1u);
                        umma_commit(bar(B_WEMPTY + slot));
                        umma_commit(bar(B_D1FULL + j));
                    }
                    __syncwarp();
                }
                if (dead) break;
                stamp();                                                    // 2: newer-tap MMAs issued
                // ---- residual and skip 1x1 convs on y: ONE instruction of N = C + S per K-step (H and the skip sum are
                //      adjacent in TMEM), a 64-channel K atom at a time as the gate epilogue delivers them; each is
                //      followed by the NEXT layer's older-tap MMAs of the same chunk (they run under the gate epilogue
                //      of the other chunk and under the next-input epilogue)
                const bool has_res = P.has_res[l] != 0;
                for (int j = 0; j < n_ch && !dead; ++j) {
                    if (!wait_u(bar(B_YFULL + j), n_lay & 1u)) { dead = true; break; }
                    tc_fence_after();
                    stamp();                                                // 3, 5: y chunk j ready
                    unsigned slot;
                    if (!wslot_wait(slot)) { dead = true; break; }
                    const unsigned long long dW = dW0 + (unsigned long long)(slot * SLOT16);
                    if (elect()) {
                        for (int kq = 0; kq < 4; ++kq) {
                            // A = y (chunk j, channels 16 kq ..) straight from TMEM: the gate epilogue wrote it, packed, over
                            // the first 8 of the 16 filter columns it had just read (D1 column 128 j + 16 kq)
                            const unsigned a = tmem_u + TM_D1 + 128 * j + 16 * kq;
                            if (has_res) umma_bf16_ts(tmem_u + TM_H, a, dW + 2u * kq, idRS, 1u);
                            else umma_bf16_ts(tmem_u + TM_SK, a, dW + (unsigned long long)(C * 8) + 2u * kq, idS, 1u);
                        }
                        umma_commit(bar(B_WEMPTY + slot));
                        if (j == n_ch - 1) umma_commit(bar(B_D2FULL));
                    }
                    __syncwarp();
                    ++wcnt;
                    if (l + 1 < L && !issue_old(n_lay + 1u, j)) { dead = true; break; }
                    stamp();                                                // 4, 6: next layer's older tap issued
                }
                if (dead) break;
            }
            if (dead) break;
            if (t >= P.t_head) {
                // ---- head: hidden = A(skip sum) . W1^T ; logits = A2(mish hidden) . W2^T
                if (!wait_u(bar(B_HAFULL), n_head & 1u)) break;
                tc_fence_after();
                for (int a = 0; a < S / 64 && !dead; ++a, ++wcnt) {
                    unsigned slot;
                    if (!wslot_wait(slot)) { dead = true; break; }
                    const unsigned long long dW = dW0 + (unsigned long long)(slot * SLOT16);
                    if (elect()) {
                        for (int kq = 0; kq < 4; ++kq)
                            umma_bf16(tmem_u + TM_D1, dY + kstep16(4 * a + kq, MROWS), dW + 2u * kq, umma_idesc(Hh), (a | kq) != 0);
                        umma_commit(bar(B_WEMPTY + slot));
                        if (a == S / 64 - 1) umma_commit(bar(B_H1DFULL));
                    }
                    __syncwarp();
                }
                if (dead) break;
                if (!wait_u(bar(B_H2AFULL), n_head & 1u)) break;
                tc_fence_after();
                for (int ca = 0; ca < P.n_h2 * (Hh / 64) && !dead; ++ca, ++wcnt) {   // row chunk c, K atom a
                    unsigned slot;
                    if (!wslot_wait(slot)) { dead = true; break; }
                    const unsigned long long dW = dW0 + (unsigned long long)(slot * SLOT16);
                    const int c = ca / (Hh / 64), a = ca - c * (Hh / 64);
                    const bool temp_chunk = (c == P.n_h2 - 1);        // the learned-temperature row, padded to 16
                    const int rows = temp_chunk ? 16 : min(128, Q - 128 * c);
                    const unsigned dcol = temp_chunk ? TM_TEMP : TM_LOGIT + 128 * c;
                    if (elect()) {
                        for (int kq = 0; kq < 4; ++kq)
                            umma_bf16(tmem_u + dcol, dXN + kstep16(4 * a + kq, MROWS), dW + 2u * kq, umma_idesc(rows), (a | kq) != 0);
                        umma_commit(bar(B_WEMPTY + slot));
                        if (ca == P.n_h2 * (Hh / 64) - 1) umma_commit(bar(B_H2DFULL));
                    }
                    __syncwarp();
                }
                if (dead) break;
                // D1 / H / SK are rewritten by the next step: wait until the epilogue has drained the logits
                if (!wait_u(bar(B_HEADDONE), n_head & 1u)) break;
                ++n_head;
            }
        }
    } else if (warp < NEW) {
        // ===================== epilogue warps: QT threads per prompt row, each owns 1/QT of every column range =====================
        const int q4 = warp & 3, hf = warp >> 2;         // TMEM lane quarter (fixed by the warp id), column part
        const int m = 32 * q4 + lane;                    // prompt row = TMEM lane
        const int b = grp * MROWS + m;
        const bool live = b < P.B;
        const unsigned tm_lane = tmem + ((unsigned)(32 * q4) << 16);
        unsigned n_lay = 0, n_head = 0, n_sync = 0, n_saved = 0;
        bool dead = false;
        long long* tr = nullptr;
        int tn = 0;
        auto stamp = [&]() { if (tr && tn < TRACE_EV) tr[tn++] = clock64(); };
        auto epi_sync = [&]() {      // barrier of the 256 epilogue threads (an mbarrier: the wait has the watchdog)
            mbar_arrive(bar(B_EPISYNC));
            dead |= !mbar_wait(bar(B_EPISYNC), n_sync & 1u, abort_flag);
            ++n_sync;
        };
        auto wait_saved = [&]() {    // the tap producer has copied the current layer-input tile out: it may be rewritten
            dead |= !mbar_wait(bar(B_XNSAVED), n_saved & 1u, abort_flag);
            ++n_saved;
        };
        const int half_c = C / QT, half_s = S / QT, half_h = Hh / QT, half_q = Q / QT;   // per-thread widths
        float* zs = reinterpret_cast<float*>(smem + SM_XN);

        for (long long t = P.t_begin; t < P.t_end; ++t) {
            // ---------------- embedding: h_0(t) = E[q_t]  (EmbeddingIO, modules/io.py:148-154) ----------------
            {
                long long q = 0;
                if (live) q = (!P.teacher_forced && t > P.t_head) ? (long long)s_idx[m] : __ldcg(P.seq + (size_t)b * P.seq_stride + t);
                q = q < 0 ? 0 : (q >= Q ? Q - 1 : q);
                const float* row = P.E + (size_t)q * C + hf * half_c;
                for (int i = 0; i < half_c / 16; ++i) {
                    float v[16];
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 e = __ldg(reinterpret_cast<const float4*>(row + 16 * i) + k4);
                        v[4 * k4] = e.x; v[4 * k4 + 1] = e.y; v[4 * k4 + 2] = e.z; v[4 * k4 + 3] = e.w;
                    }
                    const int c0 = hf * half_c + 16 * i;
                    tmem_st16(tm_lane + TM_H + c0, v);
                    *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8, MROWS)) = pack8(v);
                    *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8 + 1, MROWS)) = pack8(v + 8);
                }
                {
                    float z[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) z[k] = 0.0f;
                    for (int i = 0; i < half_s / 16; ++i) tmem_st16(tm_lane + TM_SK + hf * half_s + 16 * i, z);   // skip sum = 0
                }
                tmem_st_wait();
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(bar(B_XNFULL));
            }
            // ---------------- layers ----------------
            for (int l = 0; l < L; ++l, ++n_lay) {
                const float* bl = P.b1 + (size_t)l * 2 * C;
                tr = (P.trace && grp == 0 && tid == 0 && t == P.trace_t) ? P.trace + (size_t)(MAXL + l) * TRACE_EV : nullptr;
                tn = 0;
                stamp();                                                        // 0: layer start
                for (int j = 0; j < P.n_ch; ++j) {
                    // gate epilogue of the 64-channel chunk j: this thread's 16 channels 64 j + 16 hf + [0, 16);
                    // D1 columns of chunk j: filter rows at 128 j + [0, 64), gate rows at 128 j + 64 + [0, 64)
                    const int ch0 = 64 * j + 16 * hf;
                    float4 bf[4], bg[4];                     // biases, fetched before the wait
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        bf[k4] = __ldg(reinterpret_cast<const float4*>(bl + ch0) + k4);
                        bg[k4] = __ldg(reinterpret_cast<const float4*>(bl + C + ch0) + k4);
                    }
                    dead |= !mbar_wait(bar(B_D1FULL + j), n_lay & 1u, abort_flag);
                    tc_fence_after();
                    stamp();                                                    // gate pre-activations of chunk j complete
                    float f[16], g[16];
                    tmem_ld16(tm_lane + TM_D1 + 128 * j + 16 * hf, f);
                    tmem_ld16(tm_lane + TM_D1 + 128 * j + 64 + 16 * hf, g);
                    tmem_ld_wait();
                    float y[16];
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        y[4 * k4] = gate_fast(f[4 * k4] + bf[k4].x, g[4 * k4] + bg[k4].x);     // wavenet_v2.py:151
                        y[4 * k4 + 1] = gate_fast(f[4 * k4 + 1] + bf[k4].y, g[4 * k4 + 1] + bg[k4].y);
                        y[4 * k4 + 2] = gate_fast(f[4 * k4 + 2] + bf[k4].z, g[4 * k4 + 2] + bg[k4].z);
                        y[4 * k4 + 3] = gate_fast(f[4 * k4 + 3] + bf[k4].w, g[4 * k4 + 3] + bg[k4].w);
                    }
                    {   // y (bf16, two channels per 32-bit column) into TMEM, over this thread's own drained filter columns
                        unsigned yp[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) yp[k] = pack_bf16(y[2 * k], y[2 * k + 1]);
                        tmem_st8(tm_lane + TM_D1 + 128 * j + 16 * hf, yp);
                        tmem_st_wait();
                    }
                    tc_fence_before();
                    mbar_arrive(bar(B_YFULL + j));
                    stamp();                                                    // y chunk j written
                }
                dead |= !mbar_wait(bar(B_D2FULL), n_lay & 1u, abort_flag);
                tc_fence_after();
                stamp();                                                        // res/skip MMAs complete
                wait_saved();                                                   // this layer's input tile is in its ring slot
                if (dead) break;
                if (l < L - 1) {
                    // h_{l+1}(t) = h_l + conv_res(y) -> next layer's A tile.  The residual-conv biases are not in the
                    // stream: their effect on the next gates is folded into the gate biases (host).
                    for (int i = 0; i < half_c / 16; i += 2) {
                        const int c0 = hf * half_c + 16 * i;
                        const bool two = i + 1 < half_c / 16;
                        float v[16], w[16];
                        tmem_ld16(tm_lane + TM_H + c0, v);
                        if (two) tmem_ld16(tm_lane + TM_H + c0 + 16, w);
                        tmem_ld_wait();
                        *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8, MROWS)) = pack8(v);
                        *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8 + 1, MROWS)) = pack8(v + 8);
                        if (two) {
                            *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8 + 2, MROWS)) = pack8(w);
                            *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8 + 3, MROWS)) = pack8(w + 8);
                        }
                    }
                    fence_proxy_async_smem();
                    tc_fence_before();
                    mbar_arrive(bar(B_XNFULL));
                    stamp();                                                    // next layer input written
                }
            }
            if (dead) break;
            // ---------------- head + sampler ----------------
            if (t >= P.t_head) {
                for (int i = 0; i < half_s / 16; ++i) {          // skip sum -> bf16 A tile (in the y buffer)
                    const int c0 = hf * half_s + 16 * i;
                    float v[16];
                    tmem_ld16(tm_lane + TM_SK + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] += __ldg(P.cbs + c0 + k);
                    *reinterpret_cast<uint4*>(smem + SM_Y + tile_off(m, c0 / 8, MROWS)) = pack8(v);
                    *reinterpret_cast<uint4*>(smem + SM_Y + tile_off(m, c0 / 8 + 1, MROWS)) = pack8(v + 8);
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(bar(B_HAFULL));
                dead |= !mbar_wait(bar(B_H1DFULL), n_head & 1u, abort_flag);
                tc_fence_after();
                for (int i = 0; i < half_h / 16; ++i) {          // hidden = mish(. + b1) -> bf16 A tile (mlp.py:44-53)
                    const int c0 = hf * half_h + 16 * i;
                    float v[16];
                    tmem_ld16(tm_lane + TM_D1 + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = mish_fast(v[k] + __ldg(P.hb1 + c0 + k));
                    *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8, MROWS)) = pack8(v);
                    *reinterpret_cast<uint4*>(smem + SM_XN + tile_off(m, c0 / 8 + 1, MROWS)) = pack8(v + 8);
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(bar(B_H2AFULL));
                dead |= !mbar_wait(bar(B_H2DFULL), n_head & 1u, abort_flag);
                tc_fence_after();
                if (dead) break;
                // logits: two passes of 64 rows through the staging area, one warp decides one row at a time
                for (int pass = 0; pass < 2; ++pass) {
                    if ((q4 >> 1) == pass) {
                        float* zr = zs + (size_t)(m - 64 * pass) * ZROW;
                        for (int i = 0; i < half_q / 16; ++i) {
                            const int c0 = hf * half_q + 16 * i;
                            float v[16];
                            tmem_ld16(tm_lane + TM_LOGIT + c0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(P.hb2 + c0) + k4);
                                *reinterpret_cast<float4*>(zr + c0 + 4 * k4) =
                                    make_float4(v[4 * k4] + bb.x, v[4 * k4 + 1] + bb.y, v[4 * k4 + 2] + bb.z, v[4 * k4 + 3] + bb.w);
                            }
                        }
                        if (hf == 0) {
                            float v[16];
                            tmem_ld16(tm_lane + TM_TEMP, v);
                            tmem_ld_wait();
                            zr[Q] = v[0] + __ldg(P.hb2 + Q);
                        }
                    }
                    if (pass == 1) { tc_fence_before(); mbar_arrive(bar(B_HEADDONE)); }   // every TMEM read of the step is done
                    epi_sync();
                    for (int r = 0; r < 64 / NEW; ++r) {
                        const int mr = 64 * pass + (64 / NEW) * warp + r, br = grp * MROWS + mr;
                        if (br < P.B) {
                            const long long hstep = t - P.t_head, n_head_steps = P.t_end - P.t_head;
                            float* lout = P.logits_out ? P.logits_out + ((size_t)br * n_head_steps + hstep) * Q : nullptr;
                            const bool sample = P.temperature != nullptr;
                            float Tt = 1.0f, u = 0.0f;
                            if (sample) {
                                Tt = P.temperature[P.n_temperature == 1 ? 0 : br];
                                u = P.noise[(size_t)br * P.noise_stride + (t + 1 - P.noise_t0)];
                            }
                            const int choice = mmk::decide_warp(zs + (size_t)((64 / NEW) * warp + r) * ZROW, Q, P.min_temp, lout, sample, Tt, u);
                            if (lane == 0) {
                                s_idx[mr] = choice;
                                if (P.decisions) P.decisions[(size_t)br * n_head_steps + hstep] = choice;
                                if (!P.teacher_forced) __stcg(P.seq + (size_t)br * P.seq_stride + t + 1, (long long)choice);
                            }
                        }
                    }
                    epi_sync();
                }
                ++n_head;
                if (dead) break;
            }
            if (grp == 0 && tid == 0 && P.step_ts) P.step_ts[t - P.t_begin] = globaltimer();
        }
    }
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------
// Self-test kernel: D[128 x N] = A[128 x K] . B[N x K]^T through the same descriptors / TMEM path (K a multiple of 64).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tc_gemm_check_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                               float* __restrict__ D, int N, int K, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned sb = smem_u32(smem);
    const int a_bytes = MROWS * K * 2, b_bytes = ((N * K * 2 + 1023) / 1024) * 1024;
    unsigned char* sA = smem;                       // 128 x K bf16
    unsigned char* sB = smem + a_bytes;             // N x K bf16
    unsigned* s_misc = reinterpret_cast<unsigned*>(sB + b_bytes);
    const unsigned bar0 = smem_u32(s_misc), s_tm = smem_u32(s_misc + 4);
    for (int i = tid; i < MROWS * K / 8; i += 128) {
        const int m = i / (K / 8), kc = i % (K / 8);
        float v[8];
        for (int e = 0; e < 8; ++e) v[e] = A[(size_t)m * K + kc * 8 + e];
        *reinterpret_cast<uint4*>(sA + tile_off(m, kc, MROWS)) = pack8(v);
    }
    for (int i = tid; i < N * K / 8; i += 128) {
        const int n = i / (K / 8), kc = i % (K / 8);
        float v[8];
        for (int e = 0; e < 8; ++e) v[e] = Bm[(size_t)n * K + kc * 8 + e];
        *reinterpret_cast<uint4*>(sB + tile_off(n, kc, N)) = pack8(v);
    }
    if (tid == 0) { mbar_init(bar0, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(s_tm, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<volatile unsigned*>(s_misc + 4);
    long long c0 = 0;
    if (tid == 0) {
        c0 = clock64();
        for (int rep = 0; rep < 8; ++rep)           // the same product 8 times (timing); the last pass is the result
            for (int kk = 0; kk < K / 16; ++kk)
                umma_bf16(tmem, umma_desc(sb) + kstep16(kk, MROWS), umma_desc(sb + a_bytes) + kstep16(kk, N), umma_idesc(N), kk > 0);
        umma_commit(bar0);
    }
    unsigned spins = 0;
    while (!mbar_try_wait(bar0, 0u)) { if (++spins > 100000000u) break; }
    if (tid == 0 && cycles) *cycles = clock64() - c0;
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(tmem + ((unsigned)(32 * warp) << 16) + c, v);
        tmem_ld_wait();
        for (int k = 0; k < 16; ++k) D[(size_t)(32 * warp + lane) * N + c + k] = v[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace mmk_tc

using namespace mmk_tc;

struct wn4_handle {
    Params p{};
    std::vector<void*> allocs;
    int device = 0, max_groups = 0, smem_bytes = 0;
    long long* d_trace = nullptr; long long trace_t = 0;
};

int wn4_destroy(wn4_handle* h) {
    if (!h) return 0;
    for (void* a : h->allocs) cudaFree(a);
    delete h;
    return 0;
}

static unsigned short f2bf(float f) {   // round to nearest even
    unsigned u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}

// Writes rows [row_at, row_at + n) of a canonical bf16 tile of R rows x K from a strided fp32 matrix
// (src[r * ld + k * es]); rows >= valid are zero.
static void pack_rows(std::vector<unsigned char>& out, size_t at, int R, int K, int row_at, int n, const float* src,
                      size_t ld, size_t es, int valid) {
    for (int r = 0; r < n; ++r)
        for (int k = 0; k < K; ++k) {
            const unsigned short bfv = f2bf((src && r < valid) ? src[(size_t)r * ld + (size_t)k * es] : 0.0f);
            memcpy(out.data() + at + tile_off(row_at + r, k / 8, R) + (k % 8) * 2, &bfv, 2);
        }
}

int wn4_create(const mmk_wavenet_desc* d, int max_batch, wn4_handle** out, int* unsupported) {
    *unsupported = 1;
    const int L = d->n_layers, C = d->dilated_dim, S = d->skips_dim, Hh = d->head_hidden, Q = d->q_levels;
    const char* why = nullptr;
    if (L < 1 || L > MAXL) why = "n_layers out of range";
    else if (C != 64 && C != 128) why = "dilated_dim must be 64 or 128";
    else if (S != 64 && S != 128) why = "skips_dim must be 64 or 128";
    else if (Hh != 64 && Hh != 128) why = "head hidden_dim must be 64 or 128";
    else if (Q % 64 != 0 || Q < 64 || Q > 256) why = "q_levels must be 64, 128, 192 or 256";
    if (!why)
        for (int l = 0; l < L; ++l) {
            if ((d->conv_res_w[l] != nullptr) != (l < L - 1)) why = "every layer but the last needs a residual conv";
            if (!d->conv_skip_w || !d->conv_skip_w[l]) why = "every layer needs a skip conv";
            if (d->dilations[l] < 1) why = "dilation must be >= 1";
        }
    if (why) {
        mmk::set_error(std::string("bf16 tensor-core WaveNet kernel: unsupported configuration: ") + why);
        return 1;
    }
    *unsupported = 0;
    auto* h = new wn4_handle();
    Params& p = h->p;
    MMK_CUDA(cudaGetDevice(&h->device));
    int cc_major = 0;
    MMK_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, h->device));
    if (cc_major != 10) { delete h; MMK_FAIL("the bf16 tensor-core WaveNet kernel needs an sm_100 device (tcgen05)"); }
    p.L = L; p.C = C; p.S = S; p.Hh = Hh; p.Q = Q;
    p.n_h2 = (Q + 127) / 128 + 1;
    p.min_temp = d->min_temperature;
    h->max_groups = (max_batch + MROWS - 1) / MROWS;
    const size_t tile_bytes = (size_t)MROWS * C * 2;
    long long ring = 0;
    for (int l = 0; l < L; ++l) {
        p.dil[l] = d->dilations[l];
        p.has_res[l] = d->conv_res_w[l] ? 1 : 0;
        p.ring_off[l] = ring;
        ring += (long long)(d->dilations[l] + 1) * (long long)tile_bytes;   // d + 1 slots: the write of step t never
                                                                            // lands on the slot the tap producer reads
    }
    p.ring_group_bytes = ring;

    // ---- packed bf16 weights + stage tables
    std::vector<unsigned char> wpack;
    const int n_ch = C / 64;
    p.n_ch = n_ch; p.cb = (unsigned)(128 * C * 2); p.wb = (unsigned)((C + S) * 128);
    p.layer_bytes = (unsigned long long)2 * n_ch * p.cb + (unsigned long long)n_ch * p.wb;
    wpack.assign((size_t)L * p.layer_bytes, 0);
    for (int l = 0; l < L; ++l) {
        const float* wd = d->conv_dil_w[l];   // (2C, C, 2): [o][c][tap], tap 0 = older sample; rows o < C filter, o >= C gate
        const size_t base = (size_t)l * p.layer_bytes;
        for (int tap = 0; tap < 2; ++tap)      // block order: older-tap chunk tiles, then newer-tap chunk tiles
            for (int j = 0; j < n_ch; ++j) {
                // tile of 128 rows x C: rows 0..63 filter channels 64 j + r ; rows 64..127 gate channels 64 j + r
                const size_t at = base + (size_t)(tap * n_ch + j) * p.cb;
                pack_rows(wpack, at, 128, C, 0, 64, wd + ((size_t)(64 * j) * C) * 2 + tap, (size_t)C * 2, 2, 64);
                pack_rows(wpack, at, 128, C, 64, 64, wd + ((size_t)(C + 64 * j) * C) * 2 + tap, (size_t)C * 2, 2, 64);
            }
        for (int j = 0; j < n_ch; ++j) {       // K atom j (64 gated channels) of [residual rows | skip rows]
            const size_t at = base + (size_t)2 * n_ch * p.cb + (size_t)j * p.wb;
            pack_rows(wpack, at, C + S, 64, 0, C, d->conv_res_w[l] ? d->conv_res_w[l] + 64 * j : nullptr, (size_t)C, 1, C);
            pack_rows(wpack, at, C + S, 64, C, S, d->conv_skip_w[l] + 64 * j, (size_t)C, 1, S);
        }
    }
    p.head_off = wpack.size();
    auto add_hs = [&](size_t bytes) {
        const size_t at = wpack.size();
        p.hs_off[p.n_hs] = (unsigned)(at - p.head_off); p.hs_bytes[p.n_hs] = (unsigned)bytes; ++p.n_hs;
        wpack.resize(at + ((bytes + 1023) / 1024) * 1024, 0);
        return at;
    };
    for (int a = 0; a < S / 64; ++a) {
        const size_t at = add_hs((size_t)Hh * 128);
        pack_rows(wpack, at, Hh, 64, 0, Hh, d->head_w1 + 64 * a, (size_t)S, 1, Hh);
    }
    for (int c = 0; c < p.n_h2; ++c) {
        const bool temp_chunk = c == p.n_h2 - 1;
        const int rows = temp_chunk ? 16 : std::min(128, Q - 128 * c), row0 = temp_chunk ? Q : 128 * c;
        for (int a = 0; a < Hh / 64; ++a) {
            const size_t at = add_hs((size_t)rows * 128);
            pack_rows(wpack, at, rows, 64, 0, rows, d->head_w2 + (size_t)row0 * Hh + 64 * a, (size_t)Hh, 1, temp_chunk ? 1 : rows);
        }
    }
    if (p.n_hs > MAX_HS) { delete h; MMK_FAIL("internal: stage table overflow"); }

    std::vector<float> b1((size_t)L * 2 * C), cbr((size_t)(L + 1) * C, 0.0f), cbs(S, 0.0f);
    for (int l = 0; l < L; ++l) {
        // The residual stream is kept WITHOUT the residual-conv biases (TMEM accumulates only the MMAs); layer l's input
        // is therefore short of cbr_l = sum_{j<l} b_res_j, a constant: both taps of W1_l see it, so W1_l . cbr_l (fp64,
        // fp32 weights) is added to the gate biases instead.
        for (int i = 0; i < 2 * C; ++i) {
            double acc = d->conv_dil_b[l][i];
            for (int c = 0; c < C; ++c)
                acc += ((double)d->conv_dil_w[l][((size_t)i * C + c) * 2] + (double)d->conv_dil_w[l][((size_t)i * C + c) * 2 + 1]) *
                       (double)cbr[(size_t)l * C + c];
            b1[(size_t)l * 2 * C + i] = (float)acc;
        }
        for (int i = 0; i < C; ++i)
            cbr[(size_t)(l + 1) * C + i] = cbr[(size_t)l * C + i] + (d->conv_res_w[l] ? d->conv_res_b[l][i] : 0.0f);
        for (int i = 0; i < S; ++i) cbs[i] += d->conv_skip_b[l][i];
    }
    bool ok = true;
    auto dev_alloc = [&](size_t bytes, const void* src) -> void* {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, bytes) != cudaSuccess) { ok = false; return nullptr; }
        h->allocs.push_back(ptr);
        if (src) cudaMemcpy(ptr, src, bytes, cudaMemcpyHostToDevice); else cudaMemset(ptr, 0, bytes);
        return ptr;
    };
    p.wpack = (const unsigned char*)dev_alloc(wpack.size(), wpack.data());
    p.E = (const float*)dev_alloc((size_t)Q * C * 4, d->embedding);
    p.b1 = (const float*)dev_alloc(b1.size() * 4, b1.data());
    p.cbs = (const float*)dev_alloc(cbs.size() * 4, cbs.data());
    p.hb1 = (const float*)dev_alloc((size_t)Hh * 4, d->head_b1);
    p.hb2 = (const float*)dev_alloc((size_t)(Q + 1) * 4, d->head_b2);
    p.rings = (unsigned char*)dev_alloc((size_t)h->max_groups * (size_t)ring, nullptr);
    p.abort_flag = (unsigned*)dev_alloc(16, nullptr);
    if (const char* e = getenv("MMK_TC_TRACE_T")) {   // debug timeline of one step (see wn4_sync_check)
        h->trace_t = atoll(e);
        h->d_trace = (long long*)dev_alloc((size_t)2 * MAXL * TRACE_EV * sizeof(long long), nullptr);
    }
    if (!ok) { wn4_destroy(h); MMK_FAIL("cudaMalloc failed while creating the bf16 WaveNet handle"); }
    h->smem_bytes = SM_TOTAL;
    MMK_CUDA(cudaFuncSetAttribute(wavenet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    MMK_CUDA(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

int wn4_launch_info(wn4_handle* h, mmk_launch_info* out) {
    out->cluster_size = 1; out->n_stages = 1; out->group_size = MROWS; out->threads = NT;
    out->smem_bytes = h->smem_bytes; out->sm_used = h->max_groups;
    return 0;
}

int wn4_sync_check(wn4_handle* h, void* stream) {
    unsigned aborted = 0;
    MMK_CUDA(cudaMemcpyAsync(&aborted, h->p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MMK_CHECK(aborted == 0, "bf16 WaveNet kernel watchdog fired: a pipeline wait timed out (results invalid)");
    if (h->d_trace) {   // debug: dump the stamps of step MMK_TC_TRACE_T (role 0 = MMA issuer, 1 = epilogue thread 0)
        std::vector<long long> tr((size_t)2 * MAXL * TRACE_EV);
        MMK_CUDA(cudaMemcpy(tr.data(), h->d_trace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const char* path = getenv("MMK_TC_TRACE_FILE");
        if (FILE* f = fopen(path ? path : "tc_trace.txt", "w")) {
            for (int r = 0; r < 2; ++r)
                for (int l = 0; l < h->p.L; ++l) {
                    fprintf(f, "%d %d", r, l);
                    for (int e = 0; e < TRACE_EV; ++e) fprintf(f, " %lld", tr[((size_t)r * MAXL + l) * TRACE_EV + e]);
                    fprintf(f, "\n");
                }
            fclose(f);
        }
    }
    return 0;
}

int wn4_run(wn4_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
            int64_t t_end, int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream) {
    Params p = h->p;
    p.seq = reinterpret_cast<long long*>(d_seq) - seq_t0;
    p.seq_stride = seq_stride; p.t_begin = t_begin; p.t_head = t_head; p.t_end = t_end;
    p.B = B; p.teacher_forced = teacher_forced ? 1 : 0;
    p.temperature = d_temperature; p.n_temperature = n_temperature;
    p.noise = d_noise; p.noise_stride = noise_stride; p.noise_t0 = noise_t0;
    p.logits_out = d_logits_out; p.decisions = reinterpret_cast<long long*>(d_decisions); p.step_ts = d_step_ts;
    p.trace = h->d_trace; p.trace_t = h->trace_t;
    const int groups = (B + MROWS - 1) / MROWS;
    MMK_CHECK(groups <= h->max_groups, "batch exceeds the max_batch the handle was created for");
    MMK_CUDA(cudaMemsetAsync(p.abort_flag, 0, sizeof(unsigned), (cudaStream_t)stream));
    wavenet_tc_kernel<<<groups, NT, h->smem_bytes, (cudaStream_t)stream>>>(p);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

// Diagnostic: D (128 x N) = A (128 x K) . B (N x K)^T with bf16 operands through the UMMA path used above.
// h_cycles (nullable, host): clock cycles of 8 back-to-back passes of the K / 16 instructions (synchronises the stream).
extern "C" int mmk_tc_gemm_check(const float* d_A, const float* d_B, float* d_D, int N, int K, long long* h_cycles, void* stream) {
    MMK_CHECK(d_A && d_B && d_D, "mmk_tc_gemm_check: null pointer");
    MMK_CHECK(N % 16 == 0 && N >= 16 && N <= 256 && K % 64 == 0 && K >= 64 && K <= 256, "need N a multiple of 16 in [16, 256], K of 64 in [64, 256]");
    const size_t smem = (size_t)MROWS * K * 2 + ((size_t)N * K * 2 + 1023) / 1024 * 1024 + 64;
    MMK_CUDA(cudaFuncSetAttribute(tc_gemm_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* d_cycles = nullptr;
    if (h_cycles) MMK_CUDA(cudaMalloc(&d_cycles, sizeof(long long)));
    tc_gemm_check_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d_A, d_B, d_D, N, K, d_cycles);
    MMK_CUDA(cudaGetLastError());
    if (h_cycles) {
        MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        MMK_CUDA(cudaMemcpy(h_cycles, d_cycles, sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(d_cycles);
    }
    return 0;
}
