// Layer-pipelined tensor-core WaveNet generation kernel (sm_100a: tcgen05.mma + TMEM + DSMEM) — the bf16 mode
// (compute_mode = MMK_COMPUTE_BF16_TC) for W-30-shaped networks (128 dilated / skip / head channels).
// Function computed: as wavenet6.cu (wavenet_v2.py:131-176, 276-293, 447-452; modules/io.py:148-154; networks/mlp.py:44-63;
// modules/targets.py:40-52), operands rounded to bf16, fp32 accumulation — the arithmetic of oracle.restate.WaveNetBf16Oracle.
//
// The first tensor-core kernel here (wavenet_tc.cu) made the prompt batch the MMA M dimension: one CTA per 128 prompts,
// all 5.8 MB of weights re-streamed from L2 every step, ONE SM busy at the batch sizes BASELINE.json names.  This kernel
// turns the problem round:
//   * the WEIGHTS are the M = 128 operand and stay resident: CTA l owns layer l for the whole launch with its six
//     128 x 128 bf16 tiles (filter / gate rows of both conv taps, residual rows, skip rows: 192 KB) in shared memory in the
//     UMMA canonical K-major SWIZZLE_128B layout.  Nothing is streamed, nothing is reloaded.
//   * a group of 16 prompts is the N dimension: D[128 channels x 16 prompts] (fp32, TMEM) = W[128 x 128] . X[16 x 128]^T,
//     8 K-steps per tile.  Groups flow down the pipeline of L layer CTAs + the head CTA (one or two 16-CTA clusters); the
//     step time is the dependency chain, not the batch: B = 64 .. 128 prompts per GPU is 4 .. 8 groups in a 31-deep
//     pipeline.
//   * thread c of the four epilogue warps is channel c = TMEM lane c: tcgen05.ld gives it the 16 prompts of its row, so
//     tanh * sigmoid, the residual add (the fp32 residual stream of the group sits in its registers) and the skip sum are
//     thread-local; the gated output goes back to shared memory as the next MMA's bf16 B tile (16 two-byte stores).
//   * hand-off to the next layer: the fp32 residual stream and the fp32 running skip sum (8 KB each) as st.async.v4 into
//     the next CTA's shared memory, completing on its mbarriers; credits come back as remote mbarrier arrivals.  Across
//     the cluster boundary the same blocks go through L2 and are pulled in by the consumer's copy thread with
//     cp.async.bulk behind a release / acquire counter.
//   * the older conv tap x(t - d) is a 4 KB bf16 tile in an L2 ring, stored and fetched with bulk copies by the copy
//     thread; its MMAs are issued a unit ahead (they do not depend on this step), so only the newer tap is on the chain.
//   * one elected lane of the MMA warp issues every tcgen05.mma from warp-uniform control flow; all waits are mbarrier
//     waits with a watchdog.
#include "common.cuh"
#include "sampler.cuh"
#include "wavenet_impl.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mmk7 {

constexpr int NP = 16;             // prompts per group = MMA N
constexpr int CC = 128;            // channels = MMA M (dilated = skips = head hidden)
constexpr int NT = 320;            // layers: warps 0-7 epilogue (thread = channel x half of the group's prompts), 8 MMA issuer, 9 copy thread
                                   // head: warps 0-3 epilogue (thread = channel), 8 MMA issuer, 9 mailbox pull, every warp decides
constexpr int NEPI_H = 128;        // epilogue threads of the head
constexpr int NEPI_L = 256;        // ... of a layer
constexpr int EW_H = 4, EW_L = 8;  // epilogue warps (arrivals per credit / ack / mailbox flag)
constexpr int NPH = NP / 2;        // prompts per layer epilogue thread
constexpr int W_MMA = 8, W_CP = 9;
constexpr int MAXL = 96;
constexpr int TILE_A = CC * CC * 2;        // 32 KB: a weight tile
constexpr int TILE_B = NP * CC * 2;        // 4 KB: an activation tile (16 prompts x 128 channels, bf16)
constexpr int TILE_F = CC * NP * 4;        // 8 KB: fp32 [channel][16 prompts]
constexpr int ZROW = 260;
constexpr int TRACE_EV = 16;

// shared-memory map (bytes)
constexpr int SM_W = 0;                                   // layer: 6 weight tiles | head: W1, W2 x 3
constexpr int SM_XB = 6 * TILE_A;                         // bf16 B tile of the layer input x(t)   | head: skip sum
constexpr int SM_XO = SM_XB + TILE_B;                     // bf16 B tile of x(t - d) from the ring
constexpr int SM_YB = SM_XO + TILE_B;                     // bf16 B tile of the gated output       | head: hidden
constexpr int SM_XF = SM_YB + TILE_B;                     // fp32 residual stream from upstream
constexpr int SM_SK = SM_XF + TILE_F;                     // fp32 running skip sum from upstream   | head: its input
constexpr int SM_BAR = SM_SK + TILE_F;
constexpr int SM_TOTAL = SM_BAR + 512;
constexpr int SM_Z = 4 * TILE_A;                          // head: logits [16][ZROW] fp32 (16.6 KB), after its four weight tiles
static_assert(SM_TOTAL <= 232448, "shared memory budget");
static_assert(SM_Z + NP * ZROW * 4 <= SM_XB, "head logits staging overlaps the tiles");

enum { W_NF = 0, W_NG, W_OF, W_OG, W_RES, W_SKIP };      // layer tile order
enum {
    B_XF = 0,        // fp32 x arrived (tx)                       | head: -
    B_SK,            // fp32 skip sum arrived (tx)                | head: its input
    B_XB_FULL,       // bf16 x tile complete: sent by the layer above (bulk copy, tx) or built here (128)
    B_XB_FREE,       // built here: the ring store has read it (1)
    B_XO_FULL,       // tap tile landed (tx)
    B_XO_FREE,       // older-tap MMAs done with it (commit)
    B_GATE0, B_GATE1,   // gate accumulators complete, per TMEM set (commit)   | head: hidden
    B_Y_FULL,        // epilogue wrote the gated output tile (128)            | head: hidden tile
    B_RS0, B_RS1,    // res / skip accumulators complete (commit)             | head: logits
    B_CR_XB,         // downstream is done with the bf16 tile I sent (4 warps + its copy thread)
    B_CR_XF,         // downstream consumed the fp32 x I sent (4 warps)
    B_CR_SK,         // downstream consumed the skip sum I sent (4 warps)
    B_IN_FREE,       // mailbox-fed: this CTA's epilogue is done with XF / SK (4 warps)
    B_COUNT
};

struct Params {
    int L, CS, NCL, G, Q;
    float min_temp;
    int dil[MAXL];
    long long ring_off[MAXL];      // byte offset of layer l's ring: [d + 1 slots][G][TILE_B]
    const unsigned char* wpack;    // [L][6 tiles] then the head's 4 tiles
    const float* E;                // (Q, 128)
    const float* b1;               // [L][256] gate biases with the residual-conv biases folded in (filter rows, gate rows)
    const float* cbs;              // [128] sum of the skip-conv biases
    const float* hb1; const float* hb2;
    unsigned char* rings;
    float* mail;                   // [(slot, g, parity)][2][128][16] fp32: x block, skip block (cluster boundaries)
    unsigned* mail_flag;           // [(slot, g, parity)] warps that have published
    unsigned* ack;                 // [slot][g] warps that have consumed
    unsigned long long* samples;   // [G][16] words {index, tag}
    unsigned* abort_flag;
    // this run
    long long* seq;
    long long seq_stride, t_begin, t_head, t_end;
    int B, n_groups, teacher_forced, n_temperature;
    const float* temperature; const float* noise;
    long long noise_stride, noise_t0;
    float* logits_out; long long* decisions; unsigned long long* step_ts;
    long long* trace; long long trace_t;
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
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned raddr) {   // relaxed: see wavenet6.cu
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
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
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 ld_poll_v2(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flagged_v2(uint2* p, unsigned a, unsigned tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_async_b32(unsigned raddr, unsigned v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_v4(unsigned raddr, float4 v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rbar) : "memory");
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
__device__ __noinline__ bool count_wait(const unsigned* p, unsigned target, unsigned* abort_flag) {
    return spin_until([&] { return ld_acquire_u32(p) >= target; }, abort_flag);
}
__device__ __noinline__ bool poll_word(const uint2* p, unsigned tag, uint2& v, unsigned* abort_flag) {
    return spin_until([&] { v = ld_poll_v2(p); return v.y == tag; }, abort_flag);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared::cta -> another CTA's shared memory (async proxy end to end), completing on that CTA's mbarrier
__device__ __forceinline__ void bulk_s2s(unsigned rdst, unsigned src, unsigned bytes, unsigned rbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (see wavenet_tc.cu): a tile of R rows x K bf16 = K/64 atoms of
// R x 128-byte lines, 8 rows = 1024 bytes (SBO), chunk c of line r at position c ^ (r & 7); atoms R * 128 bytes apart.
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
__host__ __device__ __forceinline__ unsigned tile_off(int m, int kc, int R) {
    return (unsigned)((kc >> 3) * (R * 128) + (m >> 3) * 1024 + (m & 7) * 128 + (((kc & 7) ^ (m & 7)) << 4));
}
__device__ __forceinline__ float tanh_fast(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float gate_fast(float f, float g) {   // tanh(f) * sigmoid(g), sigmoid(g) = 0.5 tanh(g / 2) + 0.5
    return tanh_fast(f) * fmaf(0.5f, tanh_fast(0.5f * g), 0.5f);
}
__device__ __forceinline__ float mish_fast(float x) {
    const float sp = x > 20.0f ? x : __logf(1.0f + __expf(x));
    return x * tanh_fast(sp);
}
__device__ __forceinline__ unsigned short bf16_bits(float v) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<unsigned short*>(&b);
}
// thread c writes its 16 values (one per prompt) into a bf16 B tile [16 prompts x 128 channels] at channel c
__device__ __forceinline__ void store_column_bf16(unsigned char* tile, int c, const float (&v)[16]) {
#pragma unroll
    for (int n = 0; n < NP; ++n)
        *reinterpret_cast<unsigned short*>(tile + tile_off(n, c >> 3, NP) + (c & 7) * 2) = bf16_bits(v[n]);
}

// thread (c, half) writes its 8 values (prompts p0 .. p0 + 7) of channel c
__device__ __forceinline__ void store_half_column_bf16(unsigned char* tile, int c, int p0, const float (&v)[8]) {
#pragma unroll
    for (int n = 0; n < 8; ++n)
        *reinterpret_cast<unsigned short*>(tile + tile_off(p0 + n, c >> 3, NP) + (c & 7) * 2) = bf16_bits(v[n]);
}

// ------------------------------------------------------------------------------------------------------------
// The kernel: CTA (cluster, rank) = pipeline slot cluster * CS + rank: layers 0 .. L-1, then the head.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) wavenet7_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const unsigned rank = cluster_ctarank();
    const int cluster = blockIdx.x / P.CS;
    const int slot = cluster * P.CS + (int)rank;
    const int L = P.L, G = P.G;
    const unsigned sb = smem_u32(smem);
    auto bar = [&](int i) { return sb + (unsigned)SM_BAR + 8u * (unsigned)i; };
    auto window = [&](int r) { return mapa(sb, (unsigned)r) - sb; };
    unsigned* s_tmem = reinterpret_cast<unsigned*>(smem + SM_BAR + 8 * B_COUNT);
    unsigned* abort_flag = P.abort_flag;
    const bool is_layer = slot < L, is_head = slot == L;
    const bool first = slot == 0, last_layer = slot == L - 1;
    const bool up_local = slot > 0 && (slot - 1) / P.CS == cluster;       // fed through DSMEM (else: embedding or mailbox)
    const bool down_local = (slot + 1) / P.CS == cluster;
    const bool mail_fed = slot > 0 && !up_local;

    if (tid == 0) {
        if (sb & 1023u) atomicExch(abort_flag, 2u);       // the swizzled tiles need a 1024-byte aligned window
        mbar_init(bar(B_XF), 1); mbar_init(bar(B_SK), 1);
        const bool built_here = !is_layer || first || mail_fed;           // else the layer above writes the tile remotely
        mbar_init(bar(B_XB_FULL), built_here ? (is_layer ? NEPI_L : NEPI_H) : 1);
        mbar_init(bar(B_XB_FREE), 1);
        mbar_init(bar(B_CR_XB), EW_L + 1);                                // the layer below: its epilogue warps + its copy thread
        mbar_init(bar(B_XO_FULL), 1); mbar_init(bar(B_XO_FREE), 1);
        mbar_init(bar(B_GATE0), 1); mbar_init(bar(B_GATE1), 1);
        mbar_init(bar(B_Y_FULL), is_layer ? NEPI_L : NEPI_H);
        mbar_init(bar(B_RS0), 1); mbar_init(bar(B_RS1), 1);
        mbar_init(bar(B_CR_XF), EW_L);
        mbar_init(bar(B_CR_SK), slot + 1 == L ? EW_H : EW_L);             // the consumer of my skip sum: the head or a layer
        mbar_init(bar(B_IN_FREE), EW_L);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (up_local) {   // arm the first phase of the DSMEM-fed inputs (tx bytes may land before or after)
            if (is_layer) {
                mbar_expect_tx(bar(B_XF), TILE_F);
                mbar_expect_tx(bar(B_XB_FULL), TILE_B);
            }
            mbar_expect_tx(bar(B_SK), TILE_F);
        }
        fence_proxy_async();
    }
    // ---- resident weights (generic-proxy stores, fenced for the tensor core below)
    if (is_layer || is_head) {
        const int n_tiles = is_layer ? 6 : 4;
        const uint4* src = reinterpret_cast<const uint4*>(P.wpack + (size_t)slot * 6 * TILE_A);
        uint4* dst = reinterpret_cast<uint4*>(smem + SM_W);
        for (int i = tid; i < n_tiles * TILE_A / 16; i += NT) dst[i] = __ldg(src + i);
        if (is_layer) {   // a zero tap tile: the first launches read the ring before anything was stored
            for (int i = tid; i < TILE_B / 16; i += NT) reinterpret_cast<uint4*>(smem + SM_XO)[i] = make_uint4(0, 0, 0, 0);
        }
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(s_tmem), 128);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<volatile unsigned*>(s_tmem);
    cluster_sync_all();

    bool dead = false;
    long long* trace_row = nullptr;
    int trace_n = 0;
    auto stamp = [&]() { if (trace_row && trace_n < TRACE_EV) trace_row[trace_n++] = clock64(); };

    // downstream addresses (DSMEM): the next slot's XF / SK buffers and barriers
    const unsigned win_dn = down_local ? window((slot + 1) % P.CS) : 0u;
    const unsigned win_up = up_local ? window((slot - 1) % P.CS) : 0u;
    const size_t n_units_step = (size_t)P.n_groups;
    const long long t0_head = P.t_begin > P.t_head ? P.t_begin : P.t_head;

    if (is_layer) {
        const int l = slot;
        const int d = P.dil[l];
        unsigned char* ring = P.rings + P.ring_off[l];
        auto ring_tile = [&](long long t, int g) -> unsigned char* {
            long long s = t % (long long)(d + 1);
            if (s < 0) s += d + 1;
            return ring + ((size_t)s * G + g) * TILE_B;
        };
        if (warp < EW_L) {
            // =================================================================================================
            // epilogue warps: thread = (channel c = TMEM lane c, half ph of the group's 16 prompts): warps w and w + 4 share the
            // TMEM lane quarter w and split the columns, so an epilogue is 8 values per thread instead of 16
            // =================================================================================================
            const int c = tid & 127, ph = tid >> 7, p0 = ph * NPH;
            const unsigned tm_lane = tmem + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)p0;
            const float b_f = __ldg(P.b1 + (size_t)l * 256 + c), b_g = __ldg(P.b1 + (size_t)l * 256 + 128 + c);
            const bool built_here = first || mail_fed;
            unsigned n = 0, nh = 0;
            for (long long t = P.t_begin; t < P.t_end && !dead; ++t) {
                const unsigned delivery = (unsigned)(t - P.t_begin);
                const unsigned tag = delivery + 1u;
                const bool head_on = t >= P.t_head;
                const bool flow = P.teacher_forced || t <= P.t_head;
                for (int g = 0; g < P.n_groups && !dead; ++g, ++n) {
                    trace_row = (P.trace && t == P.trace_t && tid == 0) ? P.trace + ((size_t)slot * G + g) * TRACE_EV : nullptr;
                    trace_n = 0;
                    if (trace_row) trace_row[trace_n++] = (long long)globaltimer();
                    stamp();
                    const unsigned nb = n & 1u;
                    unsigned char* xb = smem + SM_XB;
                    float xf[NPH];
                    // ---- E0 (first layer / cluster-boundary layer only): build the bf16 B tile of the layer input here.  Every
                    //      other layer receives that tile ready-made from the layer above (st.async, straight into xb).
                    if (built_here) {
                        if (first) {
                            // embedding gather (EmbeddingIO, modules/io.py:148-154): x[c][p] = E[q_{b,t}][c]
                            long long q[NPH];
                            if (!P.teacher_forced && t > P.t_head) {
                                uint2 w[NPH];
#pragma unroll
                                for (int p = 0; p < NPH; ++p) w[p] = ld_poll_v2(reinterpret_cast<const uint2*>(P.samples + g * NP + p0 + p));
#pragma unroll
                                for (int p = 0; p < NPH; ++p) {
                                    if (g * NP + p0 + p < P.B && w[p].y != tag)
                                        dead |= !poll_word(reinterpret_cast<const uint2*>(P.samples + g * NP + p0 + p), tag, w[p], abort_flag);
                                    q[p] = (long long)w[p].x;
                                }
                            } else {
#pragma unroll
                                for (int p = 0; p < NPH; ++p) {
                                    const int b = g * NP + p0 + p;
                                    q[p] = b < P.B ? __ldcg(P.seq + (size_t)b * P.seq_stride + t) : 0;
                                }
                            }
#pragma unroll
                            for (int p = 0; p < NPH; ++p) {
                                const long long qi = q[p] < 0 ? 0 : (q[p] >= P.Q ? P.Q - 1 : q[p]);
                                xf[p] = (g * NP + p0 + p < P.B) ? __ldg(P.E + (size_t)qi * CC + c) : 0.0f;
                            }
                        } else {
                            dead |= !mbar_wait(bar(B_XF), n & 1u, abort_flag);
                            const float4* s4 = reinterpret_cast<const float4*>(smem + SM_XF) + c * 4 + 2 * ph;
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const float4 v = s4[i];
                                xf[4 * i] = v.x; xf[4 * i + 1] = v.y; xf[4 * i + 2] = v.z; xf[4 * i + 3] = v.w;
                            }
                        }
                        stamp();
                        if (n >= 1u) dead |= !mbar_wait(bar(B_XB_FREE), (n - 1u) & 1u, abort_flag);   // ring store has read the old tile
                        store_half_column_bf16(xb, c, p0, xf);
                        fence_proxy_async_smem();
                        mbar_arrive(bar(B_XB_FULL));
                    }
                    stamp();
                    // ---- E1: gate.  D_f | D_g = both taps (the older one was issued a unit ahead)
                    dead |= !mbar_wait(bar(B_GATE0 + nb), (n >> 1) & 1u, abort_flag);
                    tc_fence_after();
                    if (!built_here && tid == 0) mbar_expect_tx(bar(B_XB_FULL), TILE_B);   // the newer-tap MMAs have read the tile: re-arm
                    stamp();
                    float y[NPH];
                    {
                        float f8[NPH], g8[NPH];
                        tmem_ld8(tm_lane + 64u * nb, f8);
                        tmem_ld8(tm_lane + 64u * nb + 16u, g8);
                        tmem_ld_wait();
#pragma unroll
                        for (int p = 0; p < NPH; ++p) y[p] = gate_fast(f8[p] + b_f, g8[p] + b_g);
                    }
                    // the y tile doubled as the staging area of the bf16 tile sent downstream in the previous unit: that bulk copy is
                    // certainly over once the layer below has returned the tile (its credit also frees its x tile for this unit's)
                    if (!last_layer && down_local && n >= 1u) dead |= !mbar_wait(bar(B_CR_XB), (n - 1u) & 1u, abort_flag);
                    store_half_column_bf16(smem + SM_YB, c, p0, y);
                    fence_proxy_async_smem();
                    tc_fence_before();
                    mbar_arrive(bar(B_Y_FULL));
                    stamp();
                    // ---- E2: residual stream and skip sum of this layer
                    if (!built_here) {   // the fp32 stream of the group arrived right behind its bf16 tile
                        dead |= !mbar_wait(bar(B_XF), n & 1u, abort_flag);
                        if (tid == 0) mbar_expect_tx(bar(B_XF), TILE_F);
                        const float4* s4 = reinterpret_cast<const float4*>(smem + SM_XF) + c * 4 + 2 * ph;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float4 v = s4[i];
                            xf[4 * i] = v.x; xf[4 * i + 1] = v.y; xf[4 * i + 2] = v.z; xf[4 * i + 3] = v.w;
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(win_up + bar(B_CR_XF));   // XF may be refilled
                        // ... and the bf16 tile is returned with it.  (Returning the tile as early as the gate epilogue — the MMAs are
                        // done with it by then — produced wrong logits whenever the pipeline was back-pressured; returned here, next
                        // to the stream it belongs to, every geometry replays bit for bit.)
                        if (lane == 0) mbar_arrive_remote(win_up + bar(B_CR_XB));
                    }
                    dead |= !mbar_wait(bar(B_RS0 + nb), (n >> 1) & 1u, abort_flag);
                    tc_fence_after();
                    stamp();
                    const unsigned cnt_dn = last_layer ? nh : n;             // deliveries made downstream so far
                    const bool sends = !last_layer || head_on;
                    const size_t box_dn = (((size_t)(slot + 1) * G + g) * 2 + (delivery & 1u));
                    float rs[NPH], sk[NPH];
                    tmem_ld8(tm_lane + 64u * nb + 32u, rs);
                    tmem_ld8(tm_lane + 64u * nb + 48u, sk);
                    tmem_ld_wait();
                    if (!last_layer) {
#pragma unroll
                        for (int p = 0; p < NPH; ++p) xf[p] += rs[p];        // h_{l+1} = h_l + conv_res(y); biases folded (see b1)
                        if (down_local) {
                            // the bf16 tile first (the next layer's MMAs wait for nothing else).  It must reach the other CTA through
                            // the ASYNC proxy: a tile written with st.async (generic proxy) was not reliably seen by the tensor core
                            // there, whatever proxy fence the consumer used.  So it is staged here — in the y tile, idle since the
                            // res / skip MMAs finished — and one thread sends it with a shared::cta -> shared::cluster bulk copy.
                            store_half_column_bf16(smem + SM_YB, c, p0, xf);
                            fence_proxy_async_smem();
                            asm volatile("bar.sync 3, 256;" ::: "memory");
                            stamp();
                            if (tid == 0) bulk_s2s(win_dn + sb + (unsigned)SM_XB, sb + SM_YB, TILE_B, win_dn + bar(B_XB_FULL));
                            stamp();
                            if (cnt_dn >= 1u) dead |= !mbar_wait(bar(B_CR_XF), (cnt_dn - 1u) & 1u, abort_flag);
                            const unsigned dst = win_dn + sb + (unsigned)SM_XF + (unsigned)c * 64u + 32u * (unsigned)ph, rb = win_dn + bar(B_XF);
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                st_async_v4(dst + 16u * i, make_float4(xf[4 * i], xf[4 * i + 1], xf[4 * i + 2], xf[4 * i + 3]), rb);
                        } else {
                            if (flow && delivery >= 2u) {
                                if (lane == 0) dead |= !count_wait(P.ack + (size_t)(slot + 1) * G + g, (unsigned)EW_L * (delivery - 1u), abort_flag);
                                dead = __any_sync(0xffffffffu, dead);
                            }
                            float4* mx = reinterpret_cast<float4*>(P.mail + box_dn * (size_t)(2 * CC * NP)) + c * 4 + 2 * ph;
#pragma unroll
                            for (int i = 0; i < 2; ++i) __stcg(mx + i, make_float4(xf[4 * i], xf[4 * i + 1], xf[4 * i + 2], xf[4 * i + 3]));
                        }
                    }
                    stamp();
                    {
                        if (!first) {
                            dead |= !mbar_wait(bar(B_SK), n & 1u, abort_flag);
                            if (up_local && tid == 0) mbar_expect_tx(bar(B_SK), TILE_F);
                            const float4* s4 = reinterpret_cast<const float4*>(smem + SM_SK) + c * 4 + 2 * ph;
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const float4 v = s4[i];
                                sk[4 * i] += v.x; sk[4 * i + 1] += v.y; sk[4 * i + 2] += v.z; sk[4 * i + 3] += v.w;
                            }
                            __syncwarp();
                            if (up_local) { if (lane == 0) mbar_arrive_remote(win_up + bar(B_CR_SK)); }
                            else if (lane == 0) { mbar_arrive(bar(B_IN_FREE)); red_add_u32(P.ack + (size_t)slot * G + g, 1u); }
                        }
                        if (sends) {
                            if (down_local) {
                                if (cnt_dn >= 1u) dead |= !mbar_wait(bar(B_CR_SK), (cnt_dn - 1u) & 1u, abort_flag);
                                const unsigned dst = win_dn + sb + (unsigned)SM_SK + (unsigned)c * 64u + 32u * (unsigned)ph, rb = win_dn + bar(B_SK);
#pragma unroll
                                for (int i = 0; i < 2; ++i)
                                    st_async_v4(dst + 16u * i, make_float4(sk[4 * i], sk[4 * i + 1], sk[4 * i + 2], sk[4 * i + 3]), rb);
                            } else {
                                if (last_layer && flow && (unsigned)(t - t0_head) >= 2u) {
                                    if (lane == 0) dead |= !count_wait(P.ack + (size_t)(slot + 1) * G + g, (unsigned)EW_H * ((unsigned)(t - t0_head) - 1u), abort_flag);
                                    dead = __any_sync(0xffffffffu, dead);
                                }
                                float4* ms = reinterpret_cast<float4*>(P.mail + box_dn * (size_t)(2 * CC * NP) + CC * NP) + c * 4 + 2 * ph;
#pragma unroll
                                for (int i = 0; i < 2; ++i) __stcg(ms + i, make_float4(sk[4 * i], sk[4 * i + 1], sk[4 * i + 2], sk[4 * i + 3]));
                                __syncwarp();
                                if (lane == 0) red_release_add_u32(P.mail_flag + box_dn, 1u);   // publishes the x block as well
                            }
                        }
                    }
                    tc_fence_before();
                    stamp();
                    if (last_layer && tid == 0 && !head_on && g == P.n_groups - 1 && P.step_ts) P.step_ts[t - P.t_begin] = globaltimer();
                    if (head_on) ++nh;
                }
            }
        } else if (warp == W_MMA) {
            // =================================================================================================
            // MMA issuer: warp-uniform control flow, one elected lane issues
            // =================================================================================================
            const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            auto wait_u = [&](unsigned b, unsigned parity) -> bool {
                return __all_sync(0xffffffffu, mbar_wait(b, parity, abort_flag) ? 1 : 0) != 0;
            };
            auto elect = [&]() -> bool {
                unsigned pred;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                return pred != 0;
            };
            const unsigned id16 = umma_idesc(NP);
            const unsigned long long dW = umma_desc(sb + SM_W);
            const unsigned long long dXB0 = umma_desc(sb + SM_XB), dXO = umma_desc(sb + SM_XO), dYB = umma_desc(sb + SM_YB);
            constexpr unsigned T16 = TILE_A >> 4;
            const size_t n_units = (size_t)(P.t_end - P.t_begin) * n_units_step;
            // older tap of unit u into TMEM set u & 1 (independent of this step's activations: issued a unit ahead)
            auto issue_old = [&](size_t u) -> bool {
                if (!wait_u(bar(B_XO_FULL), (unsigned)u & 1u)) return false;
                tc_fence_after();
                const unsigned ds = tmem_u + 64u * ((unsigned)u & 1u);
                if (elect()) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) umma_bf16(ds, dW + W_OF * T16 + kstep16(kk, 128), dXO + kstep16(kk, NP), id16, kk > 0);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) umma_bf16(ds + 16u, dW + W_OG * T16 + kstep16(kk, 128), dXO + kstep16(kk, NP), id16, kk > 0);
                    umma_commit(bar(B_XO_FREE));
                }
                __syncwarp();
                return true;
            };
            if (n_units > 0) dead = !issue_old(0);
            for (size_t u = 0; u < n_units && !dead; ++u) {
                const unsigned nb = (unsigned)u & 1u, ds = tmem_u + 64u * nb;
                if (!wait_u(bar(B_XB_FULL), (unsigned)u & 1u)) break;
                tc_fence_after();
                const unsigned long long dXB = dXB0;
                if (elect()) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) umma_bf16(ds, dW + W_NF * T16 + kstep16(kk, 128), dXB + kstep16(kk, NP), id16, 1u);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) umma_bf16(ds + 16u, dW + W_NG * T16 + kstep16(kk, 128), dXB + kstep16(kk, NP), id16, 1u);
                    umma_commit(bar(B_GATE0 + nb));
                }
                __syncwarp();
                if (!wait_u(bar(B_Y_FULL), (unsigned)u & 1u)) break;
                tc_fence_after();
                if (elect()) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) umma_bf16(ds + 32u, dW + W_RES * T16 + kstep16(kk, 128), dYB + kstep16(kk, NP), id16, kk > 0);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) umma_bf16(ds + 48u, dW + W_SKIP * T16 + kstep16(kk, 128), dYB + kstep16(kk, NP), id16, kk > 0);
                    umma_commit(bar(B_RS0 + nb));
                }
                __syncwarp();
                // the other TMEM set was drained by the epilogue before it wrote the y tile waited for above
                if (u + 1 < n_units && !issue_old(u + 1)) break;
            }
        } else if (warp == W_CP) {
            // =================================================================================================
            // copy thread: ring store / tap fetch, and the mailbox pull of a cluster-boundary layer
            // =================================================================================================
            if (lane == 0) {
                unsigned n = 0;
                auto fetch_tap = [&](long long t, int g) {
                    mbar_expect_tx(bar(B_XO_FULL), TILE_B);
                    bulk_g2s(sb + SM_XO, ring_tile(t - d, g), TILE_B, bar(B_XO_FULL));
                };
                auto pull_mail = [&](long long t, int g) -> bool {
                    const unsigned delivery = (unsigned)(t - P.t_begin);
                    const size_t box = ((size_t)slot * G + g) * 2 + (delivery & 1u);
                    if (!count_wait(P.mail_flag + box, (unsigned)EW_L * (delivery / 2u + 1u), abort_flag)) return false;
                    const float* m = P.mail + box * (size_t)(2 * CC * NP);
                    mbar_expect_tx(bar(B_XF), TILE_F);
                    bulk_g2s(sb + SM_XF, m, TILE_F, bar(B_XF));
                    mbar_expect_tx(bar(B_SK), TILE_F);
                    bulk_g2s(sb + SM_SK, m + CC * NP, TILE_F, bar(B_SK));
                    return true;
                };
                if (P.t_begin < P.t_end) {
                    fetch_tap(P.t_begin, 0);
                    if (mail_fed) dead = !pull_mail(P.t_begin, 0);
                }
                for (long long t = P.t_begin; t < P.t_end && !dead; ++t) {
                    for (int g = 0; g < P.n_groups && !dead; ++g, ++n) {
                        int ng = g + 1;
                        long long nt = t;
                        if (ng == P.n_groups) { ng = 0; ++nt; }
                        // x(t) -> ring, as soon as the epilogue has built the tile
                        if (!mbar_wait(bar(B_XB_FULL), n & 1u, abort_flag)) { dead = true; break; }
                        bulk_s2g(ring_tile(t, g), sb + SM_XB, TILE_B);
                        bulk_commit();
                        bulk_wait_read();
                        if (first || mail_fed) mbar_arrive(bar(B_XB_FREE));
                        else mbar_arrive_remote(win_up + bar(B_CR_XB));
                        // next unit's tap, once the older-tap MMAs of this unit are done with the tile
                        if (nt < P.t_end) {
                            if (!mbar_wait(bar(B_XO_FREE), n & 1u, abort_flag)) { dead = true; break; }
                            bulk_wait_all();          // with one group and dilation 1 the tile just stored is the one to fetch
                            fetch_tap(nt, ng);
                            if (mail_fed) {
                                if (!mbar_wait(bar(B_IN_FREE), n & 1u, abort_flag)) { dead = true; break; }
                                if (!pull_mail(nt, ng)) { dead = true; break; }
                            }
                        }
                    }
                }
                bulk_wait_all();
            }
        }
    } else if (is_head) {
        // =====================================================================================================
        // head: hidden = mish(W1 (skips + sum of skip biases) + b1) (mlp.py:44-53), z = W2 hidden + b2, sampler
        // =====================================================================================================
        const int Q = P.Q;
        float* Z = reinterpret_cast<float*>(smem + SM_Z);
        const int c = tid & 127;
        const unsigned tm_lane = tmem + ((unsigned)(32 * (warp & 3)) << 16);
        const float cb = warp < 4 ? __ldg(P.cbs + c) : 0.0f, hb1 = warp < 4 ? __ldg(P.hb1 + c) : 0.0f;
        const float b2a = (warp < 4 && c < Q) ? __ldg(P.hb2 + c) : 0.0f;
        const float b2b = (warp < 4 && 128 + c < Q) ? __ldg(P.hb2 + 128 + c) : 0.0f;
        const float b2t = __ldg(P.hb2 + Q);
        const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        auto elect = [&]() -> bool {
            unsigned pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            return pred != 0;
        };
        const unsigned id16 = umma_idesc(NP);
        const unsigned long long dW = umma_desc(sb + SM_W);
        const unsigned long long dXB = umma_desc(sb + SM_XB), dYB = umma_desc(sb + SM_YB);
        constexpr unsigned T16 = TILE_A >> 4;
        unsigned nh = 0;
        for (long long t = t0_head; t < P.t_end && !dead; ++t) {
            const unsigned delivery = (unsigned)(t - P.t_begin);
            const unsigned tag = delivery + 1u;
            for (int g = 0; g < P.n_groups && !dead; ++g, ++nh) {
                trace_row = (P.trace && t == P.trace_t && tid == 0) ? P.trace + ((size_t)slot * G + g) * TRACE_EV : nullptr;
                trace_n = 0;
                if (trace_row) trace_row[trace_n++] = (long long)globaltimer();
                stamp();
                if (warp == W_CP && mail_fed && lane == 0) {
                    const size_t box = ((size_t)slot * G + g) * 2 + (delivery & 1u);
                    if (count_wait(P.mail_flag + box, (unsigned)EW_L * ((unsigned)(t - t0_head) / 2u + 1u), abort_flag)) {
                        mbar_expect_tx(bar(B_SK), TILE_F);
                        bulk_g2s(sb + SM_SK, P.mail + box * (size_t)(2 * CC * NP) + CC * NP, TILE_F, bar(B_SK));
                    }
                }
                if (warp < 4) {
                    dead |= !mbar_wait(bar(B_SK), nh & 1u, abort_flag);
                    if (up_local && tid == 0) mbar_expect_tx(bar(B_SK), TILE_F);
                    float a[16];
                    const float4* s4 = reinterpret_cast<const float4*>(smem + SM_SK) + c * 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = s4[i];
                        a[4 * i] = v.x + cb; a[4 * i + 1] = v.y + cb; a[4 * i + 2] = v.z + cb; a[4 * i + 3] = v.w + cb;
                    }
                    __syncwarp();
                    if (up_local) { if (lane == 0) mbar_arrive_remote(win_up + bar(B_CR_SK)); }
                    else if (lane == 0) red_add_u32(P.ack + (size_t)slot * G + g, 1u);
                    store_column_bf16(smem + SM_XB, c, a);
                    fence_proxy_async_smem();
                    mbar_arrive(bar(B_XB_FULL));
                    stamp();
                    dead |= !mbar_wait(bar(B_GATE0), nh & 1u, abort_flag);
                    tc_fence_after();
                    float hv[16];
                    tmem_ld16(tm_lane, hv);
                    tmem_ld_wait();
#pragma unroll
                    for (int p = 0; p < NP; ++p) hv[p] = mish_fast(hv[p] + hb1);
                    store_column_bf16(smem + SM_YB, c, hv);
                    fence_proxy_async_smem();
                    tc_fence_before();
                    mbar_arrive(bar(B_Y_FULL));
                    stamp();
                    dead |= !mbar_wait(bar(B_RS0), nh & 1u, abort_flag);
                    tc_fence_after();
                    float z0[16], z1[16];
                    tmem_ld16(tm_lane + 16u, z0);
                    tmem_ld16(tm_lane + 32u, z1);
                    tmem_ld_wait();
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        if (c < Q) Z[p * ZROW + c] = z0[p] + b2a;
                        if (128 + c < Q) Z[p * ZROW + 128 + c] = z1[p] + b2b;
                    }
                    if (warp == 0) {   // the learned-temperature row: row 0 (lane 0) of the third tile; the load is warp-collective
                        float zt[16];
                        tmem_ld16(tm_lane + 48u, zt);
                        tmem_ld_wait();
                        if (c == 0) {
#pragma unroll
                            for (int p = 0; p < NP; ++p) Z[p * ZROW + Q] = zt[p] + b2t;
                        }
                    }
                    tc_fence_before();
                    stamp();
                } else if (warp == W_MMA) {
                    bool ok = __all_sync(0xffffffffu, mbar_wait(bar(B_XB_FULL), nh & 1u, abort_flag) ? 1 : 0) != 0;
                    tc_fence_after();
                    if (ok && elect()) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) umma_bf16(tmem_u, dW + kstep16(kk, 128), dXB + kstep16(kk, NP), id16, kk > 0);
                        umma_commit(bar(B_GATE0));
                    }
                    __syncwarp();
                    ok = ok && __all_sync(0xffffffffu, mbar_wait(bar(B_Y_FULL), nh & 1u, abort_flag) ? 1 : 0) != 0;
                    tc_fence_after();
                    if (ok && elect()) {
                        for (int tl = 0; tl < 3; ++tl)
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma_bf16(tmem_u + 16u * (tl + 1), dW + (1 + tl) * T16 + kstep16(kk, 128), dYB + kstep16(kk, NP), id16, kk > 0);
                        umma_commit(bar(B_RS0));
                    }
                    __syncwarp();
                    dead |= !ok;
                }
                // every warp meets here: the logits of the group are staged
                if (__syncthreads_or(dead ? 1 : 0)) { dead = true; break; }
                for (int p = warp; p < NP; p += NT / 32) {
                    const int b = g * NP + p;
                    if (b >= P.B) continue;
                    const long long hstep = t - P.t_head, n_head = P.t_end - P.t_head;
                    float* lout = P.logits_out ? P.logits_out + ((size_t)b * n_head + hstep) * Q : nullptr;
                    const bool sample = P.temperature != nullptr;
                    float Tt = 1.0f, u = 0.0f;
                    if (sample) {
                        Tt = P.temperature[P.n_temperature == 1 ? 0 : b];
                        u = P.noise[(size_t)b * P.noise_stride + (t + 1 - P.noise_t0)];
                    }
                    const int choice = mmk::decide_warp(Z + p * ZROW, Q, P.min_temp, lout, sample, Tt, u);
                    if (lane == 0) {
                        if (P.decisions) P.decisions[(size_t)b * n_head + hstep] = choice;
                        if (!P.teacher_forced) {
                            st_flagged_v2(reinterpret_cast<uint2*>(P.samples + g * NP + p), (unsigned)choice, tag + 1u);
                            __stcg(P.seq + (size_t)b * P.seq_stride + t + 1, (long long)choice);
                        }
                    }
                }
                stamp();
                if (tid == 0 && g == P.n_groups - 1 && P.step_ts) P.step_ts[t - P.t_begin] = globaltimer();
                __syncthreads();     // the staging area is rewritten by the next group
            }
        }
    }
    // no CTA may exit while peers can still store into its shared memory
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, 128);
    cluster_sync_all();
}

static unsigned short f2bf(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
// rows [0, 128) of a canonical bf16 tile of 128 rows x 128 from a strided fp32 matrix (src[r * ld + k * es]); rows >= valid zero
static void pack_tile(unsigned char* out, const float* src, size_t ld, size_t es, int valid) {
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 128; ++k) {
            const unsigned short v = f2bf((src && r < valid) ? src[(size_t)r * ld + (size_t)k * es] : 0.0f);
            memcpy(out + tile_off(r, k / 8, 128) + (k % 8) * 2, &v, 2);
        }
}

}  // namespace mmk7

using namespace mmk7;

struct wn7_handle {
    Params p{};
    int device = 0, max_batch = 0;
    std::vector<void*> allocs;
    void* d_flags = nullptr;
    size_t flags_bytes = 0;
    long long* d_trace = nullptr;
    long long trace_t = -1;
    int trace_rows = 0;
};

int wn7_destroy(wn7_handle* h) {
    if (!h) return 0;
    for (void* a : h->allocs) cudaFree(a);
    delete h;
    return 0;
}

// Returns 1 with *unsupported = 1 when the configuration is not W-30-shaped (the caller falls back to wavenet_tc.cu).
int wn7_create(const mmk_wavenet_desc* d, int max_batch, wn7_handle** out, int* unsupported) {
    *unsupported = 1;
    const int L = d->n_layers, Q = d->q_levels;
    if (d->dilated_dim != CC || d->skips_dim != CC || d->head_hidden != CC || L < 1 || L > MAXL || Q < 2 || Q > 256) return 1;
    for (int l = 0; l < L; ++l) {
        if ((d->conv_res_w[l] != nullptr) != (l < L - 1)) return 1;
        if (!d->conv_skip_w || !d->conv_skip_w[l] || d->dilations[l] < 1) return 1;
    }
    if (getenv("MMK_TC_KERNEL") && atoi(getenv("MMK_TC_KERNEL")) == 4) return 1;   // force the one-CTA-per-128-prompts kernel
    auto* h = new wn7_handle();
    Params& p = h->p;
    cudaGetDevice(&h->device);
    int cc_major = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, h->device);
    if (cc_major != 10) { delete h; return 1; }
    p.L = L; p.Q = Q; p.min_temp = d->min_temperature;
    h->max_batch = max_batch;
    p.G = (max_batch + NP - 1) / NP;
    // cluster geometry: L + 1 slots in clusters of CS; every cluster must be co-resident
    MMK_CUDA(cudaFuncSetAttribute(wavenet7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    MMK_CUDA(cudaFuncSetAttribute(wavenet7_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const char* force_cs = getenv("MMK_TC_CLUSTER");
    int best = 0;
    for (int CS : {16, 8, 4, 2}) {
        if (force_cs && atoi(force_cs) != CS) continue;
        if (!force_cs && CS > 2 && L + 1 <= CS / 2) continue;
        const int ncl = (L + 1 + CS - 1) / CS;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CS * ncl); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = SM_TOTAL;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, wavenet7_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        if (n >= ncl) { best = CS; break; }
    }
    if (!best) { delete h; return 1; }
    p.CS = best; p.NCL = (L + 1 + best - 1) / best;
    *unsupported = 0;

    std::vector<unsigned char> wpack((size_t)(L + 1) * 6 * TILE_A, 0);
    long long ring = 0;
    for (int l = 0; l < L; ++l) {
        p.dil[l] = d->dilations[l];
        p.ring_off[l] = ring;
        ring += (long long)(d->dilations[l] + 1) * p.G * TILE_B;
        const float* wd = d->conv_dil_w[l];   // (256, 128, 2): [o][c][tap], tap 0 = older sample; rows o < 128 filter, o >= 128 gate
        unsigned char* base = wpack.data() + (size_t)l * 6 * TILE_A;
        pack_tile(base + W_NF * TILE_A, wd + 1, (size_t)CC * 2, 2, 128);
        pack_tile(base + W_NG * TILE_A, wd + (size_t)CC * CC * 2 + 1, (size_t)CC * 2, 2, 128);
        pack_tile(base + W_OF * TILE_A, wd, (size_t)CC * 2, 2, 128);
        pack_tile(base + W_OG * TILE_A, wd + (size_t)CC * CC * 2, (size_t)CC * 2, 2, 128);
        pack_tile(base + W_RES * TILE_A, d->conv_res_w[l], CC, 1, d->conv_res_w[l] ? 128 : 0);
        pack_tile(base + W_SKIP * TILE_A, d->conv_skip_w[l], CC, 1, 128);
    }
    {
        unsigned char* base = wpack.data() + (size_t)L * 6 * TILE_A;
        pack_tile(base, d->head_w1, CC, 1, 128);
        pack_tile(base + TILE_A, d->head_w2, CC, 1, std::min(128, Q));
        pack_tile(base + 2 * TILE_A, Q > 128 ? d->head_w2 + (size_t)128 * CC : nullptr, CC, 1, std::max(0, Q - 128));
        pack_tile(base + 3 * TILE_A, d->head_w2 + (size_t)Q * CC, CC, 1, 1);   // the learned-temperature row
    }
    // gate biases with the residual-conv biases folded in, sum of the skip biases (the arithmetic of wavenet_tc.cu and of
    // oracle.restate.WaveNetBf16Oracle: the residual stream carries no bias)
    std::vector<float> b1((size_t)L * 256), cbr((size_t)(L + 1) * CC, 0.0f), cbs(CC, 0.0f);
    for (int l = 0; l < L; ++l) {
        for (int i = 0; i < 256; ++i) {
            double acc = d->conv_dil_b[l][i];
            for (int c = 0; c < CC; ++c)
                acc += ((double)d->conv_dil_w[l][((size_t)i * CC + c) * 2] + (double)d->conv_dil_w[l][((size_t)i * CC + c) * 2 + 1]) *
                       (double)cbr[(size_t)l * CC + c];
            b1[(size_t)l * 256 + i] = (float)acc;
        }
        for (int i = 0; i < CC; ++i)
            cbr[(size_t)(l + 1) * CC + i] = cbr[(size_t)l * CC + i] + (d->conv_res_w[l] ? d->conv_res_b[l][i] : 0.0f);
        for (int i = 0; i < CC; ++i) cbs[i] += d->conv_skip_b[l][i];
    }
    bool ok = true;
    auto dev_alloc = [&](size_t bytes, const void* src) -> void* {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, std::max<size_t>(bytes, 16)) != cudaSuccess) { ok = false; return nullptr; }
        h->allocs.push_back(ptr);
        if (src) cudaMemcpy(ptr, src, bytes, cudaMemcpyHostToDevice); else cudaMemset(ptr, 0, std::max<size_t>(bytes, 16));
        return ptr;
    };
    p.wpack = (const unsigned char*)dev_alloc(wpack.size(), wpack.data());
    p.E = (const float*)dev_alloc((size_t)Q * CC * 4, d->embedding);
    p.b1 = (const float*)dev_alloc(b1.size() * 4, b1.data());
    p.cbs = (const float*)dev_alloc(cbs.size() * 4, cbs.data());
    p.hb1 = (const float*)dev_alloc((size_t)CC * 4, d->head_b1);
    p.hb2 = (const float*)dev_alloc((size_t)(Q + 1) * 4, d->head_b2);
    p.rings = (unsigned char*)dev_alloc((size_t)ring, nullptr);
    const size_t boxes = (size_t)(L + 2) * p.G * 2;
    p.mail = (float*)dev_alloc(boxes * 2 * CC * NP * sizeof(float), nullptr);
    const size_t mf_bytes = boxes * sizeof(unsigned), ack_bytes = (size_t)(L + 2) * p.G * sizeof(unsigned);
    const size_t sw_bytes = (size_t)p.G * NP * sizeof(unsigned long long);
    h->flags_bytes = sw_bytes + mf_bytes + ack_bytes + 16;
    h->d_flags = dev_alloc(h->flags_bytes, nullptr);
    if (const char* e = getenv("MMK_TC_TRACE_T")) {
        h->trace_t = atoll(e);
        h->trace_rows = (L + 1) * p.G;
        h->d_trace = (long long*)dev_alloc((size_t)h->trace_rows * TRACE_EV * sizeof(long long), nullptr);
    }
    if (!ok) { wn7_destroy(h); MMK_FAIL("cudaMalloc failed while creating the bf16 WaveNet handle"); }
    char* f = (char*)h->d_flags;
    p.samples = (unsigned long long*)f; f += sw_bytes;
    p.mail_flag = (unsigned*)f; f += mf_bytes;
    p.ack = (unsigned*)f; f += ack_bytes;
    p.abort_flag = (unsigned*)f;
    MMK_CUDA(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

int wn7_launch_info(wn7_handle* h, mmk_launch_info* out) {
    out->cluster_size = h->p.CS; out->n_stages = h->p.L + 1; out->group_size = NP; out->threads = NT;
    out->smem_bytes = SM_TOTAL; out->sm_used = h->p.L + 1;
    return 0;
}

int wn7_sync_check(wn7_handle* h, void* stream) {
    unsigned aborted = 0;
    MMK_CUDA(cudaMemcpyAsync(&aborted, h->p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MMK_CHECK(aborted == 0, "bf16 WaveNet pipeline kernel watchdog fired: a wait timed out (results invalid)");

    if (h->d_trace) {
        const size_t n = (size_t)h->trace_rows * TRACE_EV;
        std::vector<long long> tr(n);
        MMK_CUDA(cudaMemcpy(tr.data(), h->d_trace, n * sizeof(long long), cudaMemcpyDeviceToHost));
        const char* path = getenv("MMK_TC_TRACE_FILE");
        if (FILE* f = fopen(path ? path : "tc_trace.txt", "w")) {
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

int wn7_run(wn7_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
            int64_t t_end, int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Params p = h->p;
    MMK_CHECK(B <= h->max_batch, "batch exceeds the max_batch the handle was created for");
    p.seq = reinterpret_cast<long long*>(d_seq) - seq_t0;
    p.seq_stride = seq_stride; p.t_begin = t_begin; p.t_head = t_head; p.t_end = t_end;
    p.B = B; p.n_groups = (B + NP - 1) / NP; p.teacher_forced = teacher_forced ? 1 : 0;
    p.temperature = d_temperature; p.n_temperature = n_temperature;
    p.noise = d_noise; p.noise_stride = noise_stride; p.noise_t0 = noise_t0;
    p.logits_out = d_logits_out; p.decisions = reinterpret_cast<long long*>(d_decisions); p.step_ts = d_step_ts;
    p.trace = h->d_trace; p.trace_t = h->d_trace ? t_begin + h->trace_t : -1;
    MMK_CUDA(cudaMemsetAsync(h->d_flags, 0, h->flags_bytes, st));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.CS * p.NCL);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = SM_TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[] = {&p};
    MMK_CUDA(cudaLaunchKernelExC(&cfg, (const void*)wavenet7_kernel, args));
    return 0;
}
