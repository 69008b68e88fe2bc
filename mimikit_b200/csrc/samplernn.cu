// Persistent SampleRNN generation kernel (sm_100a).
//
// Replaces the reference's per-sample Python loop: SampleRNN.before_generate / generate_step
// (mimikit/networks/sample_rnn_v2.py:226-260) with SampleRNNTier.forward (83-99), FramedLinearIO / FramedConv1dIO
// (modules/io.py:106-133, 185-198), nn.GRU single steps, LinearResampler (modules/resamplers.py:13-23), the MLP head
// (networks/mlp.py:44-63) and CategoricalSampler (modules/targets.py:40-52), driven as GenerateLoopV2.run does
// (loops/generate.py:184-229).  Math: SURVEY.md App. A.2.
//
// Design:
//   * ONE cooperative launch covers the warm-up over the prompt and all generated samples.  NC CTAs (one per SM)
//     each keep a row slice of every weight matrix resident in shared memory for the whole launch: GRU rows are
//     split by hidden index (a CTA owns r/z/n rows of its indices, so the gate math is local), up-sampler and head
//     rows are split evenly.  19.3 MB of fp32 weights for the (8,2,1)/512 network live in 128 x 157 KB of smem.
//   * Activations (hidden states, up-sampled conditioning blocks, head hidden, raw logits) are tiny and live in
//     L2-resident global buffers, k-major ([dim][B]) so a prompt chunk loads as contiguous rows.  A stage ends
//     with a grid barrier (one release-add + acquire-spin on a single counter); stages per sample: head 3, plus 2
//     per firing frame tier (tier i fires when t % frame_size_i == 0).
//   * Tier inputs x = Linear(frame) (+ conditioning) are rebuilt on the fly while staging a prompt chunk — K is
//     only 1..8 — which saves one barrier per tier firing.
//   * fp32 FFMA with a fixed summation order (deterministic); gate, Mish, learned temperature, softmax and
//     inverse-CDF sampling from externally supplied noise are fused into the stages.
#include "common.cuh"
#include "sampler.cuh"
#include "tile_gemm.cuh"
#include "samplernn_impl.h"
#include "../../include/mmk_b200.h"

#include <algorithm>
#include <vector>

namespace mmk {

constexpr int SR_NT = 256;       // threads per CTA
constexpr int SR_PB = 16;        // prompts per staged chunk
constexpr int SR_MAX_TIERS = 6;  // frame tiers (the sample-level tier is separate)
constexpr int SR_MAX_RNN = 4;    // stacked recurrent layers per tier (n_rnn)
constexpr unsigned SR_SPIN_LIMIT = 1u << 26;

using Tile = TileGemm<SR_PB, SR_NT>;

struct SrTier {
    int fs, up, kdiv;             // frame size, up-sampling factor, fs_{i-1}/fs_i (1 for the top tier)
    int NU, up_rows;              // padded up-sampler columns per CTA, total up-sampler rows (up * H)
    int off_wih[SR_MAX_RNN], off_whh[SR_MAX_RNN], off_wup;  // float offsets in the CTA's weight block: W[H][N] followed by bias[N]
    const float* in_w;            // (H, fs) row-major, global
    const float* in_b;            // (H)
    float* hbuf[SR_MAX_RNN];      // per layer: [2][H][Bp] ping-pong hidden state
    float* cbuf[SR_MAX_RNN];      // LSTM: per layer [2][H][Bp] ping-pong cell state
    float* obuf;                  // [up*H][Bp] up-sampled block
};

struct SrParams {
    int n_ft, H, Hh, Q, NC, JP, NG, NH1, NZ, fs_last;
    int rnn_type, G, n_rnn, n_hh;              // cell (MMK_RNN_*), gates per cell (3 | 4 | 1), layers per tier, head hidden layers
    int off_w1, off_w2, off_wh, cta_block;     // float offsets / size of the per-CTA weight block
    int off_x, off_part, off_gi, off_lin, smem_floats;
    const float* wpack;
    const float* conv_w;          // (H, fs_last)
    const float* conv_b;          // (H)
    float* hid;                   // [2][Hh][Bp] (ping-pong over the head's hidden layers)
    float* z;                     // [Q+1][Bp]
    unsigned long long* bar;      // grid barrier counter
    unsigned* abort_flag;
    float min_temp;
    SrTier tiers[SR_MAX_TIERS];
    // this run
    int B, Bp, teacher_forced, n_temperature;
    int hsel[SR_MAX_TIERS];
    long long* seq;
    long long seq_stride, warm_begin, warm_end, warm_off, gen_begin, gen_end;
    const float* temperature;
    const float* noise;
    long long noise_stride, noise_t0;
    float* logits_out;
    long long* decisions;
    unsigned long long* step_ts;
};

__host__ __device__ __forceinline__ int part_lo(int c, int rows, int nc) { return (int)(((long long)c * rows) / nc); }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long sr_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// All CTAs of the (co-resident, cooperative) grid meet here.  `epoch` is the running arrival target.
__device__ __forceinline__ void grid_barrier(const SrParams& P, unsigned long long& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += (unsigned long long)P.NC;
        __threadfence();
        red_release_add_u64(P.bar, 1ull);
        unsigned spins = 0;
        while (ld_acquire_u64(P.bar) < epoch) {
            if (++spins > SR_SPIN_LIMIT) {          // watchdog: let everybody through and flag the launch
                atomicExch(P.abort_flag, 1u);
                atomicOr(P.bar, 1ull << 62);
                break;
            }
        }
    }
    __syncthreads();
}

// Linearizer (modules/io.py:111-112): ((q / Q) - .5) * 2 in fp32
__device__ __forceinline__ float linearize(long long q, float Qf) {
    return __fmul_rn(__fsub_rn(__fdiv_rn((float)q, Qf), 0.5f), 2.0f);
}

// xs[k][p] = (sum_f w[k][f] * lin(seq[b][tw - fs + f]) + bias[k]) (+ cond[k][b]) for the chunk starting at prompt b0
__device__ __forceinline__ void stage_frame_input(const SrParams& P, float* xs, float* lin_s, const float* w,
                                                  const float* bias, int fs, long long tw, const float* cond, int b0) {
    const int tid = threadIdx.x;
    const float Qf = (float)P.Q;
    for (int i = tid; i < SR_PB * fs; i += SR_NT) {
        const int p = i / fs, f = i - p * fs, b = b0 + p;
        long long q = 0;
        if (b < P.B) q = __ldcg(P.seq + (size_t)b * P.seq_stride + (tw - fs + f));
        lin_s[i] = linearize(q, Qf);
    }
    __syncthreads();
    for (int i = tid; i < P.H * SR_PB; i += SR_NT) {
        const int k = i / SR_PB, p = i - k * SR_PB;
        float acc = 0.0f;
        for (int f = 0; f < fs; ++f) acc = fmaf(lin_s[p * fs + f], __ldg(w + (size_t)k * fs + f), acc);
        acc += __ldg(bias + k);
        if (cond) acc += __ldcg(cond + (size_t)k * P.Bp + b0 + p);
        xs[i] = acc;
    }
    __syncthreads();
}

// xs[k][p] = src[k][b0 + p] for k < K (src is another CTA's output: read through L2)
__device__ __forceinline__ void stage_rows(float* xs, const float* src, int K, int Bp, int b0) {
    for (int i = threadIdx.x; i < K * (SR_PB / 4); i += SR_NT) {
        const int k = i / (SR_PB / 4), q4 = i - k * (SR_PB / 4);
        reinterpret_cast<float4*>(xs)[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * Bp + b0) + q4);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SR_NT, 1) samplernn_kernel(const __grid_constant__ SrParams P) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, NC = P.NC, H = P.H, Bp = P.Bp;
    float* w_s = smem;
    float* xs = smem + P.off_x;        // [H][PB] staged chunk (also: per-warp logits rows while sampling)
    float* part = smem + P.off_part;   // [NT*8]
    float* gi_s = smem + P.off_gi;     // [NG][PB]
    float* lin_s = smem + P.off_lin;   // [PB][fs_max]

    {   // resident weights: straight copy of this CTA's packed block
        const float4* src = reinterpret_cast<const float4*>(P.wpack + (size_t)c * P.cta_block);
        for (int i = tid; i < P.cta_block / 4; i += SR_NT) reinterpret_cast<float4*>(w_s)[i] = __ldg(src + i);
    }
    __syncthreads();

    const int j_lo = part_lo(c, H, NC), nj = part_lo(c + 1, H, NC) - j_lo;
    const int h1_lo = part_lo(c, P.Hh, NC), nh1 = part_lo(c + 1, P.Hh, NC) - h1_lo;
    const int z_lo = part_lo(c, P.Q + 1, NC), nz = part_lo(c + 1, P.Q + 1, NC) - z_lo;
    const int n_chunks = Bp / SR_PB;
    int hsel[SR_MAX_TIERS];
    for (int i = 0; i < SR_MAX_TIERS; ++i) hsel[i] = P.hsel[i];
    unsigned long long epoch = 0;

    for (int phase = 0; phase < 2; ++phase) {
        const bool gen = phase == 1;
        const long long t_lo = gen ? P.gen_begin : P.warm_begin, t_hi = gen ? P.gen_end : P.warm_end;
        const long long off = gen ? 0 : P.warm_off;
        for (long long t = t_lo; t < t_hi; ++t) {
            const long long tw = t + off;   // the window ends at data index tw (exclusive)
            // ---------------- frame tiers ----------------
            for (int i = 0; i < P.n_ft; ++i) {
                const SrTier& T = P.tiers[i];
                if (t % T.fs != 0) continue;
                const float* cond = nullptr;
                if (i > 0) cond = P.tiers[i - 1].obuf + (size_t)((t / T.fs) % T.kdiv) * H * Bp;
                // ---- n_rnn stacked cells (nn.GRU / nn.LSTM / nn.RNN, sample_rnn_v2.py:62-66, 92-96) on this CTA's hidden indices:
                //      layer 0 reads the tier input, layer k the new state of layer k - 1 (all rows: a grid barrier apart)
                float* hnext = nullptr;
                for (int k = 0; k < P.n_rnn; ++k) {
                    const float* hcur = T.hbuf[k] + (size_t)hsel[i] * H * Bp;
                    const float* below = hnext;                    // new hidden state of the layer below (k > 0)
                    hnext = T.hbuf[k] + (size_t)(hsel[i] ^ 1) * H * Bp;
                    const float* ccur = P.rnn_type == MMK_RNN_LSTM ? T.cbuf[k] + (size_t)hsel[i] * H * Bp : nullptr;
                    float* cnext = P.rnn_type == MMK_RNN_LSTM ? T.cbuf[k] + (size_t)(hsel[i] ^ 1) * H * Bp : nullptr;
                    const float* Wih = w_s + T.off_wih[k]; const float* bih = Wih + (size_t)H * P.NG;
                    const float* Whh = w_s + T.off_whh[k]; const float* bhh = Whh + (size_t)H * P.NG;
                    if (nj > 0) {
                        for (int pc = 0; pc < n_chunks; ++pc) {
                            const int b0 = pc * SR_PB;
                            if (k == 0) stage_frame_input(P, xs, lin_s, T.in_w, T.in_b, T.fs, tw, cond, b0);
                            else stage_rows(xs, below, H, Bp, b0);
                            const Tile tl(P.NG);
                            {
                                float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                                tl.accum(acc, Wih, P.NG, xs, H);
                                tl.store(acc, part);
                            }
                            __syncthreads();
                            for (int o = tid; o < P.G * P.JP * SR_PB; o += SR_NT) {
                                const int col = o / SR_PB, p = o - col * SR_PB;
                                gi_s[o] = Tile::reduce(part, P.NG, col, p) + bih[col];
                            }
                            __syncthreads();
                            stage_rows(xs, hcur, H, Bp, b0);
                            {
                                float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                                tl.accum(acc, Whh, P.NG, xs, H);
                                tl.store(acc, part);
                            }
                            __syncthreads();
                            for (int o = tid; o < nj * SR_PB; o += SR_NT) {
                                const int jj = o / SR_PB, p = o - jj * SR_PB;
                                float hnew;
                                if (P.rnn_type == MMK_RNN_GRU) {            // gates r, z, n
                                    const int cr = jj, cz = P.JP + jj, cn = 2 * P.JP + jj;
                                    const float hr = Tile::reduce(part, P.NG, cr, p) + bhh[cr];
                                    const float hz = Tile::reduce(part, P.NG, cz, p) + bhh[cz];
                                    const float hn = Tile::reduce(part, P.NG, cn, p) + bhh[cn];
                                    const float r = sigmoid_acc(gi_s[cr * SR_PB + p] + hr);
                                    const float zg = sigmoid_acc(gi_s[cz * SR_PB + p] + hz);
                                    const float n = tanhf(gi_s[cn * SR_PB + p] + r * hn);
                                    const float hold = xs[(j_lo + jj) * SR_PB + p];
                                    hnew = (1.0f - zg) * n + zg * hold;
                                } else if (P.rnn_type == MMK_RNN_LSTM) {    // gates i, f, g, o; c' = f c + i g; h' = o tanh(c')
                                    const int ci = jj, cf = P.JP + jj, cg = 2 * P.JP + jj, co = 3 * P.JP + jj;
                                    const float ai = gi_s[ci * SR_PB + p] + (Tile::reduce(part, P.NG, ci, p) + bhh[ci]);
                                    const float af = gi_s[cf * SR_PB + p] + (Tile::reduce(part, P.NG, cf, p) + bhh[cf]);
                                    const float ag = gi_s[cg * SR_PB + p] + (Tile::reduce(part, P.NG, cg, p) + bhh[cg]);
                                    const float ao = gi_s[co * SR_PB + p] + (Tile::reduce(part, P.NG, co, p) + bhh[co]);
                                    const float cold = (b0 + p < P.B) ? __ldcg(ccur + (size_t)(j_lo + jj) * Bp + b0 + p) : 0.0f;
                                    const float cnew = sigmoid_acc(af) * cold + sigmoid_acc(ai) * tanhf(ag);
                                    hnew = sigmoid_acc(ao) * tanhf(cnew);
                                    if (b0 + p < P.B) __stcg(cnext + (size_t)(j_lo + jj) * Bp + b0 + p, cnew);
                                } else {                                     // nn.RNN (tanh)
                                    hnew = tanhf(gi_s[jj * SR_PB + p] + (Tile::reduce(part, P.NG, jj, p) + bhh[jj]));
                                }
                                if (b0 + p < P.B) __stcg(hnext + (size_t)(j_lo + jj) * Bp + b0 + p, hnew);
                            }
                            __syncthreads();
                        }
                    }
                    grid_barrier(P, epoch);
                }
                hsel[i] ^= 1;
                // ---- LinearResampler rows of this CTA ----
                const int u_lo = part_lo(c, T.up_rows, NC), nu = part_lo(c + 1, T.up_rows, NC) - u_lo;
                if (nu > 0) {
                    const float* Wup = w_s + T.off_wup; const float* bup = Wup + (size_t)H * T.NU;
                    for (int pc = 0; pc < n_chunks; ++pc) {
                        const int b0 = pc * SR_PB;
                        stage_rows(xs, hnext, H, Bp, b0);
                        const Tile tl(T.NU);
                        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        tl.accum(acc, Wup, T.NU, xs, H);
                        tl.store(acc, part);
                        __syncthreads();
                        for (int o = tid; o < nu * SR_PB; o += SR_NT) {
                            const int col = o / SR_PB, p = o - col * SR_PB;
                            if (b0 + p < P.B)
                                __stcg(T.obuf + (size_t)(u_lo + col) * Bp + b0 + p, Tile::reduce(part, T.NU, col, p) + bup[col]);
                        }
                        __syncthreads();
                    }
                }
                grid_barrier(P, epoch);
            }
            if (!gen) continue;
            // ---------------- sample-level tier + head ----------------
            const SrTier& TL = P.tiers[P.n_ft - 1];
            if (nh1 > 0) {
                const float* cond = TL.obuf + (size_t)(t % TL.fs) * H * Bp;    // outputs[-1][:, (t % fs[-2]) - fs[-2]]
                const float* W1 = w_s + P.off_w1; const float* b1 = W1 + (size_t)H * P.NH1;
                for (int pc = 0; pc < n_chunks; ++pc) {
                    const int b0 = pc * SR_PB;
                    stage_frame_input(P, xs, lin_s, P.conv_w, P.conv_b, P.fs_last, tw, cond, b0);
                    const Tile tl(P.NH1);
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    tl.accum(acc, W1, P.NH1, xs, H);
                    tl.store(acc, part);
                    __syncthreads();
                    for (int o = tid; o < nh1 * SR_PB; o += SR_NT) {
                        const int col = o / SR_PB, p = o - col * SR_PB;
                        if (b0 + p < P.B)
                            __stcg(P.hid + (size_t)(h1_lo + col) * Bp + b0 + p,
                                   mish_acc(Tile::reduce(part, P.NH1, col, p) + b1[col]));
                    }
                    __syncthreads();
                }
            }
            grid_barrier(P, epoch);
            // MLP hidden layers (networks/mlp.py:47-50): the tuple repetition there makes it ONE Linear(Hh, Hh) + Mish applied
            // n_hidden_layers times; the buffers ping-pong
            const float* hid_in = P.hid;
            for (int r = 0; r < P.n_hh; ++r) {
                float* hid_out = P.hid + (size_t)((r + 1) & 1) * P.Hh * Bp;
                if (nh1 > 0) {
                    const float* Wh = w_s + P.off_wh; const float* bh = Wh + (size_t)P.Hh * P.NH1;
                    for (int pc = 0; pc < n_chunks; ++pc) {
                        const int b0 = pc * SR_PB;
                        stage_rows(xs, hid_in, P.Hh, Bp, b0);
                        const Tile tl(P.NH1);
                        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        tl.accum(acc, Wh, P.NH1, xs, P.Hh);
                        tl.store(acc, part);
                        __syncthreads();
                        for (int o = tid; o < nh1 * SR_PB; o += SR_NT) {
                            const int col = o / SR_PB, p = o - col * SR_PB;
                            if (b0 + p < P.B)
                                __stcg(hid_out + (size_t)(h1_lo + col) * Bp + b0 + p,
                                       mish_acc(Tile::reduce(part, P.NH1, col, p) + bh[col]));
                        }
                        __syncthreads();
                    }
                }
                hid_in = hid_out;
                grid_barrier(P, epoch);
            }
            if (nz > 0) {
                const float* W2 = w_s + P.off_w2; const float* b2 = W2 + (size_t)P.Hh * P.NZ;
                for (int pc = 0; pc < n_chunks; ++pc) {
                    const int b0 = pc * SR_PB;
                    stage_rows(xs, hid_in, P.Hh, Bp, b0);
                    const Tile tl(P.NZ);
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    tl.accum(acc, W2, P.NZ, xs, P.Hh);
                    tl.store(acc, part);
                    __syncthreads();
                    for (int o = tid; o < nz * SR_PB; o += SR_NT) {
                        const int col = o / SR_PB, p = o - col * SR_PB;
                        if (b0 + p < P.B)
                            __stcg(P.z + (size_t)(z_lo + col) * Bp + b0 + p, Tile::reduce(part, P.NZ, col, p) + b2[col]);
                    }
                    __syncthreads();
                }
            }
            grid_barrier(P, epoch);
            {   // one warp per prompt: learned temperature, argmax / inverse-CDF draw
                const int zrow = (P.Q + 1 + 3) / 4 * 4 + 4;
                const long long hstep = t - P.gen_begin, n_gen = P.gen_end - P.gen_begin;
                for (int b = c + NC * warp; b < P.B; b += NC * (SR_NT / 32)) {
                    float* zr = xs + warp * zrow;
                    for (int o = lane; o <= P.Q; o += 32) zr[o] = __ldcg(P.z + (size_t)o * Bp + b);
                    __syncwarp();
                    const bool sample = P.temperature != nullptr;
                    float Tt = 1.0f, u = 0.0f;
                    if (sample) {
                        Tt = P.temperature[P.n_temperature == 1 ? 0 : b];
                        u = P.noise[(size_t)b * P.noise_stride + (t - P.noise_t0)];
                    }
                    float* lout = P.logits_out ? P.logits_out + ((size_t)b * n_gen + hstep) * P.Q : nullptr;
                    const int choice = decide_warp(zr, P.Q, P.min_temp, lout, sample, Tt, u);
                    if (lane == 0) {
                        if (P.decisions) P.decisions[(size_t)b * n_gen + hstep] = choice;
                        if (!P.teacher_forced) __stcg(P.seq + (size_t)b * P.seq_stride + t, (long long)choice);
                    }
                    __syncwarp();
                }
            }
            grid_barrier(P, epoch);
            if (c == 0 && tid == 0 && P.step_ts) P.step_ts[t - P.gen_begin] = sr_globaltimer();
        }
    }
}

static int pad4(int v) { return (v + 3) / 4 * 4; }
static int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mmk

using namespace mmk;

struct mmk_samplernn_s {
    sr2_handle* v2 = nullptr;     // the cluster kernel (samplernn2.cu) when the configuration fits it
    SrParams p{};
    int device = 0, max_batch = 0, rf = 0;
    size_t smem_bytes = 0;
    std::vector<int> fs;
    std::vector<void*> allocs;
    size_t hbuf_floats = 0;       // floats of one [2][H][Bp] state buffer
};

static int sr_free(mmk_samplernn_s* h) {
    if (!h) return 0;
    if (h->v2) sr2_destroy(h->v2);
    for (void* a : h->allocs) cudaFree(a);
    delete h;
    return 0;
}

extern "C" int mmk_samplernn_create(const mmk_samplernn_desc* d, int max_batch, mmk_samplernn_t* out) {
    MMK_CHECK(d && out, "mmk_samplernn_create: null argument");
    mmk_samplernn_desc_ex ex{};
    ex.base = *d;
    ex.rnn_type = MMK_RNN_GRU; ex.n_rnn = 1;
    ex.w_ih = d->w_ih; ex.w_hh = d->w_hh; ex.b_ih = d->b_ih; ex.b_hh = d->b_hh;
    return mmk_samplernn_create_ex(&ex, max_batch, out);
}

extern "C" int mmk_samplernn_create_ex(const mmk_samplernn_desc_ex* dx, int max_batch, mmk_samplernn_t* out) {
    MMK_CHECK(dx && out, "mmk_samplernn_create_ex: null argument");
    const mmk_samplernn_desc* d = &dx->base;
    MMK_CHECK(dx->rnn_type == MMK_RNN_GRU || dx->rnn_type == MMK_RNN_LSTM || dx->rnn_type == MMK_RNN_TANH, "rnn_type must be MMK_RNN_GRU, _LSTM or _TANH");
    MMK_CHECK(dx->n_rnn >= 1 && dx->n_rnn <= SR_MAX_RNN, "n_rnn must be in [1, 4]");
    MMK_CHECK(dx->head_hidden_layers >= 0 && dx->head_hidden_layers <= 8, "head_hidden_layers must be in [0, 8]");
    MMK_CHECK(dx->head_hidden_layers == 0 || (dx->head_wh && dx->head_bh), "missing head hidden-layer weights");
    const int G = dx->rnn_type == MMK_RNN_GRU ? 3 : (dx->rnn_type == MMK_RNN_LSTM ? 4 : 1);
    const bool plain = dx->rnn_type == MMK_RNN_GRU && dx->n_rnn == 1 && dx->head_hidden_layers == 0 && !dx->need_set_hidden;
    MMK_CHECK(d->n_tiers >= 2 && d->n_tiers - 1 <= SR_MAX_TIERS, "n_tiers must be in [2, 7]");
    MMK_CHECK(d->hidden_dim >= 4 && d->hidden_dim % 4 == 0, "hidden_dim must be a positive multiple of 4");
    MMK_CHECK(d->head_hidden >= 4 && d->head_hidden % 4 == 0, "head_hidden must be a positive multiple of 4");
    MMK_CHECK(d->q_levels >= 2 && d->q_levels <= 1024, "q_levels must be in [2, 1024]");
    MMK_CHECK(max_batch >= 1, "max_batch must be >= 1");
    MMK_CHECK(d->frame_sizes && d->in_w && d->in_b && dx->w_ih && dx->w_hh && dx->b_ih && dx->b_hh && d->up_w && d->up_b &&
              d->conv_w && d->conv_b && d->head_w1 && d->head_b1 && d->head_w2 && d->head_b2, "missing weight pointers");
    int ndev = 0;
    MMK_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device: mmk_b200 has no CPU fallback");
    const int n_ft = d->n_tiers - 1, H = d->hidden_dim, Hh = d->head_hidden, Q = d->q_levels;
    for (int i = 0; i < d->n_tiers; ++i) MMK_CHECK(d->frame_sizes[i] >= 1 && d->frame_sizes[i] <= 64, "frame sizes must be in [1, 64]");
    for (int i = 1; i < n_ft; ++i)
        MMK_CHECK(d->frame_sizes[i - 1] % d->frame_sizes[i] == 0, "frame sizes must divide each other");

    auto* h = new mmk_samplernn_s();
    struct Guard {                 // every early return below (MMK_CHECK / MMK_CUDA / MMK_FAIL) releases the handle and what it owns
        mmk_samplernn_s* h;
        ~Guard() { }
    } guard{h};
    SrParams& p = h->p;
    MMK_CUDA(cudaGetDevice(&h->device));
    {   // MMK_SR_KERNEL: "1" = general kernel only, "2" = cluster kernel only, unset = cluster kernel when it fits
        const char* force = getenv("MMK_SR_KERNEL");
        const int tc = dx->compute_mode == MMK_COMPUTE_BF16_TC ? 1 : 0;
        // the tensor-core engine also hosts nn.LSTM tiers (the reference's default rnn_class); one layer, zero initial state, plain head
        const bool tc_form = (dx->rnn_type == MMK_RNN_GRU || dx->rnn_type == MMK_RNN_LSTM) && dx->n_rnn == 1 && dx->head_hidden_layers == 0;
        if (tc && !tc_form) { MMK_FAIL("the bf16 tensor-core mode hosts GRU / LSTM tiers with one layer and a plain head"); }
        // ... and so does the lane-major fp32 engine (sequences bit-exact): GRU or LSTM, tried first; what it cannot host falls to the
        // general kernel below
        if ((plain || tc_form) && (tc || !force || atoi(force) != 1)) {
            int unsupported = 0;
            mmk_samplernn_desc d2 = *d;      // the cluster kernel hosts the GRU / one layer / plain head form only
            d2.w_ih = dx->w_ih; d2.w_hh = dx->w_hh; d2.b_ih = dx->b_ih; d2.b_hh = dx->b_hh;
            if (sr2_create(&d2, max_batch, tc, dx->rnn_type == MMK_RNN_LSTM ? 1 : 0, dx->need_set_hidden ? 1 : 0, &h->v2, &unsupported) == 0) {
                h->max_batch = max_batch; h->rf = d->frame_sizes[0];
                guard.h = nullptr;
                *out = h;
                return 0;
            }
            h->v2 = nullptr;
            if (!unsupported) { return 1; }
            if (tc) { MMK_FAIL("the bf16 tensor-core mode needs hidden_dim in {128, 256, 512}, frame sizes / up-sampling factors in {1,2,4,8} / {1,2,4} and max_batch <= 128"); }
            if (force && atoi(force) == 2) { MMK_FAIL("configuration not supported by the cluster kernel (MMK_SR_KERNEL=2)"); }
        }
    }
    int sms = 0, max_optin = 0, coop = 0;
    MMK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    MMK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    MMK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) { MMK_FAIL("device does not support cooperative launches"); }
    // as many CTAs as SMs, trimmed so that the hidden indices split evenly (no padded GRU columns)
    int NC = std::min(sms, H);
    NC = std::min(NC, H / ceil_div(H, NC));
    if (const char* e = getenv("MMK_SR_CTAS")) NC = std::max(1, std::min(std::min(sms, H), atoi(e)));
    p.n_ft = n_ft; p.H = H; p.Hh = Hh; p.Q = Q; p.NC = NC;
    p.rnn_type = dx->rnn_type; p.G = G; p.n_rnn = dx->n_rnn; p.n_hh = dx->head_hidden_layers;
    p.JP = ceil_div(H, NC); p.NG = pad4(G * p.JP);
    p.NH1 = pad4(ceil_div(Hh, NC)); p.NZ = pad4(ceil_div(Q + 1, NC));
    p.fs_last = d->frame_sizes[n_ft];
    p.min_temp = d->min_temperature;
    h->max_batch = max_batch; h->rf = d->frame_sizes[0];
    h->fs.assign(d->frame_sizes, d->frame_sizes + d->n_tiers);
    p.Bp = ceil_div(max_batch, SR_PB) * SR_PB;

    int o = 0, fs_max = p.fs_last, widest = std::max(p.NG, std::max(p.NH1, p.NZ));
    auto take = [&](int floats) { int r = o; o += pad4(floats); return r; };
    for (int i = 0; i < n_ft; ++i) {
        SrTier& T = p.tiers[i];
        T.fs = d->frame_sizes[i];
        T.up = T.fs / (i < n_ft - 1 ? d->frame_sizes[i + 1] : 1);     // sample_rnn_v2.py:155-158
        T.kdiv = i > 0 ? d->frame_sizes[i - 1] / T.fs : 1;
        T.up_rows = T.up * H;
        T.NU = pad4(ceil_div(T.up_rows, NC));
        for (int k = 0; k < p.n_rnn; ++k) {
            T.off_wih[k] = take(H * p.NG + p.NG);
            T.off_whh[k] = take(H * p.NG + p.NG);
        }
        T.off_wup = take(H * T.NU + T.NU);
        fs_max = std::max(fs_max, T.fs);
        widest = std::max(widest, T.NU);
    }
    p.off_w1 = take(H * p.NH1 + p.NH1);
    p.off_w2 = take(Hh * p.NZ + p.NZ);
    p.off_wh = p.n_hh > 0 ? take(Hh * p.NH1 + p.NH1) : 0;
    p.cta_block = o;
    const int zrow = pad4(Q + 1) + 4;
    p.off_x = take(std::max(std::max(H, Hh) * SR_PB, (SR_NT / 32) * zrow));
    p.off_part = take(SR_NT * 8);
    p.off_gi = take(p.NG * SR_PB);
    p.off_lin = take(SR_PB * fs_max);
    p.smem_floats = o;
    h->smem_bytes = (size_t)o * sizeof(float);
    if ((SR_PB / 2) * (widest / 4) > SR_NT) { MMK_FAIL("SampleRNN rows per CTA too wide for one contraction pass"); }
    if (h->smem_bytes > (size_t)max_optin) { MMK_FAIL("SampleRNN configuration does not fit in shared memory (weights are kept resident)"); }
    MMK_CUDA(cudaFuncSetAttribute(samplernn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    int per_sm = 0;
    MMK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, samplernn_kernel, SR_NT, h->smem_bytes));
    if (per_sm * sms < NC) { MMK_FAIL("SampleRNN grid cannot be co-resident"); }

    // ---- pack per-CTA weight blocks
    std::vector<float> wpack((size_t)NC * p.cta_block, 0.0f);
    for (int c = 0; c < NC; ++c) {
        float* blk = wpack.data() + (size_t)c * p.cta_block;
        const int j_lo = part_lo(c, H, NC), nj = part_lo(c + 1, H, NC) - j_lo;
        for (int i = 0; i < n_ft; ++i) {
            const SrTier& T = p.tiers[i];
            for (int lay = 0; lay < p.n_rnn; ++lay)
                for (int m = 0; m < 2; ++m) {
                    const int e = i * p.n_rnn + lay;
                    const float* W = m == 0 ? dx->w_ih[e] : dx->w_hh[e];
                    const float* b = m == 0 ? dx->b_ih[e] : dx->b_hh[e];
                    float* Ws = blk + (m == 0 ? T.off_wih[lay] : T.off_whh[lay]);
                    float* bs = Ws + (size_t)H * p.NG;
                    for (int g = 0; g < G; ++g)
                        for (int jj = 0; jj < nj; ++jj) {
                            const int col = g * p.JP + jj, row = g * H + j_lo + jj;
                            for (int k = 0; k < H; ++k) Ws[(size_t)k * p.NG + col] = W[(size_t)row * H + k];
                            bs[col] = b[row];
                        }
                }
            const int u_lo = part_lo(c, T.up_rows, NC), nu = part_lo(c + 1, T.up_rows, NC) - u_lo;
            float* Wu = blk + T.off_wup; float* bu = Wu + (size_t)H * T.NU;
            for (int col = 0; col < nu; ++col) {
                for (int k = 0; k < H; ++k) Wu[(size_t)k * T.NU + col] = d->up_w[i][(size_t)(u_lo + col) * H + k];
                bu[col] = d->up_b[i][u_lo + col];
            }
        }
        const int h1_lo = part_lo(c, Hh, NC), nh1 = part_lo(c + 1, Hh, NC) - h1_lo;
        float* W1 = blk + p.off_w1; float* b1 = W1 + (size_t)H * p.NH1;
        for (int col = 0; col < nh1; ++col) {
            for (int k = 0; k < H; ++k) W1[(size_t)k * p.NH1 + col] = d->head_w1[(size_t)(h1_lo + col) * H + k];
            b1[col] = d->head_b1[h1_lo + col];
        }
        const int z_lo = part_lo(c, Q + 1, NC), nz = part_lo(c + 1, Q + 1, NC) - z_lo;
        float* W2 = blk + p.off_w2; float* b2 = W2 + (size_t)Hh * p.NZ;
        for (int col = 0; col < nz; ++col) {
            for (int k = 0; k < Hh; ++k) W2[(size_t)k * p.NZ + col] = d->head_w2[(size_t)(z_lo + col) * Hh + k];
            b2[col] = d->head_b2[z_lo + col];
        }
        if (p.n_hh > 0) {
            float* Wh = blk + p.off_wh; float* bh = Wh + (size_t)Hh * p.NH1;
            for (int col = 0; col < nh1; ++col) {
                for (int k = 0; k < Hh; ++k) Wh[(size_t)k * p.NH1 + col] = dx->head_wh[(size_t)(h1_lo + col) * Hh + k];
                bh[col] = dx->head_bh[h1_lo + col];
            }
        }
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
        SrTier& T = p.tiers[i];
        T.in_w = up(d->in_w[i], (size_t)H * T.fs);
        T.in_b = up(d->in_b[i], H);
        h->hbuf_floats = (size_t)2 * H * p.Bp;
        for (int k = 0; k < p.n_rnn; ++k) {
            T.hbuf[k] = up(nullptr, h->hbuf_floats);
            T.cbuf[k] = p.rnn_type == MMK_RNN_LSTM ? up(nullptr, h->hbuf_floats) : nullptr;
        }
        T.obuf = up(nullptr, (size_t)T.up_rows * p.Bp);
    }
    p.conv_w = up(d->conv_w, (size_t)H * p.fs_last);
    p.conv_b = up(d->conv_b, H);
    p.hid = up(nullptr, (size_t)2 * Hh * p.Bp);
    p.z = up(nullptr, (size_t)(Q + 1) * p.Bp);
    p.bar = (unsigned long long*)dev_alloc(64, nullptr);
    ok = ok && p.bar;
    if (!ok) { MMK_FAIL("cudaMalloc failed while creating the SampleRNN handle"); }
    p.abort_flag = (unsigned*)(p.bar + 1);
    MMK_CUDA(cudaDeviceSynchronize());
    guard.h = nullptr;
    *out = h;
    return 0;
}

extern "C" int mmk_samplernn_destroy(mmk_samplernn_t h) { return sr_free(h); }

extern "C" int mmk_samplernn_launch_info(mmk_samplernn_t h, mmk_launch_info* out) {
    MMK_CHECK(h && out, "null argument");
    if (h->v2) return sr2_launch_info(h->v2, out);
    out->cluster_size = 1; out->n_stages = h->p.NC; out->group_size = SR_PB; out->threads = SR_NT;
    out->smem_bytes = (int)h->smem_bytes; out->sm_used = h->p.NC;
    return 0;
}

extern "C" int mmk_samplernn_run(mmk_samplernn_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0,
                                 int64_t warm_begin, int64_t warm_end, int64_t warm_offset, int64_t gen_begin,
                                 int64_t gen_end, int reset_hidden, int teacher_forced, const float* d_temperature,
                                 int n_temperature, const float* d_noise, int64_t noise_stride, int64_t noise_t0,
                                 float* d_logits_out, int64_t* d_decisions, unsigned long long* d_step_ts, void* stream) {
    MMK_CHECK(h && d_seq, "mmk_samplernn_run: null argument");
    MMK_CHECK(B >= 1 && B <= h->max_batch, "batch exceeds the max_batch the handle was created for");
    MMK_CHECK(d_temperature == nullptr || (n_temperature == 1 || n_temperature == B), "temperature must have 1 or B entries");
    MMK_CHECK(d_temperature == nullptr || d_noise != nullptr, "sampling (temperature given) needs a noise tensor");
    const int rf = h->rf;
    if (warm_end < warm_begin) warm_end = warm_begin;
    if (gen_end < gen_begin) gen_end = gen_begin;
    if (warm_end > warm_begin)
        MMK_CHECK(warm_begin + warm_offset - rf >= seq_t0 && warm_end - 1 + warm_offset - seq_t0 <= seq_stride,
                  "warm-up range reads outside the sequence buffer");
    if (gen_end > gen_begin)
        MMK_CHECK(gen_begin - rf >= seq_t0 && gen_end - seq_t0 <= seq_stride, "generate range falls outside the sequence buffer");
    if (h->v2)
        return sr2_run(h->v2, d_seq, B, seq_stride, seq_t0, warm_begin, warm_end, warm_offset, gen_begin, gen_end,
                       reset_hidden, teacher_forced, d_temperature, n_temperature, d_noise, noise_stride, noise_t0,
                       d_logits_out, d_decisions, d_step_ts, stream);
    cudaStream_t st = (cudaStream_t)stream;
    SrParams p = h->p;
    if (reset_hidden) {
        for (int i = 0; i < p.n_ft; ++i) {
            for (int k = 0; k < p.n_rnn; ++k) {
                MMK_CUDA(cudaMemsetAsync(p.tiers[i].hbuf[k], 0, h->hbuf_floats * sizeof(float), st));
                if (p.tiers[i].cbuf[k]) MMK_CUDA(cudaMemsetAsync(p.tiers[i].cbuf[k], 0, h->hbuf_floats * sizeof(float), st));
            }
            h->p.hsel[i] = 0;
        }
        p = h->p;
    }
    if (warm_end == warm_begin && gen_end == gen_begin) return 0;
    MMK_CUDA(cudaMemsetAsync(p.bar, 0, 64, st));
    p.seq = reinterpret_cast<long long*>(d_seq) - seq_t0;
    p.seq_stride = seq_stride;
    p.warm_begin = warm_begin; p.warm_end = warm_end; p.warm_off = warm_offset;
    p.gen_begin = gen_begin; p.gen_end = gen_end;
    p.B = B; p.teacher_forced = teacher_forced ? 1 : 0;
    p.temperature = d_temperature; p.n_temperature = n_temperature;
    p.noise = d_noise; p.noise_stride = noise_stride; p.noise_t0 = noise_t0;
    p.logits_out = d_logits_out; p.decisions = reinterpret_cast<long long*>(d_decisions); p.step_ts = d_step_ts;
    // the prompt chunks cover only the live batch
    p.Bp = h->p.Bp;
    void* args[] = {&p};
    MMK_CUDA(cudaLaunchCooperativeKernel((const void*)samplernn_kernel, dim3(p.NC), dim3(SR_NT), args, h->smem_bytes, st));
    // the hidden ping-pong advances once per tier firing: keep the handle's view in step with the device
    for (int i = 0; i < p.n_ft; ++i) {
        const long long fs = p.tiers[i].fs;
        auto firings = [&](long long lo, long long hi) { return hi > lo ? (hi + fs - 1) / fs - (lo + fs - 1) / fs : 0; };
        const long long n = firings(warm_begin, warm_end) + firings(gen_begin, gen_end);
        h->p.hsel[i] ^= (int)(n & 1);
    }
    return 0;
}

namespace mmk {
__global__ void sr_set_hidden_kernel(float* dst, const float* src, int B, int H, int Bp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // dst [H][Bp] <- src (B, H)
    if (i < B * H) { const int b = i / H, j = i - b * H; dst[(size_t)j * Bp + b] = src[i]; }
}
}  // namespace mmk

extern "C" int mmk_samplernn_set_hidden(mmk_samplernn_t h, int tier, int layer, int which, const float* d_values, int B,
                                        void* stream) {
    MMK_CHECK(h && d_values, "mmk_samplernn_set_hidden: null argument");
    if (h->v2) {
        MMK_CHECK(layer == 0, "layer out of range");
        return sr2_set_hidden(h->v2, tier, which, d_values, B, stream);
    }
    const SrParams& p = h->p;
    MMK_CHECK(tier >= 0 && tier < p.n_ft && layer >= 0 && layer < p.n_rnn, "tier / layer out of range");
    MMK_CHECK(which == 0 || (which == 1 && p.rnn_type == MMK_RNN_LSTM), "which: 0 = hidden state, 1 = LSTM cell state");
    MMK_CHECK(B >= 1 && B <= h->max_batch, "batch exceeds the max_batch the handle was created for");
    float* base = which == 0 ? p.tiers[tier].hbuf[layer] : p.tiers[tier].cbuf[layer];
    float* dst = base + (size_t)p.hsel[tier] * p.H * p.Bp;
    sr_set_hidden_kernel<<<(B * p.H + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dst, d_values, B, p.H, p.Bp);
    MMK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmk_samplernn_sync_check(mmk_samplernn_t h, void* stream) {
    MMK_CHECK(h, "null handle");
    if (h->v2) return sr2_sync_check(h->v2, stream);
    unsigned aborted = 0;
    MMK_CUDA(cudaMemcpyAsync(&aborted, h->p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MMK_CHECK(aborted == 0, "SampleRNN kernel watchdog fired: a grid barrier timed out (results invalid)");
    return 0;
}

extern "C" int mmk_samplernn_generate(mmk_samplernn_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t prompt_len,
                                      int64_t n_steps, const float* d_temperature, int n_temperature,
                                      const float* d_noise, float* d_logits_out, unsigned long long* d_step_ts,
                                      void* stream) {
    MMK_CHECK(h, "null handle");
    MMK_CHECK(prompt_len >= h->rf, "prompt shorter than the top frame size");
    MMK_CHECK(n_steps >= 0 && prompt_len + n_steps <= seq_stride, "sequence buffer too short");
    const int64_t offset = prompt_len % h->rf;
    return mmk_samplernn_run(h, d_seq, B, seq_stride, 0, h->rf, prompt_len - offset, offset, prompt_len,
                             prompt_len + n_steps, 1, 0, d_temperature, n_temperature, d_noise, n_steps, prompt_len,
                             d_logits_out, nullptr, d_step_ts, stream);
}
