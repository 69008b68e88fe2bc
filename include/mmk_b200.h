/* mmk_b200 — C ABI of the B200-native generation + feature hot path (drop-in for ktonal/mimikit's path).
 *
 * The reference (mimikit 0.4.3) is pure Python and has NO FFI; each entry point below names the reference
 * interface (file:line under /root/reference) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes
 * binding a mimikit maintainer would add.  Conventions:
 *   - every function returns 0 on success, non-zero on error; mmk_last_error() returns the message of the last
 *     failing call on the calling thread.  There is no CPU fallback: without a CUDA device every compute call fails.
 *   - pointers named d_* are DEVICE pointers owned by the caller (torch); h_* / desc fields are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are stream-ordered and
 *     asynchronous w.r.t. the host unless stated otherwise.  One host thread per handle; a handle is bound to the
 *     device that was current when it was created.
 */
#ifndef MMK_B200_H
#define MMK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMK_ABI_VERSION 1

int mmk_abi_version(void);
const char* mmk_last_error(void);

/* Device facts the host side sizes launches with (SM count, max co-resident clusters ...). */
typedef struct {
    int sm_count;
    int cc_major, cc_minor;
    int max_smem_optin;      /* bytes per CTA */
    int l2_bytes;
} mmk_device_info;
int mmk_get_device_info(mmk_device_info* out);

/* ------------------------------------------------------------------------------------------------
 * Features
 * ---------------------------------------------------------------------------------------------- */

/* MuLawCompress.torch_func — mimikit/features/functionals.py:330-338.
 * d_x fp32 (n) -> d_q int64 (n).  Bit-exact with the reference evaluated by torch on CPU. */
int mmk_mulaw_compress(const float* d_x, int64_t* d_q, size_t n, int q_levels, float compression, void* stream);
/* Same arithmetic, uint8 output (q_levels <= 256): the compact form used inside the generation path. */
int mmk_mulaw_compress_u8(const float* d_x, uint8_t* d_q, size_t n, int q_levels, float compression, void* stream);

/* The default compress/expand kernels use a threshold table that is built with the exact arithmetic and then PROVEN
 * equal to it on every float in [-1, 1] on the device, once per (device, q_levels, compression) — on the first call
 * that needs it (one ~20 ms launch + a stream synchronisation; not possible inside a stream capture, where the exact
 * kernel runs instead).  mmk_mulaw_prepare does that step explicitly.  *table_in_use = 1 if the proven table kernels
 * will serve this parameter pair, 0 if the exact-arithmetic kernels will (proof failed: *mismatches > 0; q_levels >
 * 2048; or MMK_MULAW_EXACT=1 in the environment).  Both pointers are nullable HOST pointers. */
int mmk_mulaw_prepare(int q_levels, float compression, int* table_in_use, uint64_t* mismatches, void* stream);

/* MuLawExpand.torch_func — mimikit/features/functionals.py:361-369 (the `feature.inv` that
 * GenerateLoopV2.process_outputs applies, mimikit/loops/generate.py:245).  d_q int64 (n) -> d_x fp32 (n). */
int mmk_mulaw_expand(const int64_t* d_q, float* d_x, size_t n, int q_levels, float compression, void* stream);

/* Normalize(p=inf, dim=-1).torch_func — mimikit/features/functionals.py:236-253: F.normalize(x, p=inf, dim=-1) =
 * x / max(max|x|, 1e-12) per row (one IEEE division per sample; NaN rows stay NaN).  d_x fp32 (n_rows, row_len) with row
 * stride `row_stride` elements; d_out fp32 (n_rows, row_len) contiguous (may alias a contiguous d_x); d_norms fp32 (n_rows)
 * receives the row maxima (it is also the scratch of the first pass).  n_rows <= 65535 per call. */
int mmk_normalize_inf(const float* d_x, float* d_out, float* d_norms, int64_t n_rows, int64_t row_len, int64_t row_stride,
                      void* stream);
/* Compose(Normalize(), MuLawCompress(q_levels, compression)) — functionals.py:196-213, 236-253, 313-342 — in two passes
 * over the waveform (16 B per sample) instead of four (24 B): row maxima, then x / max -> mu-law level.  d_q int64
 * (n_rows, row_len) contiguous.  Bit-exact with the reference's composition evaluated by torch on CPU. */
int mmk_normalize_mulaw_compress(const float* d_x, int64_t* d_q, float* d_norms, int64_t n_rows, int64_t row_len,
                                 int64_t row_stride, int q_levels, float compression, void* stream);

/* RemoveDC.np_func — mimikit/features/functionals.py:216-233: scipy.signal.lfilter([1, -1], [1, -0.99], x, axis=-1) in
 * float64 (scipy's direct-form-II-transposed loop, zero initial state), cast back to float32.  ALWAYS bit-exact with that
 * chain.  Without scratch one lane per row walks the whole row (parallel over rows only).  With d_scratch of at least
 * mmk_remove_dc_scratch_bytes(n_rows, row_len) bytes (8-byte aligned) rows are also split in time: every chunk is filtered
 * speculatively after a warm-up from a zero state, every seam is checked for bitwise equality of the filter state, and
 * rows with a failed seam are recomputed sequentially on the same stream — same bits, ~12x faster for few long rows.
 * d_x fp32 (n_rows, row_len) with row stride; d_out fp32 (n_rows, row_len) contiguous, must not alias d_x.
 * (RemoveDC.torch_func cannot run in the reference: it passes lfilter's arguments in the wrong order.) */
size_t mmk_remove_dc_scratch_bytes(int64_t n_rows, int64_t row_len);
int mmk_remove_dc(const float* d_x, float* d_out, int64_t n_rows, int64_t row_len, int64_t row_stride, void* d_scratch,
                  size_t scratch_bytes, void* stream);
/* Diagnostic: rows the last speculative mmk_remove_dc call on this scratch recomputed sequentially (synchronises). */
int mmk_remove_dc_recomputed_rows(const void* d_scratch, int64_t n_rows, int64_t row_len, int64_t* h_count, void* stream);

/* MagSpec.torch_func -> STFT(coordinate="mag").torch_func — mimikit/features/functionals.py:468-524, 576-606,
 * fused with MelSpec.np_func — functionals.py:649-668 (librosa mel filterbank @ magnitudes).
 *   d_x        fp32 (n_clips, clip_len), row stride `clip_stride` elements
 *   n_fft      power of two in [64, 4096]; hop > 0; center: 0/1 (zero "constant" padding of n_fft/2 per side)
 *   align      0 = none, 1 = "end" (keep the LAST target samples), 2 = "start"   (STFT._fix_length)
 *   d_mag_out  nullable, fp32 (n_clips, n_frames, n_fft/2+1)
 *   d_mel_fb   nullable, fp32 (n_mels, n_fft/2+1) dense filterbank (mmk_mel_filterbank builds the reference's)
 *   d_mel_out  nullable, fp32 (n_clips, n_frames, n_mels)
 * n_frames follows the reference: see mmk_stft_n_frames. */
int mmk_stft_mag_mel(const float* d_x, int n_clips, int64_t clip_len, int64_t clip_stride, int n_fft, int hop,
                     int center, int align, float* d_mag_out, const float* d_mel_fb, int n_mels, float* d_mel_out,
                     void* stream);
/* Frame count and kept length after STFT._fix_length (functionals.py:468-486, item_spec.py:58-98). Host-only. */
int mmk_stft_n_frames(int64_t clip_len, int n_fft, int hop, int center, int align, int64_t* n_frames,
                      int64_t* kept_len);
/* librosa.filters.mel(sr=22050, n_fft, n_mels, fmin, fmax, htk, norm="slaney") as reached from
 * functionals.py:665-668 (the reference never passes sr).  Host-only: writes (n_mels, n_fft/2+1) fp32 to h_out.
 * fmax <= 0 means sr/2. */
int mmk_mel_filterbank(int n_fft, int n_mels, float fmin, float fmax, int htk, float* h_out);

/* MelSpec.np_func on PRECOMPUTED magnitudes — functionals.py:665-668: librosa.feature.melspectrogram(S=inputs.T, ...).T =
 * inputs @ mel_basis^T.  d_mag fp32 (n_frames, n_bins) with row stride `mag_stride` elements (n_bins = n_fft/2+1),
 * d_mel_fb fp32 (n_mels, n_bins) dense filterbank on the device (mmk_mel_filterbank builds the reference's on the host),
 * d_mel_out fp32 (n_frames, n_mels) contiguous.  One warp per frame, sparse filter supports, HBM-bound. */
int mmk_mel_apply(const float* d_mag, int64_t n_frames, int n_bins, int64_t mag_stride, const float* d_mel_fb, int n_mels,
                  float* d_mel_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * WaveNet — WaveNet.generate_step / forward (mimikit/networks/wavenet_v2.py:447-452, 276-293), WNLayer.forward
 * (131-176), EmbeddingIO (modules/io.py:148-154), MLP head (networks/mlp.py:44-63), CategoricalSampler
 * (modules/targets.py:40-52), driven as GenerateLoopV2.run does (loops/generate.py:184-229).
 * Supported: single mu-law embedding input, kernel_size 2, gated tanh/sigmoid, pad_side 0, optional residual and
 * skip 1x1 convs, MLP head with n_hidden_layers = 0 and a learned temperature.  Anything else: error.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mmk_wavenet_s* mmk_wavenet_t;

typedef struct {
    int n_layers;
    int dilated_dim;              /* C: dims_dilated[0] (== residuals_dim when residual convs exist) */
    int skips_dim;                /* S, 0 = no skip convs */
    int head_hidden;              /* MLPIO.hidden_dim */
    int q_levels;                 /* Q; head emits Q+1 values (learned temperature channel) */
    float min_temperature;        /* MLP.min_temp buffer.  A head built with min_temperature=None (mlp.py:29, 54-62: Q outputs, no
                                   * division) is passed as Q + 1 rows whose last row is zero with bias 40 and min_temperature 0:
                                   * sigmoid(40) rounds to 1.0f, so the division is exact (mimikit_b200/arm.py:_head_last) */
    const int* dilations;         /* [n_layers] */
    /* HOST pointers, fp32, in the reference state_dict layouts (key names in comments) */
    const float* embedding;                /* input_modules.0.0.weight        (Q, C)     */
    const float* const* conv_dil_w;        /* layers.l.conv_dil.0.0.weight    (2C, C, 2) */
    const float* const* conv_dil_b;        /* layers.l.conv_dil.0.0.bias      (2C)       */
    const float* const* conv_skip_w;       /* layers.l.conv_skip.weight       (S, C, 1)  or NULL array when S == 0 */
    const float* const* conv_skip_b;       /* layers.l.conv_skip.bias         (S)        */
    const float* const* conv_res_w;        /* layers.l.conv_res.weight        (C, C, 1)  entry NULL = layer has none; on the last
                                            * layer (reverse_layer_order, blocks=()) only the fp32 general kernel takes one */
    const float* const* conv_res_b;        /* layers.l.conv_res.bias          (C)        */
    const float* head_w1;                  /* output_modules.0.estimator.0.fc.0.weight (Hh, S|C) */
    const float* head_b1;                  /* ....fc.0.bias   (Hh)      */
    const float* head_w2;                  /* ....fc.2.weight (Q+1, Hh) */
    const float* head_b2;                  /* ....fc.2.bias   (Q+1)     */
} mmk_wavenet_desc;

/* Repacks the weights for the device and sizes the persistent kernel for up to max_batch prompts. */
int mmk_wavenet_create(const mmk_wavenet_desc* desc, int max_batch, mmk_wavenet_t* out);
/* Same with an explicit arithmetic:
 *   MMK_COMPUTE_FP32     fp32 FFMA kernels (what mmk_wavenet_create builds): argmax / noise-sampled sequences are
 *                        bit-exact with the oracle, logits within 1e-3 relative.
 *   MMK_COMPUTE_BF16_TC  bf16 operands, fp32 accumulation on the tensor cores (tcgen05.mma, accumulators in TMEM, weights
 *                        streamed by cp.async.bulk): the batch is the MMA M dimension, one CTA per 128 prompts, for the
 *                        large-batch regime (BASELINE.json configs[3]).  Logits within 5e-2 relative (the north star's
 *                        bf16 tolerance); sequences follow the bf16 logits.  Needs residual and skip convs, channel
 *                        counts of 64 or 128 and q_levels a multiple of 64 up to 256; any other
 *                        configuration is an error (no silent change of precision). */
#define MMK_COMPUTE_FP32 0
#define MMK_COMPUTE_BF16_TC 1
int mmk_wavenet_create_ex(const mmk_wavenet_desc* desc, int max_batch, int compute_mode, mmk_wavenet_t* out);
/* The rest of the WaveNet configuration surface that changes only ring depth, tap count and one add:
 * per-layer kernel sizes (wavenet_v2.py:295-327; conv_dil_w[l] is then (2C, C, k_l)), layerwise_inputs (:283-284),
 * n_hidden_layers > 0 in the MLP head (networks/mlp.py:47-50: one shared Linear) and with_affine_residuals.  These run in the general fp32 kernel;
 * a desc with kernel sizes all 2, no layerwise inputs and a plain head takes the same route as mmk_wavenet_create_ex.
 * (pad_side = 1 needs nothing here: the generation loop evaluates the last position of an rf-long window, where the
 * padded and the unpadded network agree.) */
#define MMK_ACT_DEFAULT 0
#define MMK_ACT_TANH 1
#define MMK_ACT_SIGMOID 2
#define MMK_ACT_MISH 3
#define MMK_ACT_RELU 4
#define MMK_ACT_SOFTPLUS 5     /* beta = 1, threshold = 20 (torch.nn.Softplus defaults) */
#define MMK_ACT_IDENTITY 6
#define MMK_ACT_ABS 7
#define MMK_ACT_SIN 8
#define MMK_ACT_COS 9
typedef struct {
    mmk_wavenet_desc base;
    const int* kernel_sizes;      /* [n_layers], each 2..4; NULL = all 2 */
    int layerwise_inputs;         /* 0 / 1 */
    int head_hidden_layers;       /* MLP n_hidden_layers; base.head_w2 is then the LAST Linear (fc.{2 + 2 n}) */
    const float* head_wh;         /* output_modules.0.estimator.0.fc.2.weight (Hh, Hh) when head_hidden_layers > 0 */
    const float* head_bh;         /* ...fc.2.bias (Hh) */
    /* with_affine_residuals (wavenet_v2.py:121-122, 148-149, 164-165; networks/parametrized.py:34-47): every layer's input
     * first goes through aff_res, z = x_hat * a + b with (x_hat | a | b) the three C-row chunks of one 1x1 conv; the dilated
     * conv and the residual add read z.  NULL arrays = the network has none; otherwise one entry per layer. */
    const float* const* aff_res_w;   /* layers.l.aff_res.params.weight (3C, C, 1) */
    const float* const* aff_res_b;   /* layers.l.aff_res.params.bias   (3C)       */
    /* act_f / act_g of the gated unit y = act_f(a_f) * act_g(a_g) (wavenet_v2.py:198-199, 224-225, 151): the point-wise members
     * of ActivationEnum (modules/activations.py:26-40).  0 = the default (Tanh filter, Sigmoid gate). */
    int act_f;                       /* MMK_ACT_* */
    int act_g;                       /* MMK_ACT_* */
} mmk_wavenet_desc_ex;
int mmk_wavenet_create_cfg(const mmk_wavenet_desc_ex* desc, int max_batch, int compute_mode, mmk_wavenet_t* out);
/* Diagnostic for the tensor-core path: d_D (128, N) fp32 = d_A (128, K) . d_B (N, K)^T with operands rounded to bf16,
 * through the same shared-memory descriptors (K-major SWIZZLE_128B), tcgen05.mma and TMEM loads as the bf16 kernel.
 * N a multiple of 16 <= 256, K a multiple of 64 <= 256.  h_cycles: nullable HOST pointer; receives the clock cycles of 8
 * back-to-back passes of the K / 16 instructions (the call then synchronises the stream). */
int mmk_tc_gemm_check(const float* d_A, const float* d_B, float* d_D, int N, int K, long long* h_cycles, void* stream);
int mmk_wavenet_destroy(mmk_wavenet_t h);
int mmk_wavenet_rf(mmk_wavenet_t h);   /* WaveNet.rf, wavenet_v2.py:337-339 */

/* Blocks until the handle's stream work is done and reports whether the last launch hit its watchdog (a spin on an
 * inter-stage flag timed out: results are invalid).  Returns non-zero with a message in that case. */
int mmk_wavenet_sync_check(mmk_wavenet_t h, void* stream);

/* Launch geometry chosen at create time, for reporting. */
typedef struct {
    int cluster_size, n_stages, group_size, threads, smem_bytes, sm_used;
} mmk_launch_info;
int mmk_wavenet_launch_info(mmk_wavenet_t h, mmk_launch_info* out);

/* One call = one persistent-kernel launch covering time indices [t_begin, t_end) of d_seq:
 *   for t in [t_begin, t_head)  : the layers consume sample t (teacher-forced) and refresh the dilation rings
 *                                 — this is the prefill; pass t_begin = t_head - rf + ... see mmk_wavenet_generate
 *   for t in [t_head, t_end)    : additionally the head predicts sample t+1; unless `teacher_forced`, the decision
 *                                 is written to d_seq[:, t+1] and fed back.
 *   d_seq          int64 (B, seq_stride) mu-law indices, in place; column j holds time seq_t0 + j (time also selects
 *                  the ring slot t mod dilation, so a caller stepping one sample at a time keeps t running and
 *                  slides seq_t0 along with a small scratch buffer)
 *   d_temperature  NULL => argmax (targets.py:42-43); else fp32 (n_temperature in {1, B}) (targets.py:27-34)
 *   d_noise        fp32 (B, noise_stride): uniform [0,1) draw for the decision made at time t is
 *                  d_noise[b, t + 1 - noise_t0]; required iff d_temperature != NULL
 *   d_logits_out   nullable fp32 (B, t_end - t_head, Q): head logits (after the learned-temperature divide)
 *   d_decisions    nullable int64 (B, t_end - t_head): the decision for every head step (useful when teacher_forced)
 *   d_step_ts      nullable uint64 (t_end - t_begin): globaltimer (ns) when the last prompt group left the last stage
 */
int mmk_wavenet_run(mmk_wavenet_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin,
                    int64_t t_head, int64_t t_end, int teacher_forced, const float* d_temperature, int n_temperature,
                    const float* d_noise, int64_t noise_stride, int64_t noise_t0, float* d_logits_out,
                    int64_t* d_decisions, unsigned long long* d_step_ts, void* stream);

/* Convenience = GenerateLoopV2.run's inner loop for one batch (loops/generate.py:195-219): prefill over the last
 * rf prompt samples, then n_steps autoregressive steps.  d_seq is (B, >= P + n_steps); noise is (B, n_steps). */
int mmk_wavenet_generate(mmk_wavenet_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t prompt_len,
                         int64_t n_steps, const float* d_temperature, int n_temperature, const float* d_noise,
                         float* d_logits_out, unsigned long long* d_step_ts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SampleRNN — SampleRNN.before_generate / generate_step (mimikit/networks/sample_rnn_v2.py:226-260),
 * SampleRNNTier.forward (83-99), FramedLinearIO / FramedConv1dIO (modules/io.py:106-133, 185-198),
 * LinearResampler (modules/resamplers.py:13-23), nn.GRU cell, MLP head, CategoricalSampler.
 * Supported: single mu-law framed_linear input, rnn_class "gru", n_rnn 1, h0 zeros, inputs_mode "sum".
 * ---------------------------------------------------------------------------------------------- */
typedef struct mmk_samplernn_s* mmk_samplernn_t;

typedef struct {
    int n_tiers;                  /* len(frame_sizes), >= 2; the last tier is the sample-level MLP input */
    const int* frame_sizes;       /* [n_tiers] */
    int hidden_dim;               /* H */
    int head_hidden;
    int q_levels;
    float min_temperature;
    /* HOST pointers per frame tier i < n_tiers-1 (weight-norm already folded) */
    const float* const* in_w;     /* tiers.i.input_module.heads.0.2.weight (H, fs_i) */
    const float* const* in_b;     /* ...bias (H) */
    const float* const* w_ih;     /* tiers.i.rnn.weight_ih_l0 (3H, H) */
    const float* const* w_hh;     /* tiers.i.rnn.weight_hh_l0 (3H, H) */
    const float* const* b_ih;     /* (3H) */
    const float* const* b_hh;     /* (3H) */
    const float* const* up_w;     /* tiers.i.up_sampler.fc.weight (H*up_i, H) */
    const float* const* up_b;     /* (H*up_i) */
    const float* conv_w;          /* tiers.(n-1).input_module.heads.0.2.2.cv.weight (H, 1, fs_last) */
    const float* conv_b;          /* (H) */
    const float* head_w1; const float* head_b1; const float* head_w2; const float* head_b2;
} mmk_samplernn_desc;

int mmk_samplernn_create(const mmk_samplernn_desc* desc, int max_batch, mmk_samplernn_t* out);

/* The rest of SampleRNNTier's configuration surface (sample_rnn_v2.py:40-66, 101-119; networks/mlp.py:44-50):
 * rnn_class "lstm" (the reference default) / "gru" / "rnn", n_rnn stacked layers, a non-zero initial state, and
 * n_hidden_layers > 0 in the MLP head.  The GRU or LSTM / one layer / zero state / plain head form runs in the cluster kernel
 * (csrc/samplernn2.cu: LSTM at hidden_dim 128, 256, 512); everything else in the general kernel (csrc/samplernn.cu). */
#define MMK_RNN_GRU 0
#define MMK_RNN_LSTM 1
#define MMK_RNN_TANH 2
typedef struct {
    mmk_samplernn_desc base;      /* geometry, input / up-sampler / bottom-tier weights, head fc.0 and the LAST head Linear
                                     (head_w2 = fc.{2 + 2 n_hidden}.weight); base.w_ih .. base.b_hh are ignored */
    int rnn_type;                 /* MMK_RNN_* : nn.GRU (gates r, z, n) | nn.LSTM (i, f, g, o) | nn.RNN (tanh) */
    int n_rnn;                    /* stacked layers per frame tier, 1..4 */
    /* HOST pointers, entry [tier * n_rnn + layer] = tiers.i.rnn.weight_ih_l{layer} (G*H, H), weight_hh_l{layer} (G*H, H),
     * bias_ih_l{layer}, bias_hh_l{layer} (G*H), G = 3 | 4 | 1 */
    const float* const* w_ih; const float* const* w_hh; const float* const* b_ih; const float* const* b_hh;
    int head_hidden_layers;       /* MLP n_hidden_layers.  mlp.py:47-50 repeats a tuple holding ONE nn.Linear(Hh, Hh), so all the
                                     hidden layers share the weights given here */
    const float* head_wh;         /* output_modules.0.estimator.0.fc.2.weight (Hh, Hh) when head_hidden_layers > 0 */
    const float* head_bh;         /* ...fc.2.bias (Hh) */
    int need_set_hidden;          /* 1 = mmk_samplernn_set_hidden will be used (h0_init "ones" / "randn"): never the tile engine */
    int compute_mode;             /* MMK_COMPUTE_FP32 (0) | MMK_COMPUTE_BF16_TC (1): frame-tier GRU and up-sampler contractions on
                                     tcgen05 with bf16 operands and fp32 accumulation (logits within 5e-2 of fp32).  GRU or LSTM, one layer,
                                     plain head, hidden_dim in {128, 256, 512}, frame sizes in {1, 2, 4, 8, 16}, max_batch <= 128; else an error */
} mmk_samplernn_desc_ex;
int mmk_samplernn_create_ex(const mmk_samplernn_desc_ex* desc, int max_batch, mmk_samplernn_t* out);
/* Installs an initial state — SampleRNNTier._reset_hidden / _init_h0 (sample_rnn_v2.py:101-119) with h0_init "ones" or
 * "randn": d_values fp32 (B, H) on the device; which = 0 the hidden state, 1 the LSTM cell state.  Call after a run with
 * reset_hidden = 1 and zero-length ranges (or on a fresh handle), then run with reset_hidden = 0. */
int mmk_samplernn_set_hidden(mmk_samplernn_t h, int tier, int layer, int which, const float* d_values, int B, void* stream);
int mmk_samplernn_destroy(mmk_samplernn_t h);
int mmk_samplernn_launch_info(mmk_samplernn_t h, mmk_launch_info* out);
/* As mmk_wavenet_sync_check: waits for the stream and fails if a grid barrier of the last launch timed out. */
int mmk_samplernn_sync_check(mmk_samplernn_t h, void* stream);

/* One persistent-kernel launch covering two consecutive ranges of generate_step calls (sample_rnn_v2.py:236-260):
 *   warm-up  : logical t in [warm_begin, warm_end) — frame tiers only, reading the window that ends at data index
 *              t + warm_offset (before_generate's loop, :232-234, with offset = P % frame_sizes[0])
 *   generate : absolute t in [gen_begin, gen_end) — frame tiers on their clocks (t % fs_i == 0), then the bottom
 *              tier, the head and the sampler produce sample t; unless `teacher_forced` it is written to d_seq[:, t]
 *   reset_hidden: 1 = zero the GRU states first (SampleRNN.reset_hidden, :266-268)
 *   d_seq    int64 (B, seq_stride); column j holds time seq_t0 + j
 *   d_noise  fp32 (B, noise_stride): the draw for sample t is d_noise[b, t - noise_t0]
 *   d_logits_out nullable fp32 (B, gen_end - gen_begin, Q); d_decisions nullable int64 (B, gen_end - gen_begin)
 *   d_step_ts nullable uint64 (gen_end - gen_begin): globaltimer (ns) at the end of each generate step */
int mmk_samplernn_run(mmk_samplernn_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0,
                      int64_t warm_begin, int64_t warm_end, int64_t warm_offset, int64_t gen_begin, int64_t gen_end,
                      int reset_hidden, int teacher_forced, const float* d_temperature, int n_temperature,
                      const float* d_noise, int64_t noise_stride, int64_t noise_t0, float* d_logits_out,
                      int64_t* d_decisions, unsigned long long* d_step_ts, void* stream);

/* before_generate (hidden reset + warm-up over the prompt, sample_rnn_v2.py:226-234) followed by n_steps of
 * generate_step driven as loops/generate.py:207-219.  d_seq is (B, >= P + n_steps); noise is (B, n_steps). */
int mmk_samplernn_generate(mmk_samplernn_t h, int64_t* d_seq, int B, int64_t seq_stride, int64_t prompt_len,
                           int64_t n_steps, const float* d_temperature, int n_temperature, const float* d_noise,
                           float* d_logits_out, unsigned long long* d_step_ts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMK_B200_H */
