// Internal interface between the WaveNet C ABI (wavenet.cu, which also holds the general fp32 kernel) and the specialised
// kernels.  Every create returns 0 and a handle when the configuration fits; 1 with *unsupported = 1 when the caller should
// fall back to the general kernel; 1 with *unsupported = 0 on a real error (message set).
#pragma once
#include "../../include/mmk_b200.h"

// The bf16 tensor-core kernel (wavenet_tc.cu, tcgen05 + TMEM): same contract; used when compute_mode = MMK_COMPUTE_BF16_TC.
struct wn4_handle;
int wn4_create(const mmk_wavenet_desc* d, int max_batch, wn4_handle** out, int* unsupported);
int wn4_destroy(wn4_handle* h);
int wn4_launch_info(wn4_handle* h, mmk_launch_info* out);
int wn4_sync_check(wn4_handle* h, void* stream);
int wn4_run(wn4_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
            int64_t t_end, int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream);

// The layer-pipelined fp32 kernel (wavenet6.cu): one CTA pair per layer, critical weights in registers; same contract; tried first.
struct wn6_handle;
int wn6_create(const mmk_wavenet_desc* d, int max_batch, wn6_handle** out, int* unsupported);
int wn6_destroy(wn6_handle* h);
int wn6_launch_info(wn6_handle* h, mmk_launch_info* out);
int wn6_sync_check(wn6_handle* h, void* stream);
int wn6_run(wn6_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
            int64_t t_end, int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream);

// The layer-pipelined tensor-core kernel (wavenet7.cu: weights resident as the MMA M operand, 16-prompt groups as N): same
// contract; tried first for compute_mode = MMK_COMPUTE_BF16_TC, W-30-shaped networks only (else wavenet_tc.cu).
struct wn7_handle;
int wn7_create(const mmk_wavenet_desc* d, int max_batch, wn7_handle** out, int* unsupported);
int wn7_destroy(wn7_handle* h);
int wn7_launch_info(wn7_handle* h, mmk_launch_info* out);
int wn7_sync_check(wn7_handle* h, void* stream);
int wn7_run(wn7_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t t_begin, int64_t t_head,
            int64_t t_end, int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream);
