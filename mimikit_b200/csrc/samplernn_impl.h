// Internal interface between the SampleRNN C ABI (samplernn.cu) and the cluster kernel (samplernn2.cu).
#pragma once
#include "../../include/mmk_b200.h"

struct sr2_handle;

// Returns 0 and a handle when the configuration fits the cluster kernel; returns 1 with *unsupported = 1 when the
// caller should use the general kernel, or 1 with *unsupported = 0 on a real error (message set).
int sr2_create(const mmk_samplernn_desc* d, int max_batch, int tc, int lstm, int need_set_hidden, sr2_handle** out, int* unsupported);
int sr2_set_hidden(sr2_handle* h, int tier, int which, const float* d_values, int B, void* stream);
int sr2_destroy(sr2_handle* h);
int sr2_launch_info(sr2_handle* h, mmk_launch_info* out);
int sr2_sync_check(sr2_handle* h, void* stream);
// Same contract as mmk_samplernn_run (arguments already validated by the caller).
int sr2_run(sr2_handle* h, int64_t* d_seq, int B, int64_t seq_stride, int64_t seq_t0, int64_t warm_begin,
            int64_t warm_end, int64_t warm_offset, int64_t gen_begin, int64_t gen_end, int reset_hidden,
            int teacher_forced, const float* d_temperature, int n_temperature, const float* d_noise,
            int64_t noise_stride, int64_t noise_t0, float* d_logits_out, int64_t* d_decisions,
            unsigned long long* d_step_ts, void* stream);
