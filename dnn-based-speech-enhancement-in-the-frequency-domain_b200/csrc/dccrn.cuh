#pragma once
#include "common.cuh"
#include "elementwise.cuh"
#include "lstm.cuh"
#include "stft.cuh"

struct sefd_plan;
sefd_plan* sefd_plan_create_impl(int B, int L, int mask_mode, int flags);
int sefd_forward_impl(const sefd_plan* P, const float* prm, float* bnbuf, const float* noisy, const float* target,
                      int train, float* out_real, float* out_imag, float* out_wav, void* ws, size_t ws_bytes,
                      cudaStream_t st);
// tail_ready (optional event): recorded when every gradient at flat offset >= sefd_dccrn_grad_split(plan) - decoder, LSTM and
// projection parameters, which the backward finishes first - is final, so their all-reduce can overlap the encoder backward
int sefd_backward_impl(const sefd_plan* P, const float* prm, const float* dwav, const float* dreal, const float* dimag,
                       float* grads, void* ws, size_t ws_bytes, cudaStream_t st, cudaEvent_t tail_ready = nullptr);

// CRN (crn.cu): same plan type (kind = 1), own forward / backward
sefd_plan* sefd_crn_plan_create_impl(int B, int L);
int sefd_crn_forward_impl(const sefd_plan* P, const float* prm, float* bnbuf, const float* noisy, const float* target,
                          int train, float* est_mags, float* target_mags, float* out_wav, void* ws, size_t ws_bytes,
                          cudaStream_t st);
int sefd_crn_backward_impl(const sefd_plan* P, const float* prm, const float* dwav, const float* dmags, float* grads, void* ws,
                           size_t ws_bytes, cudaStream_t st);
int sefd_crn_tensor_info(const sefd_plan* P, const char* name, long long* off, int* ndim, long long shape[4]);
