// FullSubNet model level (fsnet.cu): same opaque plan type as DCCRN / CRN (kind = 2).
#pragma once
#include "common.cuh"

struct sefd_plan;
sefd_plan* sefd_fsn_plan_create_impl(int B, int frames);
void sefd_fsn_plan_free_ext(sefd_plan* P);
int sefd_fsn_forward_impl(const sefd_plan* P, const float* prm, const float* noisy_mag, int train, float dropout_p,
                          const float* mask_fb, const float* mask_sb, unsigned long long seed, float* crm, void* ws,
                          size_t ws_bytes, cudaStream_t st);
int sefd_fsn_backward_impl(const sefd_plan* P, const float* prm, const float* d_crm, float* grads, void* ws, size_t ws_bytes,
                           cudaStream_t st);
int sefd_fsn_tensor_info(const sefd_plan* P, const char* name, long long* off, int* ndim, long long shape[4]);
