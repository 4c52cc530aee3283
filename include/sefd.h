/* sefd — C ABI of the B200-native DCCRN speech-enhancement train-step path (libsefd.so).
 *
 * Every entry point replaces one piece of the reference's PyTorch path (file:line are relative
 * to the reference checkout).  Conventions:
 *   - all pointers are DEVICE pointers to fp32 unless stated; buffers are caller-owned
 *     (the Python host passes torch CUDA tensors' data_ptr()); the library allocates nothing
 *     on the device;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it;
 *   - return value 0 = success, negative = failure with the message in sefd_last_error();
 *     nothing throws across the ABI; there is no CPU fallback;
 *   - activations are channels-last: element (b, f, t, c) of a [B][F][T][C] tensor is contiguous in c;
 *   - STFT geometry is the reference default win 400 / hop 100 / fft 512 / periodic Hann
 *     (config.py:55-61); waveform length L must be a multiple of 100 and T = L/100 + 3 frames.
 */
#ifndef SEFD_H
#define SEFD_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sefd_plan sefd_plan;

/* masking modes (config.py:27, models.py:258-276) and losses (config.py:23, models.py:315-323) */
enum { SEFD_MODE_E = 1, SEFD_MODE_C = 2, SEFD_MODE_R = 3, SEFD_MODE_DIRECT = 5 /* 'Direct(None make)', models.py:232-250 */ };
enum { SEFD_MSE = 0, SEFD_SDR = 1, SEFD_SI_SNR = 2, SEFD_SI_SDR = 3 };

int sefd_abi_version(void);
const char* sefd_last_error(void);

/* Non-sticky CUDA errors that were pending in the calling thread when one of this library's launches began (left by
 * other code in the process; cudaGetLastError semantics): they are absorbed and counted here instead of being reported
 * as a failure of the unrelated kernel that happened to be launched next. */
int sefd_stale_cuda_errors(void);
const char* sefd_last_stale_cuda_error(void);

/* ---- op level ------------------------------------------------------------------------------- */

/* ConvSTFT.forward, 'complex' (tools_for_model.py:54-61).  wav [B][L] -> spec [B][257][T][2] (re, im). */
int sefd_stft_forward(const float* wav, float* spec, int B, int L, void* stream);

/* ConviSTFT.forward (tools_for_model.py:90-112) of spec [B][257][T][2] -> wav [B][L] (no clamp). */
int sefd_istft_forward(const float* spec, float* wav, int B, int L, void* stream);
/* its adjoint: dwav [B][L] -> dspec [B][257][T][2] */
int sefd_istft_backward(const float* dwav, float* dspec, int B, int L, void* stream);

/* mask apply + ISTFT + clamp (models.py:253-282).  mask [B][256][T][2] holds (re, im) for bins 1..256
 * (the DC mask is zero, models.py:255-256).  out_real/out_imag [B][257][T] may be NULL. raw_wav [B][L]
 * (pre-clamp, needed by the backward) may be NULL. */
int sefd_mask_istft_forward(const float* spec, const float* mask, int mode, int B, int L, float* out_real,
                            float* out_imag, float* out_wav, float* raw_wav, void* stream);
int sefd_mask_istft_backward(const float* dwav, const float* raw_wav, const float* spec, const float* mask, int mode,
                             int B, int L, float* dmask, void* stream);

/* The same three ops for either transform geometry the reference's config allows (config.py:55-61; init_kernels
 * tools_for_model.py:16-33 is generic in win_len / win_inc / fft_len): nfft = 512 -> win 400 / hop 100 / F = 257 bins,
 * nfft = 1024 -> win 800 / hop 200 / F = 513 bins; T = L / hop + 3; spec [B][F][T][2]; mask [B][F-1][T][2] (DC mask zero).
 * sefd_mask_istft_forward_n: mode 0 = plain ISTFT of spec (mask unused), output clamped to [-1, 1] like the model's.
 * sefd_stft_mask_istft_fused: wave -> STFT -> mask -> ISTFT -> wave in ONE kernel, the spectrum never leaves the SM
 * (BASELINE configs[4]'s fused form: 8 (F-1) + 8 hop bytes per frame); wav and out_wav must not alias. */
int sefd_stft_forward_n(const float* wav, float* spec, int B, int L, int nfft, void* stream);
int sefd_mask_istft_forward_n(const float* spec, const float* mask, int mode, int B, int L, int nfft, float* out_wav,
                              void* stream);
int sefd_stft_mask_istft_fused(const float* wav, const float* mask, int mode, int B, int L, int nfft, float* out_wav,
                               void* stream);

/* DCCRN.loss, non-perceptual branch (models.py:315-323; tools_for_loss.py:29-94).
 * scratch: 8*B doubles; coef: 2*B floats (kept for the backward); loss: 1 float. */
int sefd_loss_forward(const float* est, const float* target, int B, int L, int kind, double* scratch, float* loss,
                      float* coef, void* stream);
/* d_est[b][n] = gout * (coef[b][0]*est + coef[b][1]*target); gout (1 float, device) may be NULL (= 1). */
int sefd_loss_backward(const float* est, const float* target, const float* coef, const float* gout, float* d_est,
                       int B, int L, void* stream);

/* LMS perceptual loss (get_array_lms_loss, tools_for_loss.py:111-249; called from DCCRN.loss models.py:305-312):
 * log-mel-spectrum RMSE at mel scales 16/32/64 between sqrt(|STFT(target)|^2 + 1e-7) and
 * sqrt(out_real^2 + out_imag^2 + 1e-7), rows formed by the reference's x.view(-1, 257) reshape.
 * est_real / est_imag [B][257][T]; clean_spec [B][257][T][2] (sefd_stft_forward of the target);
 * F [257][112] = the three melFilterBank(M, 512)^T side by side (M = 16, 32, 64), Ft [112][257] its transpose (built
 * by the host exactly like the reference builds them); scratch: 1 double; loss: 1 float.
 * inputs_are_mags = 1: est_real and clean_spec already hold magnitudes [B][257][T] (the generic
 * get_array_lms_loss(clean_mags, est_mags) signature); est_imag / d_imag are then unused. */
int sefd_lms_forward(const float* est_real, const float* est_imag, const float* clean_spec, const float* F, int B, int T,
                     int inputs_are_mags, double* scratch, float* loss, void* stream);
int sefd_lms_backward(const float* est_real, const float* est_imag, const float* clean_spec, const float* F, const float* Ft,
                      const float* gout, int B, int T, int inputs_are_mags, float* d_real, float* d_imag, void* stream);

/* PMSQE perceptual loss: get_array_pmsqe_loss(clean_array, est_array), tools_for_loss.py:255-269 =
 * PITLossWrapper(SingleSrcPMSQE(), 'pw_pt') on mag(Encoder(STFTFB(512, 512, stride 256))) of the waveforms cut into
 * 1-second chunks.  PARITY UNPINNED (the arithmetic is asteroid's, absent from the reference tree): see csrc/pmsqe.cu and
 * oracle/pmsqe_oracle.py.  est_wav / clean_wav [N][L], L = S * 16000 with S <= 4.
 * tables: sefd_pmsqe_table_floats() floats = [bark_matrix 257 x 49 | abs_thresh_power 49 | modified_zwicker_power 49 |
 * width_of_band_bark 49 | mask_sll 257] (SingleSrcPMSQE's buffers; the host builds them from the ITU-T P.862 tables or
 * passes asteroid's own).  ws: sefd_pmsqe_workspace_bytes(N, L) bytes, filled by forward and read by backward.
 * backward: d_est [N][L] = gout[0] * d loss / d est_wav (gout may be NULL = 1). */
int sefd_pmsqe_table_floats(void);
size_t sefd_pmsqe_workspace_bytes(int N, int L);
int sefd_pmsqe_forward(const float* est_wav, const float* clean_wav, int N, int L, const float* tables, void* ws,
                       size_t ws_bytes, float* loss, void* stream);
int sefd_pmsqe_backward(const float* gout, int N, int L, const float* tables, void* ws, size_t ws_bytes, float* d_est,
                        void* stream);

/* FullSubNet feature / target side of trainer.fullsubnet_train (trainer.py:97-104): torch.stft geometry n_fft 512, hop 300,
 * win 400 (centred, reflect padding); T = sefd_fsn_frames(L) = L / 300 + 1.
 *   sefd_fsn_features: tools.stft x2 + tools.mag_phase + tools.build_complex_ideal_ratio_mask fused (tools_for_model.py:
 *                      628-717): noisy, clean [B][L] -> noisy_mag [B][257][T], cirm [B][257][T][2]
 *   sefd_fsn_stft / _mag_phase / _cirm / _decompress_cirm: the same functions one by one (spec = complex viewed as [..][2]) */
int sefd_fsn_frames(int L);
int sefd_fsn_features(const float* noisy, const float* clean, int B, int L, float* noisy_mag, float* cirm, void* stream);
int sefd_fsn_stft(const float* wav, int B, int L, float* spec, void* stream);
int sefd_fsn_mag_phase(const float* spec, long long n, float* mag, float* phase, void* stream);
int sefd_fsn_cirm(const float* noisy_spec, const float* clean_spec, long long n, float* cirm, void* stream);
int sefd_fsn_decompress_cirm(const float* mask, long long n, float* out, void* stream);
/* compress_cIRM (tools_for_model.py:707-717, K = 10, C = 0.1) by itself: out = K (1 - e^{-C m}) / (1 + e^{-C m}), m clipped at -100 */
int sefd_fsn_compress_cirm(const float* mask, long long n, float* out, void* stream);
/* tools.istft (tools_for_model.py:651-679, torch.istft 512 / 300 / 400, centred, Hann): spec [B][257][T][2], or magnitude and
 * phase [B][257][T] when phase != NULL (use_mag_phase=True), -> wav [B][len] */
int sefd_fsn_istft(const float* spec_or_mag, const float* phase, int B, int T, int len, float* wav, void* stream);

/* ComplexConv2d (tools_for_model.py:199-269: kernel (5,2), stride (2,1), pad (2,0), causal pad 1) and
 * ComplexConvTranspose2d (tools_for_model.py:272-338: + output_padding (1,0)); channels-last tensors.
 *   conv : x [B][F][T][Cin]          -> y [B][F/2][T][Cout]
 *   convT: x0,x1 [B][F][T][Cin/2]    -> y [B][2F][T+1][Cout]   (x1 = skip tensor of complex_cat, may be NULL
 *                                                                  when Cin/2 is the whole input)
 * wr, wi, br, bi are the reference's real_conv / imag_conv parameters in their own layouts.
 * ws: sefd_cconv_workspace_bytes(Cin, Cout) bytes of scratch (holds the packed block-real operand). */
size_t sefd_cconv_workspace_bytes(int Cin, int Cout);
int sefd_cconv2d_forward(const float* x, const float* wr, const float* br, const float* wi, const float* bi, float* y,
                         int B, int F, int T, int Cin, int Cout, void* ws, void* stream);
int sefd_cconv2d_backward(const float* x, const float* wr, const float* wi, const float* dy, float* dx, float* dwr,
                          float* dbr, float* dwi, float* dbi, int B, int F, int T, int Cin, int Cout, void* ws,
                          void* stream);
int sefd_cconvT2d_forward(const float* x0, const float* x1, const float* wr, const float* br, const float* wi,
                          const float* bi, float* y, int B, int F, int T, int Cin, int Cout, void* ws, void* stream);
int sefd_cconvT2d_backward(const float* x0, const float* x1, const float* wr, const float* wi, const float* dy,
                           float* dx0, float* dx1, float* dwr, float* dbr, float* dwi, float* dbi, int B, int F, int T,
                           int Cin, int Cout, void* ws, void* stream);

/* nn.BatchNorm2d (train mode) + nn.PReLU (models.py:76-78) on y [rows][C]; z [rows][C].
 * save [2][C] (mean, inv-std); running_* may be NULL; scratch: (2*C+1) doubles. */
int sefd_bn_prelu_forward(const float* y, float* z, long long rows, int C, const float* gamma, const float* beta,
                          const float* alpha, float* save, float* running_mean, float* running_var, double* scratch,
                          void* stream);
int sefd_bn_prelu_backward(const float* y, const float* dz, float* dy, long long rows, int C, const float* gamma,
                           const float* beta, const float* alpha, const float* save, float* dgamma, float* dbeta,
                           float* dalpha, double* scratch, void* stream);

/* ComplexBatchNorm (tools_for_model.py:430-603) + PReLU on channels-last rows [rows][C], C = 2 h: real parts in [0, h),
 * imaginary parts in [h, 2 h).  w3h = Wrr | Wri | Wii, b2h = Br | Bi, running5h = RMr | RMi | RVrr | RVri | RVii (each h
 * floats; train mode lerps the batch moments into it when non-NULL, use_running = 1 normalises with it); alpha: PReLU slope
 * (1 float on the device; 1.0 gives the bare module); save: 9 h floats kept for the backward; scratch: 5 h doubles (forward),
 * 6 h + 1 doubles (backward); coef: 9 h floats of backward scratch. */
int sefd_cbn_prelu_forward(const float* y, float* z, long long rows, int C, const float* w3h, const float* b2h, const float* alpha,
                           float* save, float* running5h, int use_running, double* scratch, void* stream);
int sefd_cbn_prelu_backward(const float* y, const float* dz, float* dy, long long rows, int C, const float* w3h, const float* b2h,
                            const float* alpha, const float* save, float* dw3h, float* db2h, float* dalpha, double* scratch,
                            float* coef, void* stream);

/* recurrent part of two nn.LSTM(H=128) run side by side (tools_for_model.py:147-150,167-170):
 * gates [2][rows][T][512] holds x W_ih^T + b_ih + b_hh on entry and the activated gates on exit. */
int sefd_lstm_forward(const float* w_hh, float* gates, float* h, float* c, int rows, int T, void* stream);
int sefd_lstm_backward(const float* w_hh, const float* gates, const float* c, const float* dh, float* dgates, int rows,
                       int T, void* stream);

/* torch.optim.Adam, defaults of train_interface.py:59 (no weight decay, no amsgrad); step counts from 1.
 * grads are multiplied by gscale first (1/world_size after an allreduce-sum). */
int sefd_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, int step, float gscale, void* stream);

/* same update with the step count kept ON THE DEVICE (step_dev: 1 int, starts at 0 and is incremented by the call;
 * bc_scratch2: 2 floats of scratch): no host-computed value changes from step to step, so a whole train step can be captured
 * in one CUDA graph and replayed. */
int sefd_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                       float beta1, float beta2, float eps, int* step_dev, float* bc_scratch2, float gscale, void* stream);

/* y[i] = a x[i] + b y[i]: the mixing of the perceptual step, loss = (main + perceptual) / 2 and the matching sum of
 * the two waveform gradients (trainer.py:166-169), as a kernel of this library instead of framework element-wise ops. */
int sefd_axpby(float* y, const float* x, float a, float b, long long n, void* stream);
/* *counters_dev[i] += inc for n int64 device counters whose addresses sit in a DEVICE array: BatchNorm2d's
 * num_batches_tracked of every layer after a train-mode forward (torch BatchNorm semantics) in one launch. */
int sefd_counters_inc(long long* const* counters_dev, int n, long long inc, void* stream);

/* ---- model level: DCCRN.forward / autograd backward (models.py:176-284) ---------------------- */
sefd_plan* sefd_dccrn_plan_create(int B, int L, int masking_mode);
/* flags: SEFD_PLAN_NO_SKIP builds the decoder of cfg.skip_type = False (models.py:138-169, 227-230): the transposed
 * convolutions read the previous output only (Cin = kernel_num[idx], no complex_cat with the encoder output) */
#define SEFD_PLAN_NO_SKIP 1
/* SEFD_PLAN_REAL_LSTM: cfg.lstm = 'real' (models.py:96-105, 213-218): the recurrent part is self.enhance = nn.LSTM(1024 -> 256,
 * 2 layers) + self.tranform = Linear(256 -> 1024) on the [T, B, C * D] view of the encoder output (state_dict keys
 * enhance.weight_ih_l0 ... tranform.bias) instead of the two NavieComplexLSTM layers */
#define SEFD_PLAN_REAL_LSTM 2
/* SEFD_PLAN_CBN: DCCRN(use_cbn=True) (models.py:26, 76, 120, 151): every BatchNorm2d(C) becomes ComplexBatchNorm(C)
 * (tools_for_model.py:430-603) - parameters .1.Wrr / Wri / Wii / Br / Bi and buffers .1.RMr / RMi / RVrr / RVri / RVii over
 * C / 2 complex features (2x2 whitening + affine), train and eval mode, forward and backward (csrc/cbn.cu). */
#define SEFD_PLAN_CBN 4
sefd_plan* sefd_dccrn_plan_create_ex(int B, int L, int masking_mode, int flags);
void sefd_dccrn_plan_destroy(sefd_plan* plan);
size_t sefd_dccrn_workspace_bytes(const sefd_plan* plan);
long long sefd_dccrn_param_floats(const sefd_plan* plan);    /* size of the flat parameter / gradient buffer */
long long sefd_dccrn_buffer_floats(const sefd_plan* plan);   /* size of the flat BN running-stat buffer */
int sefd_dccrn_num_params(const sefd_plan* plan);
int sefd_dccrn_num_buffers(const sefd_plan* plan);
/* idx-th parameter (kind 0) or BN buffer (kind 1): reference state_dict key, offset into the flat buffer, shape */
int sefd_dccrn_entry_info(const sefd_plan* plan, int kind, int idx, char* name, int name_cap, long long* offset,
                          long long* numel, int* ndim, long long shape[4]);
/* named intermediate in the workspace (tests / debugging): "spec", "enc3.y", "enc3.z", "dec0.y", "U", ... */
int sefd_dccrn_tensor_info(const sefd_plan* plan, const char* name, long long* offset_floats, int* ndim,
                           long long shape[4]);

/* params: flat parameter buffer; bn_buffers: flat running stats (updated in train mode; read in eval mode);
 * noisy [B][L]; target [B][L] or NULL (when given, the loss dot products are accumulated on the fly);
 * out_real/out_imag [B][257][T] (may be NULL), out_wav [B][L]. */
int sefd_dccrn_forward(const sefd_plan* plan, const float* params, float* bn_buffers, const float* noisy,
                       const float* target, int train, float* out_real, float* out_imag, float* out_wav, void* ws,
                       size_t ws_bytes, void* stream);
/* d_wav [B][L] -> grads (flat, same layout as params; every entry is overwritten) */
int sefd_dccrn_backward(const sefd_plan* plan, const float* params, const float* d_wav, float* grads, void* ws,
                        size_t ws_bytes, void* stream);
/* same as sefd_dccrn_backward, for data-parallel steps: `tail_ready_event` (a cudaEvent_t, may be NULL) is recorded as soon as
 * every gradient at flat offset >= sefd_dccrn_grad_split(plan) is final (decoder, projection and LSTM parameters: 76 % of
 * the buffer, finished before the encoder backward starts), so that slice can be all-reduced on another stream while the
 * encoder backward runs; the head slice [0, split) is final when the call's work on `stream` is. */
long long sefd_dccrn_grad_split(const sefd_plan* plan);
int sefd_dccrn_backward_overlap(const sefd_plan* plan, const float* params, const float* d_wav, float* grads, void* ws,
                                size_t ws_bytes, void* stream, void* tail_ready_event);
/* same, with an additional gradient arriving at the masked spectrum out_real / out_imag [B][257][T] (perceptual losses,
 * models.py:305-312); d_wav or the pair (d_out_real, d_out_imag) may be NULL */
int sefd_dccrn_backward_spec(const sefd_plan* plan, const float* params, const float* d_wav, const float* d_out_real,
                             const float* d_out_imag, float* grads, void* ws, size_t ws_bytes, void* stream);
/* loss on the plan's own outputs, reusing dot products gathered by the forward epilogue when `target` was given */
int sefd_dccrn_loss(const sefd_plan* plan, const float* out_wav, const float* target, int kind, int reuse_dots,
                    float* loss, float* coef, void* ws, void* stream);

/* ---- model level: CRN.forward / autograd backward (models.py:329-565; RealConv2d / RealConvTranspose2d,
 * tools_for_model.py:341-425).  The plan is the same opaque type: every sefd_dccrn_* accessor above
 * (workspace_bytes, param_floats, buffer_floats, num_params, num_buffers, entry_info, tensor_info, plan_destroy,
 * loss) also serves a CRN plan; state_dict keys are the reference's (encoder.i.0.conv.weight, enhance.weight_ih_l0,
 * tranform.weight, ...).  Mask: est_mags = tanh(out) * |X| with the noisy phase (models.py:518-524).
 * est_mags / target_mags [B][257][T] may be NULL (target_mags also needs `target`). */
sefd_plan* sefd_crn_plan_create(int B, int L);
int sefd_crn_forward(const sefd_plan* plan, const float* params, float* bn_buffers, const float* noisy,
                     const float* target, int train, float* est_mags, float* target_mags, float* out_wav, void* ws,
                     size_t ws_bytes, void* stream);
int sefd_crn_backward(const sefd_plan* plan, const float* params, const float* d_wav, float* grads, void* ws,
                      size_t ws_bytes, void* stream);
/* same, with an additional gradient arriving at est_mags [B][257][T] (CRN.loss perceptual branch, models.py:553-555);
 * d_wav or d_est_mags may be NULL */
int sefd_crn_backward_spec(const sefd_plan* plan, const float* params, const float* d_wav, const float* d_est_mags,
                           float* grads, void* ws, size_t ws_bytes, void* stream);

/* ---- model level: FullSubNet.forward / autograd backward (models.py:568-682; SequenceModel tools_for_model.py:726-795,
 * BaseModel.unfold :806-837, offline_laplace_norm :997-1011).  Same opaque plan type (workspace_bytes, param_floats,
 * num_params, entry_info, tensor_info, plan_destroy serve it); state_dict keys are the reference's
 * (fb_model.sequence_model.weight_ih_l0, ..., sb_model.fc_output_layer.bias).  Configuration of config.py:71-80:
 * 15 sub-band neighbours, 0 full-band neighbours, look-ahead 2, LSTM 257 -> 512 -> 512 -> Linear 257 + ReLU (full band),
 * LSTM 32 -> 384 -> 384 -> Linear 2 (sub band), offline_laplace_norm.
 *   noisy_mag [B][257][frames] (tools.mag_phase of tools.stft, or sefd_fsn_features) -> crm [B][257][frames][2].
 * nn.LSTM(dropout = 0.8) between the stacked layers (tools_for_model.py:746) is active when train != 0 and dropout_p > 0:
 * mask_fb [T][B][512] / mask_sb [T][B*257][384] (T = frames + 2) are optional injected multipliers (0 or 1 / (1 - p));
 * when NULL the masks come from Philox4x32-10 keyed by `seed` (the backward regenerates them; injected masks must stay
 * alive until the backward has run). */
sefd_plan* sefd_fsn_plan_create(int B, int frames);
int sefd_fsn_forward(const sefd_plan* plan, const float* params, const float* noisy_mag, int train, float dropout_p,
                     const float* mask_fb, const float* mask_sb, unsigned long long seed, float* crm, void* ws, size_t ws_bytes,
                     void* stream);
/* the inter-layer dropout op by itself: y = x * m, m = mask[i] when mask != NULL, else 0 or 1 / (1 - p) from
 * Philox4x32-10(counter = (i / 4, stream_id), key = seed); n must be a multiple of 4; x == y is allowed */
int sefd_dropout_forward(const float* x, float* y, long long n, float p, const float* mask, unsigned long long seed,
                         unsigned int stream_id, void* stream);
/* d_crm [B][257][frames][2] -> grads (flat, same layout as params; every entry is overwritten) */
int sefd_fsn_backward(const sefd_plan* plan, const float* params, const float* d_crm, float* grads, void* ws, size_t ws_bytes,
                      void* stream);

/* ---- data-parallel plumbing without torch: one NCCL communicator per process / GPU; the only collective on
 * the path is a sum all-reduce of (slices of) the flat fp32 gradient buffer.  libnccl.so.2 is bound at run time (dlopen;
 * SEFD_NCCL_LIB overrides the name).  Rank 0 creates the 128-byte id with sefd_nccl_unique_id and hands it to the other ranks
 * by any host-side channel (file, MPI, torchrun's store, ...). */
typedef struct sefd_comm sefd_comm;
int sefd_nccl_unique_id_bytes(void);
int sefd_nccl_unique_id(void* out128);
sefd_comm* sefd_nccl_init(int rank, int world, const void* unique_id128);
int sefd_nccl_allreduce(sefd_comm* comm, float* buf, long long n, void* stream);   /* in place, sum, asynchronous on stream */
int sefd_nccl_world(const sefd_comm* comm);
void sefd_nccl_destroy(sefd_comm* comm);

/* ---- measurement support (bench.py): CUDA-event timing per kernel category on the launching stream.
 * categories: 0 tap-GEMM (conv/convT/linear fwd + dgrad), 1 weight gradients, 2 BN+PReLU passes,
 * 3 LSTM recurrence, 4 STFT/ISTFT/loss, 5 packing/reductions/Adam, 6 CUDA-core kernels of the 2-channel layers
 * (encoder 0 / decoder 5 forward, data and weight gradients).  flops/bytes are the ALGORITHMIC figures
 * of the recorded launches (DESIGN.md states the formulas). */
/* GEMM engine: 1 = tcgen05 TF32 tensor cores wherever a contraction is eligible (default), 0 = fp32 CUDA cores
 * everywhere (bit-for-bit the reference's fp32 arithmetic order aside; used by the parity tests as the exact engine). */
int sefd_set_engine(int engine);
int sefd_get_engine(void);
long long sefd_launch_count(void);   /* kernels launched by this library since load */
int sefd_prof_enable(int on);
int sefd_prof_reset(void);
int sefd_prof_get(int category, double* ms, long long* launches, double* flops, double* bytes);
int sefd_prof_dump(const char* csv_path);   /* one row per recorded launch */

#ifdef __cplusplus
}
#endif
#endif
