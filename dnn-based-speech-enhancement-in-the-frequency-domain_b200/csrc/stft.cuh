#pragma once
#include "common.cuh"

// SEFD_MASK_MAG: CRN's real T-F mask (models.py:518-524): one float per bin, est_mag = tanh(m) * |X|, noisy phase
// SEFD_MASK_DIRECT: spectral mapping, masking_mode 'Direct(None make)' (models.py:232-250): the decoder output IS the
// enhanced spectrum (bins 1..256; DC zero-padded)
enum { SEFD_MASK_NONE = 0, SEFD_MASK_E = 1, SEFD_MASK_C = 2, SEFD_MASK_R = 3, SEFD_MASK_MAG = 4, SEFD_MASK_DIRECT = 5 };
enum { SEFD_LOSS_MSE = 0, SEFD_LOSS_SDR = 1, SEFD_LOSS_SISNR = 2, SEFD_LOSS_SISDR = 3 };

struct MaskIstftParams {
    const float* spec;       // [B][257][T][2] noisy spectrum (or the spectrum itself for SEFD_MASK_NONE)
    const float* mask;       // element (b, k-1, t + m_tshift) at mask + b*mB + (k-1)*mF + (t+m_tshift)*mT, 2 floats
    long long mB, mF, mT;
    int m_tshift;
    int mode;
    int B, L, T;
    float *out_real, *out_imag;   // [B][257][T] or nullptr (SEFD_MASK_MAG: out_real receives est_mags, out_imag is unused)
    float* out_wav;               // [B][L] clamped
    float* raw_wav;               // [B][L] before the clamp (kept for the backward) or nullptr
    const float* target;          // [B][L] or nullptr
    double* dots;                 // [B][8]
};

struct MaskIstftBwdParams {
    const float *dreal, *dimag;   // optional [B][257][T]: gradient arriving directly at the masked spectrum (perceptual
                                  // (SEFD_MASK_MAG: dreal = gradient at est_mags, dimag unused)
                                  // losses on out_real / out_imag, models.py:305-312); added to the ISTFT adjoint
    const float* dwav;       // [B][L] (may be nullptr: no gradient through the waveform)
    const float* raw_wav;    // [B][L] or nullptr (no clamp gating)
    const float* spec;       // [B][257][T][2]
    const float* mask;       // as above (only read for mode E)
    float* dmask;            // same addressing as mask
    long long mB, mF, mT;
    int m_tshift;
    int mode;
    int B, L, T;
};

int sefd_stft_launch(const float* wav, float* spec, int B, int L, int T, cudaStream_t st);
// ConvSTFT 'real' magnitudes (tools_for_model.py:63-66): spec [n][2] -> mag [n]
int sefd_spec_mag_launch(const float* spec, float* mag, long long n, cudaStream_t st);
int sefd_mask_istft_launch(const MaskIstftParams& p, cudaStream_t st);
int sefd_mask_istft_bwd_launch(const MaskIstftBwdParams& p, cudaStream_t st);
// either transform geometry of config.py:55-61 (nfft 512: win 400 / hop 100 / 257 bins; nfft 1024: win 800 / hop 200 / 513 bins);
// spec [B][nfft/2+1][T][2], T = L / hop + 3.  fused_wav != null: wave -> STFT -> mask -> ISTFT in one kernel (p.spec unused).
int sefd_stft_launch_n(const float* wav, float* spec, int B, int L, int nfft, cudaStream_t st);
int sefd_mask_istft_launch_n(const MaskIstftParams& p, const float* fused_wav, int nfft, cudaStream_t st);
int sefd_loss_fwd_launch(const float* est, const float* tgt, int B, int L, int kind, double* dots, int dots_ready,
                         float* loss, float* coef, cudaStream_t st);
int sefd_loss_bwd_launch(const float* est, const float* tgt, const float* coef, const float* gout, float* dest,
                         int B, int L, cudaStream_t st);
