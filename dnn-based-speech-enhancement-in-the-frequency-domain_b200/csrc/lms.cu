// LMS perceptual loss (log-mel-spectrum distance at three mel scales) of the reference:
//   get_array_lms_loss / perceptual_distance / perceptual_transform / melFilterBank, tools_for_loss.py:111-249,
//   called from DCCRN.loss (models.py:305-312) on  clean_mags = sqrt(Re^2 + Im^2 + 1e-7) of STFT(target)  and
//   est_mags = sqrt(out_real^2 + out_imag^2 + 1e-7)  of the masked spectrum.
// Arithmetic restated (per batch item; x is the [257][T] magnitude array):
//   rows      = x.view(-1, 257)         -> row r holds the flat elements [257 r, 257 r + 257) of the [257][T] array
//                                          (a reshape, not a transpose: tools_for_loss.py:205 - reproduced as is)
//   P_s[r][m] = (1/512) sum_k rows[r][k] F_s[k][m],  F_s = melFilterBank(M_s, 512)^T,  M_s in {16, 32, 64}
//   L_s[r][m] = log(P_s[r][m] + 1e-7)
//   d_s[r]    = sqrt(mean_m (Lpred - Ltrue)^2 + 1e-7);   loss = mean_b mean_s mean_r d_s[r]
// One CTA (128 threads) per (item, row): thread m < 112 owns one mel coefficient of one scale.  The dense 257 x 112
// filter matrix (the three melFilterBank(M, 512) matrices side by side, tools_for_loss.py:144-188) is an INPUT: the
// host side builds it with numpy exactly as the reference does (its float32 rounding of the band edges depends on the
// numpy version, like the reference's own) and passes it as F [257][112] and its transpose Ft [112][257].
#include <string.h>

#include "../../include/sefd.h"
#include "common.cuh"
#include "prof.cuh"

namespace {

constexpr int NB = 257, NM = 112, FFT = 512;      // bins, mel coefficients of the three scales (16 + 32 + 64)

__device__ __forceinline__ int scale_of(int m, int& M, int& m0) {
    if (m < 16) { M = 16; m0 = 0; return 0; }
    if (m < 48) { M = 32; m0 = 16; return 1; }
    M = 64; m0 = 48; return 2;
}

struct LmsParams {
    const float *est_real, *est_imag;   // [B][257][T] masked spectrum (flat per item)
    const float* clean_spec;            // [B][257][T][2] STFT of the target
    const float *F, *Ft;
    int B, T;
    int mags;                           // 1: est_real / clean_spec already hold magnitudes [B][257][T] (est_imag unused)
    double* acc;                        // [1] sum over (item, scale, row) of d_s[r]
    const float* gout;                  // backward: upstream gradient (1 float) or nullptr
    float *d_real, *d_imag;             // backward outputs [B][257][T]
};

// shared per-row work: magnitudes into smem, the 112 log-mel values of true and pred, d_s[r]
__device__ __forceinline__ void lms_row(const LmsParams& p, int b, int r, float* xt, float* xp, float* red, float& lt, float& lp,
                                        float& pt_, float& pp_, float d[3]) {
    const int tid = threadIdx.x;
    const long long base = ((long long)b * NB * p.T) + (long long)NB * r;
    for (int k = tid; k < NB; k += 128) {
        if (p.mags) {
            xp[k] = __ldg(p.est_real + base + k);
            xt[k] = __ldg(p.clean_spec + base + k);
        } else {
            const float re = __ldg(p.est_real + base + k), im = __ldg(p.est_imag + base + k);
            xp[k] = sqrtf(re * re + im * im + 1e-7f);
            const float2 c = __ldg(reinterpret_cast<const float2*>(p.clean_spec) + base + k);
            xt[k] = sqrtf(c.x * c.x + c.y * c.y + 1e-7f);
        }
    }
    __syncthreads();
    float st = 0.f, sp = 0.f;
    if (tid < NM) {
        for (int k = 0; k < NB; ++k) {
            const float f = __ldg(p.F + k * NM + tid);
            st = fmaf(xt[k], f, st);
            sp = fmaf(xp[k], f, sp);
        }
    }
    pt_ = st * (1.f / FFT);
    pp_ = sp * (1.f / FFT);
    lt = logf(pt_ + 1e-7f);
    lp = logf(pp_ + 1e-7f);
    const float e = tid < NM ? (lp - lt) * (lp - lt) : 0.f;
    red[tid] = e;
    __syncthreads();
    if (tid < 3) {
        const int m0 = tid == 0 ? 0 : (tid == 1 ? 16 : 48), M = 16 << tid;
        float s = 0.f;
        for (int m = 0; m < M; ++m) s += red[m0 + m];
        red[128 + tid] = sqrtf(s / M + 1e-7f);
    }
    __syncthreads();
    d[0] = red[128]; d[1] = red[129]; d[2] = red[130];
}

__global__ void __launch_bounds__(128) lms_fwd_kernel(const LmsParams p) {
    __shared__ float xt[NB + 3], xp[NB + 3], red[132];
    const int b = blockIdx.y, r = blockIdx.x;
    float lt, lp, pt_, pp_, d[3];
    lms_row(p, b, r, xt, xp, red, lt, lp, pt_, pp_, d);
    if (threadIdx.x == 0) atomicAdd(p.acc, (double)d[0] + (double)d[1] + (double)d[2]);
}

__global__ void lms_finalize_kernel(const double* acc, int B, int T, float* loss) {
    loss[0] = (float)(acc[0] / (3.0 * (double)T * (double)B));
}

__global__ void __launch_bounds__(128) lms_bwd_kernel(const LmsParams p) {
    __shared__ float xt[NB + 3], xp[NB + 3], red[132], coef[NM];
    const int b = blockIdx.y, r = blockIdx.x, tid = threadIdx.x;
    float lt, lp, pt_, pp_, d[3];
    lms_row(p, b, r, xt, xp, red, lt, lp, pt_, pp_, d);
    const float go = (p.gout ? p.gout[0] : 1.f) / (3.f * (float)p.T * (float)p.B);
    if (tid < NM) {
        int M, m0;
        const int s = scale_of(tid, M, m0);
        // d loss / d P_pred[r][m] = go / d_s * (lp - lt) / M / (P + 1e-7);  P = x . F / 512
        coef[tid] = go / d[s] * (lp - lt) / (float)M / (pp_ + 1e-7f) * (1.f / FFT);
    }
    __syncthreads();
    const long long base = ((long long)b * NB * p.T) + (long long)NB * r;
    for (int k = tid; k < NB; k += 128) {
        float g = 0.f;
        for (int m = 0; m < NM; ++m) g = fmaf(coef[m], __ldg(p.Ft + m * NB + k), g);
        if (p.mags) {
            p.d_real[base + k] = g;
        } else {                        // est_mag = sqrt(re^2 + im^2 + 1e-7)
            const float re = __ldg(p.est_real + base + k), im = __ldg(p.est_imag + base + k);
            const float inv = g / xp[k];
            p.d_real[base + k] = inv * re;
            p.d_imag[base + k] = inv * im;
        }
    }
}

}  // namespace

extern "C" {

// scratch: 1 double.  loss: 1 float.
int sefd_lms_forward(const float* est_real, const float* est_imag, const float* clean_spec, const float* F, int B, int T,
                     int inputs_are_mags, double* scratch, float* loss, void* stream) {
    SEFD_REQUIRE(est_real && (est_imag || inputs_are_mags) && clean_spec && F && scratch && loss && B > 0 && T > 0,
                 "lms_forward: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    LmsParams p;
    memset(&p, 0, sizeof(p));
    p.F = F;
    p.est_real = est_real; p.est_imag = est_imag; p.clean_spec = clean_spec; p.B = B; p.T = T; p.acc = scratch;
    p.mags = inputs_are_mags;
    cudaMemsetAsync(scratch, 0, sizeof(double), st);
    SefdProfScope prof(SEFD_PROF_STFT, 0, 16.0 * B * NB * T, st);
    lms_fwd_kernel<<<dim3(T, B), 128, 0, st>>>(p);
    SEFD_TRY(sefd_check_launch("lms_fwd"));
    lms_finalize_kernel<<<1, 1, 0, st>>>(scratch, B, T, loss);
    return sefd_check_launch("lms_finalize");
}

// d_real / d_imag [B][257][T]: gradient of the loss with respect to the masked spectrum; gout (1 float, device) may be NULL
int sefd_lms_backward(const float* est_real, const float* est_imag, const float* clean_spec, const float* F, const float* Ft,
                      const float* gout, int B, int T, int inputs_are_mags, float* d_real, float* d_imag, void* stream) {
    SEFD_REQUIRE(est_real && clean_spec && F && Ft && d_real && B > 0 && T > 0 && (inputs_are_mags || (est_imag && d_imag)),
                 "lms_backward: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    LmsParams p;
    memset(&p, 0, sizeof(p));
    p.F = F; p.Ft = Ft;
    p.est_real = est_real; p.est_imag = est_imag; p.clean_spec = clean_spec; p.B = B; p.T = T;
    p.gout = gout; p.d_real = d_real; p.d_imag = d_imag; p.mags = inputs_are_mags;
    SefdProfScope prof(SEFD_PROF_STFT, 0, 24.0 * B * NB * T, st);
    lms_bwd_kernel<<<dim3(T, B), 128, 0, st>>>(p);
    return sefd_check_launch("lms_bwd");
}

}  // extern "C"
