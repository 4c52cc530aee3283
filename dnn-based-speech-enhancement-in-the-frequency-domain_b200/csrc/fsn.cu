// FullSubNet feature / target side of trainer.fullsubnet_train (trainer.py:97-104), SURVEY.md 8 a14:
//   tools.stft                          tools_for_model.py:628-648  torch.stft(y, 512, hop 300, win 400, hann_window(400)),
//                                       centred frames, reflect padding, window zero-padded to 512 (56 | 400 | 56)
//   tools.mag_phase                     tools_for_model.py:682-683
//   tools.build_complex_ideal_ratio_mask / compress_cIRM / decompress_cIRM   tools_for_model.py:686-723
// The train loop needs |STFT(noisy)| and cIRM(noisy, clean): one fused kernel rides the noisy frame on the real part and
// the clean frame on the imaginary part of ONE 512-point complex FFT (fft512.cuh), un-mixes the two spectra and writes the
// magnitude and the compressed mask - both waveforms are read once, no complex spectrum goes to HBM
// (algorithmic bytes per frame: 2 * 300 * 4 in, 257 * 12 out).  The FullSubNet model itself is not built yet.
#include <string.h>

#include "../../include/sefd.h"
#include "common.cuh"
#include "fft512.cuh"
#include "prof.cuh"

namespace {

constexpr int NF = 512, NB = 257, FHOP = 300, FWIN = 400, LPAD = (NF - FWIN) / 2, FPC = 16;   // frames per CTA
constexpr float CIRM_EPS = 1.1920928955078125e-07f;      // np.finfo(np.float32).eps

struct FsnSmem {
    float2 fft[4][NF];
    float2 tw[NF];
    float win[NF];
    float mag[NB][FPC + 1];
    float2 msk[NB][FPC + 1];
};

__device__ __forceinline__ void fsn_tables(float2* tw, float* win) {
    for (int j = threadIdx.x; j < NF; j += blockDim.x) {
        float s, c;
        sincospif((float)j / 256.f, &s, &c);
        tw[j] = make_float2(c, -s);
        const int m = j - LPAD;                           // periodic Hann(400) centred in the 512-sample frame
        win[j] = (m >= 0 && m < FWIN) ? 0.5f - 0.5f * cospif(2.f * (float)m / (float)FWIN) : 0.f;
    }
}
__device__ __forceinline__ int reflect(int k, int L) {   // F.pad(mode="reflect"): no edge repeat
    if (k < 0) k = -k;
    if (k >= L) k = 2 * (L - 1) - k;
    return k;
}
__device__ __forceinline__ float compress(float m) {      // compress_cIRM, K = 10, C = 0.1
    m = m <= -100.f ? -100.f : m;
    const float e = expf(-0.1f * m);
    return 10.f * (1.f - e) / (1.f + e);
}

// MODE 0: features (a = noisy, b = clean) -> mag [B][257][T], cirm [B][257][T][2]
// MODE 1: plain STFT of a (two frames per transform) -> spec [B][257][T][2]
template <int MODE>
__global__ void __launch_bounds__(256) fsn_stft_kernel(const float* __restrict__ a, const float* __restrict__ b, int L, int T,
                                                       float* __restrict__ mag, float* __restrict__ out2) {
    extern __shared__ __align__(16) unsigned char raw[];
    FsnSmem& sm = *reinterpret_cast<FsnSmem*>(raw);
    const int bi = blockIdx.y, t0 = blockIdx.x * FPC;
    const int tid = threadIdx.x, g = tid >> 6, t64 = tid & 63;
    const float* wa = a + (size_t)bi * L;
    const float* wb = MODE == 0 ? b + (size_t)bi * L : nullptr;
    fsn_tables(sm.tw, sm.win);
    __syncthreads();
    constexpr int PER = MODE == 0 ? 1 : 2;                 // frames per transform
    for (int r = 0; r < FPC / (4 * PER); ++r) {
        const int f0 = (4 * r + g) * PER, ta = t0 + f0, tb = ta + 1;
        float2* s = sm.fft[g];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = t64 + 64 * i;
            float x = 0.f, y = 0.f;
            const float w = sm.win[n];
            if (w != 0.f) {
                if (ta < T) x = __ldg(wa + reflect(ta * FHOP + n - NF / 2, L));
                if (MODE == 0) { if (ta < T) y = __ldg(wb + reflect(ta * FHOP + n - NF / 2, L)); }
                else if (tb < T) y = __ldg(wa + reflect(tb * FHOP + n - NF / 2, L));
            }
            s[fft_at(n)] = make_float2(w * x, w * y);
        }
        __syncthreads();
        fft512_cta<false>(s, sm.tw, t64);
        for (int k = t64; k <= 256; k += 64) {
            const float2 z = s[fft_at(k)], zc = s[fft_at((NF - k) & (NF - 1))];
            const float2 xa = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));      // spectrum of the real part
            const float2 xb = make_float2(0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));     // spectrum of the imaginary part
            if (MODE == 0) {
                const float p = xa.x * xa.x + xa.y * xa.y;
                sm.mag[k][f0] = sqrtf(p);
                const float den = p + CIRM_EPS;
                sm.msk[k][f0] = make_float2(compress((xa.x * xb.x + xa.y * xb.y) / den),
                                            compress((xa.x * xb.y - xa.y * xb.x) / den));
            } else {
                sm.msk[k][f0] = xa;
                sm.msk[k][f0 + 1] = xb;
            }
        }
        __syncthreads();
    }
    const int nf = min(FPC, T - t0);
    float2* o2 = reinterpret_cast<float2*>(out2) + (size_t)bi * NB * T;
    for (int e = tid; e < NB * FPC; e += 256) {
        const int k = e / FPC, f = e % FPC;
        if (f < nf) {
            o2[(size_t)k * T + t0 + f] = sm.msk[k][f];
            if (MODE == 0) mag[((size_t)bi * NB + k) * T + t0 + f] = sm.mag[k][f];
        }
    }
}

__global__ void fsn_mag_phase_kernel(const float2* __restrict__ spec, long long n, float* __restrict__ mag, float* __restrict__ phase) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 z = spec[i];
        mag[i] = hypotf(z.x, z.y);
        if (phase) phase[i] = atan2f(z.y, z.x);
    }
}
__global__ void fsn_cirm_kernel(const float2* __restrict__ noisy, const float2* __restrict__ clean, long long n, float2* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 x = noisy[i], c = clean[i];
        const float den = x.x * x.x + x.y * x.y + CIRM_EPS;
        out[i] = make_float2(compress((x.x * c.x + x.y * c.y) / den), compress((x.x * c.y - x.y * c.x) / den));
    }
}
__global__ void fsn_compress_kernel(const float* __restrict__ m, long long n, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = compress(m[i]);
}
__global__ void fsn_decompress_kernel(const float* __restrict__ m, long long n, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = m[i];
        v = v >= 9.9f ? 9.9f : (v <= -9.9f ? -9.9f : v);                 // decompress_cIRM, K = 10, limit = 9.9
        out[i] = -10.f * logf((10.f - v) / (10.f + v));
    }
}

// tools.istft (tools_for_model.py:651-679) = torch.istft(features, 512, 300, 400, hann_window(400), length): per frame
// irfft_512 (1/512, imaginary parts of DC / Nyquist ignored), times the centred window, overlap-add at hop 300, divided by the
// overlap-added squared window, trimmed by n_fft / 2 at the front.  Output-centric: one CTA owns IH hops of output samples
// and transforms the IH + 2 frames that touch them (two frames per complex inverse FFT, Hermitian packing), so the
// overlap-add needs no atomics and is deterministic.  MAGPH: the input is (mag, phase) instead of a complex spectrum.
constexpr int IH = 8;
struct IfsnSmem {
    float2 fft[4][NF];
    float2 tw[NF];
    float win[NF];
    float fr[IH + 4][NF];
};

template <bool MAGPH>
__global__ void __launch_bounds__(256) fsn_istft_kernel(const float* __restrict__ in0, const float* __restrict__ in1, int T,
                                                        int len, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char raw[];
    IfsnSmem& sm = *reinterpret_cast<IfsnSmem*>(raw);
    const int bi = blockIdx.y, h0 = blockIdx.x * IH;          // hops h0 .. h0 + IH - 1 of the padded signal
    const int tid = threadIdx.x, g = tid >> 6, t64 = tid & 63;
    fsn_tables(sm.tw, sm.win);
    __syncthreads();
    const int fbase = h0 - 1;                                  // frames fbase .. fbase + IH + 2 (IH + 3 would be unused)
    for (int r = 0; r < (IH + 4) / 8 + ((IH + 4) % 8 ? 1 : 0); ++r) {
        const int j0 = (4 * r + g) * 2;                        // local frame pair (j0, j0 + 1)
        const int fa = fbase + j0, fb = fa + 1;
        const bool va = j0 < IH + 4 && fa >= 0 && fa < T, vb = j0 + 1 < IH + 4 && fb >= 0 && fb < T;
        float2* s = sm.fft[g];
        for (int k = t64; k <= 256; k += 64) {
            float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
            const size_t base = ((size_t)bi * NB + k) * T;
            if (MAGPH) {
                if (va) { float sn, cs; sincosf(in1[base + fa], &sn, &cs); const float m = in0[base + fa]; a = make_float2(m * cs, m * sn); }
                if (vb) { float sn, cs; sincosf(in1[base + fb], &sn, &cs); const float m = in0[base + fb]; b = make_float2(m * cs, m * sn); }
            } else {
                const float2* sp = reinterpret_cast<const float2*>(in0);
                if (va) a = sp[base + fa];
                if (vb) b = sp[base + fb];
            }
            if (k == 0 || k == 256) {
                s[fft_at(k)] = make_float2(a.x, b.x);
            } else {
                s[fft_at(k)] = make_float2(a.x - b.y, a.y + b.x);
                s[fft_at(NF - k)] = make_float2(a.x + b.y, -a.y + b.x);
            }
        }
        __syncthreads();
        fft512_cta<true>(s, sm.tw, t64);
        if (j0 < IH + 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = t64 + 64 * i;
                const float w = sm.win[n] * (1.f / NF);
                sm.fr[j0][n] = va ? w * s[fft_at(n)].x : 0.f;
                if (j0 + 1 < IH + 4) sm.fr[j0 + 1][n] = vb ? w * s[fft_at(n)].y : 0.f;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < IH * FHOP; i += 256) {
        const int p = h0 * FHOP + i;                           // padded coordinate
        const int o = p - NF / 2;
        if (o < 0 || o >= len) continue;
        const int f1 = p / FHOP, f2 = f1 - 1;                  // the (at most) two frames covering p
        float v = 0.f, env = 0.f;
        if (f1 < T) { const int n = p - f1 * FHOP; v += sm.fr[f1 - fbase][n]; env += sm.win[n] * sm.win[n]; }
        if (f2 >= 0 && f2 < T) {
            const int n = p - f2 * FHOP;
            if (n < NF) { v += sm.fr[f2 - fbase][n]; env += sm.win[n] * sm.win[n]; }
        }
        out[(size_t)bi * len + o] = env > 1e-11f ? v / env : 0.f;
    }
}

template <int MODE>
int launch_stft(const float* a, const float* b, int B, int L, float* mag, float* out2, cudaStream_t st) {
    const int T = L / FHOP + 1;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(fsn_stft_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FsnSmem));
        attr = true;
    }
    SefdProfScope prof(SEFD_PROF_STFT, 0, 4.0 * B * ((MODE == 0 ? 2.0 : 1.0) * L + (MODE == 0 ? 3.0 : 2.0) * NB * T), st);
    fsn_stft_kernel<MODE><<<dim3((T + FPC - 1) / FPC, B), 256, sizeof(FsnSmem), st>>>(a, b, L, T, mag, out2);
    return sefd_check_launch("fsn_stft");
}
inline unsigned ew_blocks(long long n) { return (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8); }

}  // namespace

extern "C" {

int sefd_fsn_frames(int L) { return L > NF / 2 ? L / FHOP + 1 : 0; }

/* noisy, clean [B][L] -> noisy_mag [B][257][T], cirm [B][257][T][2], T = sefd_fsn_frames(L) */
int sefd_fsn_features(const float* noisy, const float* clean, int B, int L, float* noisy_mag, float* cirm, void* stream) {
    SEFD_REQUIRE(noisy && clean && noisy_mag && cirm && B > 0, "fsn_features: bad argument");
    SEFD_REQUIRE(L > NF / 2, "fsn_features: reflect padding needs more than %d samples (got %d)", NF / 2, L);
    return launch_stft<0>(noisy, clean, B, L, noisy_mag, cirm, (cudaStream_t)stream);
}
/* wav [B][L] -> spec [B][257][T][2] (torch.stft(..., return_complex=True) viewed as real) */
int sefd_fsn_stft(const float* wav, int B, int L, float* spec, void* stream) {
    SEFD_REQUIRE(wav && spec && B > 0, "fsn_stft: bad argument");
    SEFD_REQUIRE(L > NF / 2, "fsn_stft: reflect padding needs more than %d samples (got %d)", NF / 2, L);
    return launch_stft<1>(wav, nullptr, B, L, nullptr, spec, (cudaStream_t)stream);
}
/* spec [n][2] -> mag [n], phase [n] (phase may be NULL) */
int sefd_fsn_mag_phase(const float* spec, long long n, float* mag, float* phase, void* stream) {
    SEFD_REQUIRE(spec && mag && n > 0, "fsn_mag_phase: bad argument");
    sefd_absorb_stale_error();
    fsn_mag_phase_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec), n, mag, phase);
    return sefd_check_launch("fsn_mag_phase");
}
/* noisy_spec, clean_spec [n][2] -> compressed complex ideal ratio mask [n][2] */
int sefd_fsn_cirm(const float* noisy_spec, const float* clean_spec, long long n, float* cirm, void* stream) {
    SEFD_REQUIRE(noisy_spec && clean_spec && cirm && n > 0, "fsn_cirm: bad argument");
    sefd_absorb_stale_error();
    fsn_cirm_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(noisy_spec),
                                                                    reinterpret_cast<const float2*>(clean_spec), n,
                                                                    reinterpret_cast<float2*>(cirm));
    return sefd_check_launch("fsn_cirm");
}
/* spec [B][257][T][2] (or mag, phase [B][257][T] when phase != NULL) -> wav [B][len] */
int sefd_fsn_istft(const float* spec_or_mag, const float* phase, int B, int T, int len, float* wav, void* stream) {
    SEFD_REQUIRE(spec_or_mag && wav && B > 0 && T > 0 && len > 0, "fsn_istft: bad argument");
    // samples no frame covers (len beyond 300 (T - 1) + 256) come out as zeros, like torch.istft's right padding
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(fsn_istft_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IfsnSmem));
        cudaFuncSetAttribute(fsn_istft_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IfsnSmem));
        attr = true;
    }
    const int hops = (NF / 2 + len + FHOP - 1) / FHOP;
    dim3 grid((hops + IH - 1) / IH, B);
    SefdProfScope prof(SEFD_PROF_STFT, 0, 4.0 * B * (2.0 * NB * T + len), st);
    if (phase) fsn_istft_kernel<true><<<grid, 256, sizeof(IfsnSmem), st>>>(spec_or_mag, phase, T, len, wav);
    else fsn_istft_kernel<false><<<grid, 256, sizeof(IfsnSmem), st>>>(spec_or_mag, nullptr, T, len, wav);
    return sefd_check_launch("fsn_istft");
}
int sefd_fsn_compress_cirm(const float* mask, long long n, float* out, void* stream) {
    SEFD_REQUIRE(mask && out && n > 0, "fsn_compress_cirm: bad argument");
    sefd_absorb_stale_error();
    fsn_compress_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(mask, n, out);
    return sefd_check_launch("fsn_compress_cirm");
}
int sefd_fsn_decompress_cirm(const float* mask, long long n, float* out, void* stream) {
    SEFD_REQUIRE(mask && out && n > 0, "fsn_decompress_cirm: bad argument");
    sefd_absorb_stale_error();
    fsn_decompress_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(mask, n, out);
    return sefd_check_launch("fsn_decompress_cirm");
}

}  // extern "C"
