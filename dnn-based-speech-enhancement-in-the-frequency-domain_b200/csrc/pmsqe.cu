// PMSQE perceptual loss of the reference's perceptual train step:
//   tools_for_loss.py:255-256  pmsqe_stft = Encoder(STFTFB(kernel_size=512, n_filters=512, stride=256)),
//                              pmsqe_loss = PITLossWrapper(SingleSrcPMSQE(), pit_from='pw_pt')
//   tools_for_loss.py:259-269  get_array_pmsqe_loss: wav.view(N, -1, 16000) -> mag(stft(.)) -> pmsqe_loss(est, clean)
// PARITY UNPINNED: the arithmetic lives in asteroid / asteroid_filterbanks, which are neither in the reference tree nor in
// this image; what is restated here is the published algorithm (Martin-Donas et al., IEEE SPL 2018) as asteroid implements
// it, see oracle/pmsqe_oracle.py.  Every perceptual table (Bark matrix, thresholds, Zwicker powers, band widths, SLL mask) is
// an INPUT, so a site with asteroid installed can pass asteroid's own tensors.
//
// Per utterance n the three 1-second chunks are "sources" s; est chunk i is scored against clean chunk j for all S x S
// pairs, the permutation with the lowest mean wins (PIT), the batch mean is the loss.
//   K1 stft_mag   : one CTA per (chunk, est|clean): 61 frames x 512-point FFT (two real frames per complex transform),
//                   |X| = sqrt(re^2 + im^2 + 1e-8), sum of SLL-masked magnitudes
//   K2 bark       : R[t][k] = sum_f |X|[t][f] M[f][k]   (the 1e7 * Sp / mean scaling is a scalar per chunk, applied later)
//   K3 pair       : frequency / gain equalisation, Zwicker loudness, symmetric + asymmetric disturbance, frame mean
//   K4 pit        : best permutation per utterance, batch mean
//   K5 pair_bwd   : K3 recomputed for the winning pairs + its hand-written reverse sweep -> dR, d(mean)
//   K6 mag_istft  : dR -> d|X| -> d spectrum -> inverse FFT, window, overlap-add -> d est_wav
#include <string.h>

#include "../../include/sefd.h"
#include "common.cuh"
#include "fft512.cuh"
#include "prof.cuh"

namespace {

constexpr int NB = 257, NK = 49, PF = 61, SEG = 16000, NF = 512, PHOP = 256, MAXS = 4;
constexpr int T_BARK = 0, T_THR = NB * NK, T_ZW = T_THR + NK, T_WID = T_ZW + NK, T_MASK = T_WID + NK, T_END = T_MASK + NB;
constexpr float SP = 6.910853e-006f, SL = 1.866055e-001f, ALPHA = 0.1f, BETA = 0.0309f, PEPS = 1e-8f;

struct Ws {                       // workspace carve (float offsets)
    size_t spec, magE, magC, RE, RC, dR, msum, coef, pw, perm, acc, end;
};
__host__ __device__ inline Ws carve(int NS, int N, int S) {
    Ws w;
    size_t c = 0;
    w.spec = c; c += (size_t)NS * PF * NB * 2;
    w.magE = c; c += (size_t)NS * PF * NB;
    w.magC = c; c += (size_t)NS * PF * NB;
    w.RE = c; c += (size_t)NS * PF * NK;
    w.RC = c; c += (size_t)NS * PF * NK;
    w.dR = c; c += (size_t)NS * PF * NK;
    w.msum = c; c += (size_t)2 * NS;
    w.coef = c; c += (size_t)NS;
    w.pw = c; c += (size_t)N * S * S;
    w.perm = c; c += (size_t)N * MAXS;
    c = (c + 1) & ~(size_t)1;
    w.acc = c; c += 2;           // one double
    w.end = c;
    return w;
}

__device__ __forceinline__ void init_tw(float2* tw, float* win) {
    for (int j = threadIdx.x; j < NF; j += blockDim.x) {
        float s, c;
        sincospif((float)j / 256.f, &s, &c);
        tw[j] = make_float2(c, -s);                       // e^{-2 pi i j / 512}
        win[j] = sinpif((float)j / 512.f);                // sqrt(hanning(513)[j]) = sin(pi j / 512)
    }
}
__device__ __forceinline__ float bin_scale(int k) {       // STFTFB: 1/16, DC and Nyquist real rows / sqrt(2)
    return (k == 0 || k == NF / 2) ? 0.0625f * 0.70710678118654752440f : 0.0625f;
}

// ---------------------------------------------------------------------------------------------------------------
struct StftSmem {
    float2 fft[4][NF];
    float2 tw[NF];
    float win[NF];
    float red[8];
};

__global__ void __launch_bounds__(256) pmsqe_stft_kernel(const float* __restrict__ est, const float* __restrict__ clean,
                                                         int L, int S, const float* __restrict__ tables, float* ws) {
    __shared__ StftSmem sm;
    const int seg = blockIdx.x, which = blockIdx.y, NS = gridDim.x;
    const Ws W = carve(NS, NS / S, S);
    const int tid = threadIdx.x, g = tid >> 6, t64 = tid & 63;
    const float* w = (which ? clean : est) + (size_t)(seg / S) * L + (size_t)(seg % S) * SEG;
    float* mag = ws + (which ? W.magC : W.magE) + (size_t)seg * PF * NB;
    float2* spec = reinterpret_cast<float2*>(ws + W.spec) + (size_t)seg * PF * NB;
    const float* mask = tables + T_MASK;
    init_tw(sm.tw, sm.win);
    __syncthreads();
    float sum = 0.f;
    for (int r = 0; r < 8; ++r) {
        const int fa = 2 * (4 * r + g), fb = fa + 1;
        float2* s = sm.fft[g];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = t64 + 64 * i;
            const float xa = fa < PF ? __ldg(w + PHOP * fa + n) : 0.f, xb = fb < PF ? __ldg(w + PHOP * fb + n) : 0.f;
            s[fft_at(n)] = make_float2(sm.win[n] * xa, sm.win[n] * xb);
        }
        __syncthreads();
        fft512_cta<false>(s, sm.tw, t64);
        for (int k = t64; k <= 256; k += 64) {            // unpack the two real transforms
            const float2 z = s[fft_at(k)], zc = s[fft_at((NF - k) & (NF - 1))];
            const float sc = bin_scale(k);
            float2 xa = make_float2(sc * 0.5f * (z.x + zc.x), sc * 0.5f * (z.y - zc.y));
            float2 xb = make_float2(sc * 0.5f * (z.y + zc.y), -sc * 0.5f * (z.x - zc.x));
            if (k == 0 || k == 256) xa.y = xb.y = 0.f;    // the filter rows are exactly zero there
            const float mk = __ldg(mask + k);
            if (fa < PF) {
                const float m = sqrtf(xa.x * xa.x + xa.y * xa.y + PEPS);
                mag[fa * NB + k] = m;
                if (!which) spec[fa * NB + k] = xa;
                sum = fmaf(m, mk, sum);
            }
            if (fb < PF) {
                const float m = sqrtf(xb.x * xb.x + xb.y * xb.y + PEPS);
                mag[fb * NB + k] = m;
                if (!which) spec[fb * NB + k] = xb;
                sum = fmaf(m, mk, sum);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0) sm.red[tid >> 5] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += sm.red[i];
        ws[W.msum + (size_t)which * NS + seg] = t;        // sum over frames and bins of mask * |X|
    }
}

__global__ void __launch_bounds__(256) pmsqe_bark_kernel(int S, const float* __restrict__ tables, float* ws) {
    const int seg = blockIdx.x, which = blockIdx.y, NS = gridDim.x;
    const Ws W = carve(NS, NS / S, S);
    const float* mag = ws + (which ? W.magC : W.magE) + (size_t)seg * PF * NB;
    float* R = ws + (which ? W.RC : W.RE) + (size_t)seg * PF * NK;
    const float* M = tables + T_BARK;
    for (int e = threadIdx.x; e < PF * NK; e += 256) {
        const int t = e / NK, k = e % NK;
        float a = 0.f;
        for (int f = 0; f < NB; ++f) a = fmaf(__ldg(mag + t * NB + f), __ldg(M + f * NK + k), a);
        R[e] = a;
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct PairSmem {
    float Br[PF * NK], Bd[PF * NK], D[PF * NK];
    float thr[NK], zw[NK], wid[NK], at[NK], eq[NK], eraw[NK], sdv[NK], dsd[NK];
    float ns[PF + 3], apr[PF + 3], fval[PF + 3];
    float red[64];
};

__device__ __forceinline__ float loud(float b, float thr, float zw, float at) {
    return b < thr ? 0.f : SL * at * (powf(0.5f + 0.5f * b / thr, zw) - 1.f);
}

// one CTA (64 threads) scores est chunk i against clean chunk j of utterance n.  BWD: only the winning pairs, and the
// reverse sweep leaves dR (gradient w.r.t. the unscaled Bark spectrum of the est chunk) and the d(mean) coefficient.
template <bool BWD>
__global__ void __launch_bounds__(64) pmsqe_pair_kernel(int N, int S, const float* __restrict__ tables, float* ws,
                                                        const float* __restrict__ gout) {
    extern __shared__ __align__(16) unsigned char raw[];
    PairSmem& sm = *reinterpret_cast<PairSmem*>(raw);
    const int NS = N * S, n = blockIdx.x, tid = threadIdx.x;
    const Ws W = carve(NS, N, S);
    int i, j;
    if (BWD) {
        i = blockIdx.y;
        j = reinterpret_cast<const int*>(ws + W.perm)[n * MAXS + i];
    } else {
        i = blockIdx.y / S;
        j = blockIdx.y % S;
    }
    const int segE = n * S + i, segC = n * S + j;
    const float norm = 1e7f * SP * (float)(PF * NB);
    const float msE = ws[W.msum + segE], msC = ws[W.msum + NS + segC];
    const float cE = norm / msE, cC = norm / msC;               // 1e7 * Sp / mean(mask * |X|)
    if (tid < NK) {
        sm.thr[tid] = tables[T_THR + tid];
        sm.zw[tid] = tables[T_ZW + tid];
        sm.wid[tid] = tables[T_WID + tid];
        sm.at[tid] = powf(sm.thr[tid] / 0.5f, sm.zw[tid]);
    }
    const float* RE = ws + W.RE + (size_t)segE * PF * NK;
    const float* RC = ws + W.RC + (size_t)segC * PF * NK;
    for (int e = tid; e < PF * NK; e += 64) {
        sm.Br[e] = cC * RC[e];
        sm.Bd[e] = cE * RE[e];
    }
    __syncthreads();
    float sqrtW = 0.f;
    for (int k = 0; k < NK; ++k) sqrtW += sm.wid[k];
    sqrtW = sqrtf(sqrtW);
    // audible power of the reference: x100 threshold decides speech-active frames, x1 is the gain / weight term
    if (tid < PF) {
        float a100 = 0.f, a1 = 0.f;
        for (int k = 0; k < NK; ++k) {
            const float b = sm.Br[tid * NK + k];
            if (b > sm.thr[k] * 100.f) a100 += b;
            if (b > sm.thr[k]) a1 += b;
        }
        sm.ns[tid] = a100 >= 1e7f ? 1.f : 0.f;
        sm.apr[tid] = a1;
    }
    __syncthreads();
    // frequency equalisation: ratio of the thresholded band powers over the speech-active frames
    if (tid < NK) {
        float sr = 0.f, sd = 0.f;
        const float th = sm.thr[tid] * 100.f;
        for (int t = 0; t < PF; ++t) {
            const float b = sm.Br[t * NK + tid];
            if (sm.ns[t] != 0.f && b >= th) { sr += b; sd += sm.Bd[t * NK + tid]; }
        }
        const float e = (sr + 1000.f) / (sd + 1000.f);
        sm.eraw[tid] = e;
        sm.sdv[tid] = sd;
        sm.eq[tid] = fminf(fmaxf(e, 0.01f), 100.f);
    }
    __syncthreads();
    float graw = 0.f, gl = 0.f, apd = 0.f, acc2 = 0.f, acc1 = 0.f, wt = 1.f;
    if (tid < PF) {
        const int t = tid;
        for (int k = 0; k < NK; ++k) {
            const float b1 = sm.eq[k] * sm.Bd[t * NK + k];
            if (b1 > sm.thr[k]) apd += b1;
        }
        graw = (sm.apr[t] + 5e3f) / (apd + 5e3f);
        gl = fminf(fmaxf(graw, 3e-4f), 5.f);
        for (int k = 0; k < NK; ++k) {
            const float br = sm.Br[t * NK + k], b = gl * sm.eq[k] * sm.Bd[t * NK + k];
            const float lr = loud(br, sm.thr[k], sm.zw[k], sm.at[k]), ld = loud(b, sm.thr[k], sm.zw[k], sm.at[k]);
            const float sym = fmaxf(fabsf(ld - lr) - 0.25f * fminf(lr, ld), PEPS);
            const float asym = powf((b + 50.f) / (br + 50.f), 1.2f);
            const float af = asym < 3.f ? 0.f : fminf(asym, 12.f);
            const float sw = sym * sm.wid[k];
            acc2 += sw * sw + PEPS;
            acc1 += af * sw;
        }
        wt = powf((sm.apr[t] + 1e5f) / 1e7f, 0.04f);
        const float wd = fminf(sqrtf(acc2) * sqrtW / wt, 45.f), wda = fminf(acc1 / wt, 45.f);
        sm.fval[t] = ALPHA * wd + BETA * wda;
    }
    __syncthreads();
    if (!BWD) {
        if (tid == 0) {
            float s = 0.f;
            for (int t = 0; t < PF; ++t) s += sm.fval[t];
            ws[W.pw + ((size_t)n * S + i) * S + j] = s / PF;
        }
        return;
    }
    // ---- reverse sweep ----
    const float go = (gout ? gout[0] : 1.f) / ((float)N * S * PF);
    if (tid < PF) {
        const int t = tid;
        const float root = sqrtf(acc2);
        const float dd = (root * sqrtW / wt < 45.f) ? go * ALPHA / wt : 0.f;      // d / d d_frame
        const float dda = (acc1 / wt < 45.f) ? go * BETA / wt : 0.f;              // d / d da_frame
        float dg = 0.f;
        for (int k = 0; k < NK; ++k) {
            const float br = sm.Br[t * NK + k], b1 = sm.eq[k] * sm.Bd[t * NK + k], b = gl * b1;
            const float th = sm.thr[k], zw = sm.zw[k], wk = sm.wid[k];
            const float lr = loud(br, th, zw, sm.at[k]), ld = loud(b, th, zw, sm.at[k]);
            const float rm = fabsf(ld - lr) - 0.25f * fminf(lr, ld);
            const float sym = fmaxf(rm, PEPS);
            const float asym = powf((b + 50.f) / (br + 50.f), 1.2f);
            const float af = asym < 3.f ? 0.f : fminf(asym, 12.f);
            const float dsym = dd * sqrtW * sym * wk * wk / root + dda * wk * af;
            float db = 0.f;
            if (asym >= 3.f && asym < 12.f) db += dda * wk * sym * 1.2f * asym / (b + 50.f);
            if (rm > PEPS) {
                const float sgn = ld > lr ? 1.f : (ld < lr ? -1.f : 0.f);
                const float dld = dsym * sgn - (ld < lr ? 0.25f * dsym : 0.f);
                if (b >= th) db += dld * SL * sm.at[k] * zw * powf(0.5f + 0.5f * b / th, zw - 1.f) * 0.5f / th;
            }
            sm.D[t * NK + k] = db;                          // d / d (gain- and frequency-equalised Bark spectrum)
            dg += db * b1;
        }
        const float dapd = (graw > 3e-4f && graw < 5.f) ? -dg * graw / (apd + 5e3f) : 0.f;
        for (int k = 0; k < NK; ++k) {
            const float b1 = sm.eq[k] * sm.Bd[t * NK + k];
            sm.D[t * NK + k] = gl * sm.D[t * NK + k] + (b1 > sm.thr[k] ? dapd : 0.f);   // d / d (frequency-equalised)
        }
    }
    __syncthreads();
    if (tid < NK) {
        float deq = 0.f;
        for (int t = 0; t < PF; ++t) deq += sm.D[t * NK + tid] * sm.Bd[t * NK + tid];
        const float e = sm.eraw[tid];
        sm.dsd[tid] = (e > 0.01f && e < 100.f) ? -deq * e / (sm.sdv[tid] + 1000.f) : 0.f;
    }
    __syncthreads();
    float dc = 0.f;
    float* dR = ws + W.dR + (size_t)segE * PF * NK;
    for (int e = tid; e < PF * NK; e += 64) {
        const int t = e / NK, k = e % NK;
        float d = sm.eq[k] * sm.D[e];
        if (sm.ns[t] != 0.f && sm.Br[e] >= sm.thr[k] * 100.f) d += sm.dsd[k];
        dR[e] = cE * d;
        dc = fmaf(d, sm.Bd[e], dc);                         // Bd = cE * R  ->  d cE = sum d * R = sum d * Bd / cE
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dc += __shfl_xor_sync(0xffffffffu, dc, o);
    if ((tid & 31) == 0) sm.red[tid >> 5] = dc;
    __syncthreads();
    if (tid == 0) ws[W.coef + segE] = -(sm.red[0] + sm.red[1]) / msE;   // cE = norm / msum: d msum = -(d cE) cE / msum
}

__global__ void pmsqe_pit_kernel(int N, int S, float* ws, float* loss) {
    const Ws W = carve(N * S, N, S);
    int* perm_out = reinterpret_cast<int*>(ws + W.perm);
    __shared__ float red[256];
    float total = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float* pw = ws + W.pw + (size_t)n * S * S;
        int p[MAXS] = {0, 1, 2, 3}, best[MAXS] = {0, 1, 2, 3};
        float bestv = 3.0e38f;
        // lexicographic enumeration (itertools.permutations order); strict < keeps the first minimum like torch.min
        for (;;) {
            float v = 0.f;
            for (int i = 0; i < S; ++i) v += pw[i * S + p[i]];
            v /= S;
            if (v < bestv) { bestv = v; for (int i = 0; i < S; ++i) best[i] = p[i]; }
            int a = S - 2;
            while (a >= 0 && p[a] > p[a + 1]) --a;
            if (a < 0) break;
            int b = S - 1;
            while (p[b] < p[a]) --b;
            int tmp = p[a]; p[a] = p[b]; p[b] = tmp;
            for (int l = a + 1, r = S - 1; l < r; ++l, --r) { tmp = p[l]; p[l] = p[r]; p[r] = tmp; }
        }
        for (int i = 0; i < S; ++i) perm_out[n * MAXS + i] = best[i];
        total += bestv;
    }
    red[threadIdx.x] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)blockDim.x; ++i) s += red[i];
        loss[0] = (float)(s / N);
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct IstftSmem {
    float2 fft[4][NF];
    float2 tw[NF];
    float win[NF];
    float dR[PF * NK];
    float fr[PF][NF];
};

__global__ void __launch_bounds__(256) pmsqe_mag_istft_bwd_kernel(int L, int S, const float* __restrict__ tables,
                                                                  const float* __restrict__ ws, float* __restrict__ d_est) {
    extern __shared__ __align__(16) unsigned char raw[];
    IstftSmem& sm = *reinterpret_cast<IstftSmem*>(raw);
    const int seg = blockIdx.x, NS = gridDim.x;
    const Ws W = carve(NS, NS / S, S);
    const int tid = threadIdx.x, g = tid >> 6, t64 = tid & 63;
    const float* mag = ws + W.magE + (size_t)seg * PF * NB;
    const float2* spec = reinterpret_cast<const float2*>(ws + W.spec) + (size_t)seg * PF * NB;
    const float* M = tables + T_BARK;
    const float* mask = tables + T_MASK;
    const float coef = ws[W.coef + seg];
    init_tw(sm.tw, sm.win);
    for (int e = tid; e < PF * NK; e += 256) sm.dR[e] = ws[W.dR + (size_t)seg * PF * NK + e];
    __syncthreads();
    for (int r = 0; r < 8; ++r) {
        const int fa = 2 * (4 * r + g), fb = fa + 1;
        float2* s = sm.fft[g];
        for (int k = t64; k <= 256; k += 64) {
            float2 ga = make_float2(0.f, 0.f), gb = make_float2(0.f, 0.f);
            const float mk = coef * __ldg(mask + k), sc = bin_scale(k);
            float da = mk, db = mk;
            if (fa < PF) {
                for (int q = 0; q < NK; ++q) da = fmaf(sm.dR[fa * NK + q], __ldg(M + k * NK + q), da);
                const float2 x = spec[fa * NB + k];
                const float f = sc * da / mag[fa * NB + k];          // d|X| * X / |X|, then the filter scale
                ga = make_float2(f * x.x, f * x.y);
            }
            if (fb < PF) {
                for (int q = 0; q < NK; ++q) db = fmaf(sm.dR[fb * NK + q], __ldg(M + k * NK + q), db);
                const float2 x = spec[fb * NB + k];
                const float f = sc * db / mag[fb * NB + k];
                gb = make_float2(f * x.x, f * x.y);
            }
            // Re sum_{k=0}^{256} G_k e^{+i theta k n}  =  inverse transform of the Hermitian H: H_0 = Re G_0, H_256 = Re G_256,
            // H_k = G_k / 2, H_{512-k} = conj(G_k) / 2; two frames ride one complex transform as H_a + i H_b
            if (k == 0 || k == 256) {
                s[fft_at(k)] = make_float2(ga.x, gb.x);
            } else {
                ga.x *= 0.5f; ga.y *= 0.5f; gb.x *= 0.5f; gb.y *= 0.5f;
                s[fft_at(k)] = make_float2(ga.x - gb.y, ga.y + gb.x);
                s[fft_at(NF - k)] = make_float2(ga.x + gb.y, -ga.y + gb.x);
            }
        }
        __syncthreads();
        fft512_cta<true>(s, sm.tw, t64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = t64 + 64 * i;
            if (fa < PF) sm.fr[fa][n] = sm.win[n] * s[fft_at(n)].x;
            if (fb < PF) sm.fr[fb][n] = sm.win[n] * s[fft_at(n)].y;
        }
        __syncthreads();
    }
    float* out = d_est + (size_t)(seg / S) * L + (size_t)(seg % S) * SEG;
    for (int i = tid; i < SEG; i += 256) {                 // overlap-add in a fixed order (hop 256, two frames per sample)
        const int t = i / PHOP, n = i % PHOP;
        float v = 0.f;
        if (t < PF) v += sm.fr[t][n];
        if (t >= 1 && t - 1 < PF) v += sm.fr[t - 1][PHOP + n];
        out[i] = v;
    }
}

}  // namespace

extern "C" {

int sefd_pmsqe_table_floats(void) { return T_END; }

size_t sefd_pmsqe_workspace_bytes(int N, int L) {
    if (N <= 0 || L <= 0 || L % SEG != 0 || L / SEG > MAXS) return 0;
    const int S = L / SEG;
    return carve(N * S, N, S).end * sizeof(float);
}

int sefd_pmsqe_forward(const float* est_wav, const float* clean_wav, int N, int L, const float* tables, void* ws,
                       size_t ws_bytes, float* loss, void* stream) {
    SEFD_REQUIRE(est_wav && clean_wav && tables && ws && loss && N > 0, "pmsqe_forward: bad argument");
    SEFD_REQUIRE(L > 0 && L % SEG == 0 && L / SEG <= MAXS,
                 "pmsqe_forward: the waveform must be 1..%d whole seconds at 16 kHz (tools_for_loss.py:264), got %d samples", MAXS, L);
    const int S = L / SEG, NS = N * S;
    SEFD_REQUIRE(ws_bytes >= carve(NS, N, S).end * sizeof(float), "pmsqe_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* w = (float*)ws;
    SefdProfScope prof(SEFD_PROF_STFT, 0, 8.0 * N * L, st);
    pmsqe_stft_kernel<<<dim3(NS, 2), 256, 0, st>>>(est_wav, clean_wav, L, S, tables, w);
    SEFD_TRY(sefd_check_launch("pmsqe_stft"));
    pmsqe_bark_kernel<<<dim3(NS, 2), 256, 0, st>>>(S, tables, w);
    SEFD_TRY(sefd_check_launch("pmsqe_bark"));
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(pmsqe_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairSmem));
        cudaFuncSetAttribute(pmsqe_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairSmem));
        cudaFuncSetAttribute(pmsqe_mag_istft_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IstftSmem));
        attr = true;
    }
    pmsqe_pair_kernel<false><<<dim3(N, S * S), 64, sizeof(PairSmem), st>>>(N, S, tables, w, nullptr);
    SEFD_TRY(sefd_check_launch("pmsqe_pair"));
    pmsqe_pit_kernel<<<1, 256, 0, st>>>(N, S, w, loss);
    return sefd_check_launch("pmsqe_pit");
}

/* ws must be the workspace the matching forward filled.  d_est [N][L] receives gout * d loss / d est_wav. */
int sefd_pmsqe_backward(const float* gout, int N, int L, const float* tables, void* ws, size_t ws_bytes, float* d_est,
                        void* stream) {
    SEFD_REQUIRE(tables && ws && d_est && N > 0 && L > 0 && L % SEG == 0 && L / SEG <= MAXS, "pmsqe_backward: bad argument");
    const int S = L / SEG, NS = N * S;
    SEFD_REQUIRE(ws_bytes >= carve(NS, N, S).end * sizeof(float), "pmsqe_backward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* w = (float*)ws;
    SefdProfScope prof(SEFD_PROF_STFT, 0, 8.0 * N * L, st);
    pmsqe_pair_kernel<true><<<dim3(N, S), 64, sizeof(PairSmem), st>>>(N, S, tables, w, gout);
    SEFD_TRY(sefd_check_launch("pmsqe_pair_bwd"));
    pmsqe_mag_istft_bwd_kernel<<<NS, 256, sizeof(IstftSmem), st>>>(L, S, tables, w, d_est);
    return sefd_check_launch("pmsqe_mag_istft_bwd");
}

}  // extern "C"
