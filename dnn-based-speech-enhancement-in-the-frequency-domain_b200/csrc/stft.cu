// STFT, mask + ISTFT (+ clamp + loss dot products), the fused wave -> STFT -> mask -> ISTFT kernel and the adjoint, for
// win 400 / hop 100 / N 512 (every model) and win 800 / hop 200 / N 1024 (op level).
//
// Reference semantics reproduced (file:line relative to the reference checkout):
//  * ConvSTFT  (tools_for_model.py:54-61): zero-pad 300|300, frames of 400 x periodic Hann, zero-padded at
//    the END to 512, rFFT  ->  (sum x w cos, -sum x w sin).
//  * ConviSTFT (tools_for_model.py:90-112): synthesis with pinv of the truncated basis (NOT irfft). Closed form
//    used here (derivation in DESIGN.md, checked to 1e-14 against numpy pinv):
//        y[n] = Re sum_{k=0}^{256} S[k] e^{+2 pi i k n/512},  n < 400
//        frame[n] = w[n]/256 * ( y[n] - P_{n mod 2}/456 ),  P_q = sum_{n = q mod 2} y[n]
//    overlap-add, divide by coff = sum of 4 shifted w^2 + 1e-8, trim 300|300.
//  * DC mask bin is zero (models.py:255-256), mask modes C/E/R (models.py:258-276), clamp (models.py:282).
// Two real frames share one complex transform computed by one warp (fftw.cuh).
#include "fftw.cuh"
#include "stft.cuh"
#include "prof.cuh"

namespace {

constexpr int HOP = 100, NBIN = 257;      // the models' geometry (fft 512); the kernels themselves are templates over the fft length

// ------------------------------------------------------------------------------------------------
// mask application (models.py:253-276) and its Jacobian
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 apply_mask(int mode, float2 x, float2 m) {
    if (mode == SEFD_MASK_C) return make_float2(x.x * m.x - x.y * m.y, x.x * m.y + x.y * m.x);
    if (mode == SEFD_MASK_R) return make_float2(x.x * m.x, x.y * m.y);
    if (mode == SEFD_MASK_E) {
        const float smag = sqrtf(x.x * x.x + x.y * x.y + 1e-8f);
        const float sph = atan2f(x.y, x.x);
        const float mm = sqrtf(m.x * m.x + m.y * m.y);
        const float rp = m.x / (mm + 1e-8f), ip = m.y / (mm + 1e-8f);
        const float ph = sph + atan2f(ip, rp);
        const float em = tanhf(mm) * smag;
        float sn, cs;
        sincosf(ph, &sn, &cs);
        return make_float2(em * cs, em * sn);
    }
    if (mode == SEFD_MASK_DIRECT) return m;
    if (mode == SEFD_MASK_MAG) {          // m.x = real mask; m.y carries nothing
        const float mag = sqrtf(x.x * x.x + x.y * x.y);
        const float ph = atan2f(x.y, x.x);
        const float em = tanhf(m.x) * mag;
        float sn, cs;
        sincosf(ph, &sn, &cs);
        return make_float2(em * cs, em * sn);
    }
    return x;   // SEFD_MASK_NONE: plain ISTFT of `spec`
}
__device__ __forceinline__ float2 mask_bwd(int mode, float2 x, float2 m, float2 ds) {
    if (mode == SEFD_MASK_C) return make_float2(x.x * ds.x + x.y * ds.y, -x.y * ds.x + x.x * ds.y);
    if (mode == SEFD_MASK_R) return make_float2(x.x * ds.x, x.y * ds.y);
    if (mode == SEFD_MASK_E) {
        const float smag = sqrtf(x.x * x.x + x.y * x.y + 1e-8f);
        const float sph = atan2f(x.y, x.x);
        const float mm = sqrtf(m.x * m.x + m.y * m.y);
        const float den = mm + 1e-8f;
        const float rp = m.x / den, ip = m.y / den;
        const float ph = sph + atan2f(ip, rp);
        const float th = tanhf(mm);
        const float em = th * smag;
        float sn, cs;
        sincosf(ph, &sn, &cs);
        const float d_em = ds.x * cs + ds.y * sn;
        const float d_ph = em * (-ds.x * sn + ds.y * cs);
        float d_mm = d_em * smag * (1.f - th * th);
        const float r2 = rp * rp + ip * ip;
        float d_ip = 0.f, d_rp = 0.f;
        if (r2 > 0.f) { d_ip = rp / r2 * d_ph; d_rp = -ip / r2 * d_ph; }
        float dmx = d_rp / den, dmy = d_ip / den;
        d_mm += -(d_rp * m.x + d_ip * m.y) / (den * den);
        if (mm > 0.f) { dmx += d_mm * m.x / mm; dmy += d_mm * m.y / mm; }
        return make_float2(dmx, dmy);
    }
    if (mode == SEFD_MASK_DIRECT) return ds;
    if (mode == SEFD_MASK_MAG) {
        const float mag = sqrtf(x.x * x.x + x.y * x.y);
        const float ph = atan2f(x.y, x.x);
        const float th = tanhf(m.x);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        return make_float2((1.f - th * th) * mag * (ds.x * cs + ds.y * sn), 0.f);
    }
    return ds;   // NONE: gradient with respect to the spectrum itself
}

// ================================================================================================
// The kernels (fftw.cuh): one warp = one complex transform = two real frames; CTA = 8 warps = 16 frames per round.
// Both transform geometries of the reference's config (config.py:55-61): NFFT 512 (win 400 / hop 100, 257 bins) and
// NFFT 1024 (win 800 / hop 200, 513 bins); in both win = 4 hop and the zero padding is win - hop on either side.
// ================================================================================================
template <int NFFT>
struct Geo {
    static constexpr int WIN = NFFT / 32 * 25, HOP = WIN / 4, NBIN = NFFT / 2 + 1, PAD = WIN - HOP, R1 = NFFT / 32;
    static constexpr float INV_HALF = 2.0f / NFFT, INV_PARITY = 1.0f / (NFFT / 2 + WIN / 2);
};
constexpr int WF = 16;                  // frames per round of a CTA (8 warps x 2)
constexpr int OUT_PITCH = 18;           // float2 per bin row of the frame-contiguous staging tile: 16-byte aligned pairs,
                                        // 36-word stride = conflict-free for the 16-byte column writes and the row reads
constexpr int LU = 17;                  // tile elements whose global loads a thread keeps in flight together: all 17 bin rows of
                                        // its column for 257 bins (one DRAM latency per round), two batches for 513
constexpr int TILE_PITCH = 17;          // float2 per bin row of the synthesis tile: conflict-free 8-byte column reads

template <int NFFT>
__device__ __forceinline__ void init_window(float* win, float* coff) {
    using G = Geo<NFFT>;
    for (int n = threadIdx.x; n < G::WIN; n += blockDim.x) win[n] = 0.5f - 0.5f * cospif(2.0f * n / G::WIN);
    if (coff) {
        __syncthreads();
        for (int m = threadIdx.x; m < G::HOP; m += blockDim.x) {
            float a = 0.f;
            for (int r = 0; r < 4; ++r) a += win[m + G::HOP * r] * win[m + G::HOP * r];
            coff[m] = a + 1e-8f;
        }
    }
    __syncthreads();
}

// The two one-sided spectra of a frame pair from the transform Z of a + i b left in the warp's buffer:
// XA = (Z[k] + conj Z[N-k]) / 2, XB = (Z[k] - conj Z[N-k]) / (2i); lane holds bins lane, lane + 32, ... <= N / 2
template <int NFFT>
__device__ __forceinline__ void unpack_pair(const float2* s, float4 (&o)[NFFT / 64 + 1], int lane) {
#pragma unroll
    for (int j = 0; j < NFFT / 64 + 1; ++j) {
        const int k = lane + 32 * j;
        if (k <= NFFT / 2) {
            const float2 z = s[k], zc = s[(NFFT - k) & (NFFT - 1)];
            o[j] = make_float4(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y), 0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));
        }
    }
}
// ... and into columns (col, col + 1) of the frame-contiguous staging tile
template <int NFFT>
__device__ __forceinline__ void stage_pair(float2* stage, const float4 (&o)[NFFT / 64 + 1], int col, int lane) {
#pragma unroll
    for (int j = 0; j < NFFT / 64 + 1; ++j) {
        const int k = lane + 32 * j;
        if (k <= NFFT / 2) *reinterpret_cast<float4*>(stage + k * OUT_PITCH + col) = o[j];
    }
}
// windowed samples of frames (fa, fa + 1) of a zero-padded signal: v[n1] = w[n] * (x_a[n], x_b[n]), n = 32 n1 + lane
template <int NFFT>
__device__ __forceinline__ void load_pair(float2 (&v)[NFFT / 32], const float* __restrict__ w, const float* win, int fa,
                                          int T, int L, int lane) {
    using G = Geo<NFFT>;
    const int base = fa * G::HOP - G::PAD;
    if (base >= 0 && base + G::HOP + G::WIN <= L && fa + 1 < T) {      // both frames inside the signal (warp-uniform)
        const float* q = w + base + lane;
#pragma unroll
        for (int n1 = 0; n1 < G::R1; ++n1) {
            const int n = 32 * n1 + lane;
            v[n1] = make_float2(0.f, 0.f);
            if (32 * n1 < G::WIN && (32 * n1 + 31 < G::WIN || n < G::WIN)) {
                const float wn = win[n];
                v[n1] = make_float2(wn * __ldg(q + 32 * n1), wn * __ldg(q + 32 * n1 + G::HOP));
            }
        }
        return;
    }
#pragma unroll
    for (int n1 = 0; n1 < G::R1; ++n1) {
        const int n = 32 * n1 + lane;
        v[n1] = make_float2(0.f, 0.f);
        if (32 * n1 < G::WIN && n < G::WIN) {
            const int ia = base + n, ib = ia + G::HOP;
            const float wn = win[n];
            if (fa < T && ia >= 0 && ia < L) v[n1].x = wn * __ldg(w + ia);
            if (fa + 1 < T && ib >= 0 && ib < L) v[n1].y = wn * __ldg(w + ib);
        }
    }
}

template <int NFFT>
struct StftWSmem {
    static constexpr int FFT_LEN = 8 * fftw::Buf<NFFT>::LEN, OUT_LEN = Geo<NFFT>::NBIN * OUT_PITCH;
    float2 buf[FFT_LEN > OUT_LEN ? FFT_LEN : OUT_LEN];   // transform buffers of the 8 warps, then the staging tile
    float win[Geo<NFFT>::WIN];
};

// STFT forward: wav [B][L] -> spec [B][NBIN][T][2]; persistent over (utterance, 16-frame chunk) units
template <int NFFT>
__global__ void __launch_bounds__(256, NFFT == 512 ? 3 : 2) stft_fwd_w_kernel(const float* __restrict__ wav, float* __restrict__ spec,
                                                            int B, int L, int T) {
    using G = Geo<NFFT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftWSmem<NFFT>& S = *reinterpret_cast<StftWSmem<NFFT>*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    fftw::Twiddles<NFFT> tw;
    tw.init(lane);
    init_window<NFFT>(S.win, nullptr);
    float2* s = S.buf + warp * fftw::Buf<NFFT>::LEN;
    const int chunks = (T + WF - 1) / WF;
    // the next unit's samples travel while this one is staged and stored (N = 1024 has no registers to spare for that)
    constexpr bool PREFETCH = NFFT == 512;
    float2 v[G::R1];
    int u = blockIdx.x;
    if (PREFETCH && u < B * chunks) {
        const int b = u / chunks, t0 = (u - b * chunks) * WF;
        load_pair<NFFT>(v, wav + (long long)b * L, S.win, t0 + 2 * warp, T, L, lane);
    }
    for (; u < B * chunks; u += gridDim.x) {
        const int b = u / chunks, t0 = (u - b * chunks) * WF;
        if (!PREFETCH) load_pair<NFFT>(v, wav + (long long)b * L, S.win, t0 + 2 * warp, T, L, lane);
        fftw::fft_warp<NFFT, false>(v, s, tw, lane);
        float4 o[NFFT / 64 + 1];
        unpack_pair<NFFT>(s, o, lane);
        const int un = u + gridDim.x;
        if (PREFETCH && un < B * chunks) {
            const int bn = un / chunks, tn = (un - bn * chunks) * WF;
            load_pair<NFFT>(v, wav + (long long)bn * L, S.win, tn + 2 * warp, T, L, lane);
        }
        __syncthreads();                                   // every warp has left its transform buffer: the tile may alias it
        stage_pair<NFFT>(S.buf, o, 2 * warp, lane);
        __syncthreads();
        float2* out = reinterpret_cast<float2*>(spec) + (long long)b * G::NBIN * T;
        const int f = tid & 15, kq = tid >> 4;             // 16 lanes = the 128 contiguous bytes of one bin row
        if (t0 + f < T) {
            float2* op = out + (long long)kq * T + t0 + f;
            const float2* sp = S.buf + kq * OUT_PITCH + f;
            const long long step = 16LL * T;
#pragma unroll 4
            for (int k = kq; k < G::NBIN; k += 16, op += step, sp += 16 * OUT_PITCH) *op = *sp;
        }
        __syncthreads();
    }
}

// element i of the Hermitian-packed input A + iB of the inverse transform that synthesises two real frames at once;
// (sa, sb) = the one-sided spectra of the two frames at bin k = min(i, N - i)
template <int NFFT>
__device__ __forceinline__ float2 pack_hermitian(int i, float2 sa, float2 sb) {
    if (i == 0 || i == NFFT / 2) return make_float2(sa.x, sb.x);
    if (i < NFFT / 2) return make_float2(0.5f * (sa.x - sb.y), 0.5f * (sa.y + sb.x));
    return make_float2(0.5f * (sa.x + sb.y), 0.5f * (sb.x - sa.y));
}
// P_q / (N/2 + win/2) of the pair (ya, yb) in the warp's buffer, q = the lane's parity = the parity of its samples
template <int NFFT>
__device__ __forceinline__ float2 parity_means(const float2* s, int lane) {
    using G = Geo<NFFT>;
    float ea = 0.f, eb = 0.f;
#pragma unroll
    for (int i = 0; i < (G::WIN + 31) / 32; ++i) {
        const int n = lane + 32 * i;
        if (n < G::WIN) { const float2 e = s[n]; ea += e.x; eb += e.y; }
    }
#pragma unroll
    for (int o = 16; o >= 2; o >>= 1) {
        ea += __shfl_xor_sync(0xffffffffu, ea, o);
        eb += __shfl_xor_sync(0xffffffffu, eb, o);
    }
    return make_float2(ea * G::INV_PARITY, eb * G::INV_PARITY);
}

// overlap-add of the pair into the CTA's accumulator: frames whose index differs by a multiple of 4 never overlap
// (win = 4 hop), so four phases separated by CTA barriers give a fixed summation order without atomics
template <int NFFT>
__device__ __forceinline__ void overlap_add_pair(const float2* s, float2 par, const float* win, float* ola, int la, int lane) {
    using G = Geo<NFFT>;
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const int which = ((la & 3) == ph) ? 0 : (((la + 1) & 3) == ph) ? 1 : -1;
        if (which >= 0) {
            float* dst = ola + (la + which) * G::HOP;
#pragma unroll
            for (int i = 0; i < (G::WIN + 31) / 32; ++i) {
                const int n = lane + 32 * i;
                if (n < G::WIN) {
                    const float2 e = s[n];
                    dst[n] += win[n] * G::INV_HALF * (which ? e.y - par.y : e.x - par.x);
                }
            }
        }
        __syncthreads();
    }
}

constexpr int WIF = 32;                 // frames per CTA of the synthesis kernels (two rounds)
constexpr int WIHB = WIF - 3;           // hop blocks of finished output per CTA

template <int NFFT>
struct IstftWSmem {
    static constexpr int FFT_LEN = 8 * fftw::Buf<NFFT>::LEN, TILE_LEN = Geo<NFFT>::NBIN * TILE_PITCH;
    float2 buf[FFT_LEN > TILE_LEN ? FFT_LEN : TILE_LEN];  // masked-spectrum tile of a round, then the transform buffers
    float ola[(WIF + 3) * Geo<NFFT>::HOP];
    float win[Geo<NFFT>::WIN];
    float coff[Geo<NFFT>::HOP];
    float red[3][8];
};

template <int NFFT>
__device__ __forceinline__ float2 load_mask_g(const MaskIstftParams& p, int b, int k, int t) {
    const float* q = p.mask + b * p.mB + (long long)(k - 1) * p.mF + (long long)(t + p.m_tshift) * p.mT;
    if (p.mode == SEFD_MASK_MAG) return make_float2(__ldg(q), 0.f);
    return __ldg(reinterpret_cast<const float2*>(q));
}

// finalize hop blocks [3, 3 + WIHB) of a CTA's accumulator: / coff, clamp, store, loss dot products
template <int NFFT>
__device__ __forceinline__ void finish_chunk(const MaskIstftParams& p, const float* ola, const float* coff, float (*red)[8],
                                             int b, int f0) {
    using G = Geo<NFFT>;
    const int tid = threadIdx.x;
    float d12 = 0.f, d22 = 0.f, d11 = 0.f;
    for (int q = G::PAD + tid; q < G::PAD + WIHB * G::HOP; q += 256) {
        const int n = f0 * G::HOP + q - G::PAD;
        if (n < p.L) {
            const float raw = ola[q] / coff[q % G::HOP];
            const float v = fminf(fmaxf(raw, -1.f), 1.f);
            const long long o = (long long)b * p.L + n;
            p.out_wav[o] = v;
            if (p.raw_wav) p.raw_wav[o] = raw;
            if (p.target) {
                const float tg = __ldg(p.target + o);
                d12 += v * tg; d22 += tg * tg; d11 += v * v;
            }
        }
    }
    if (p.target) {
        d12 = warp_sum(d12); d22 = warp_sum(d22); d11 = warp_sum(d11);
        if ((tid & 31) == 0) { red[0][tid >> 5] = d12; red[1][tid >> 5] = d22; red[2][tid >> 5] = d11; }
        __syncthreads();
        if (tid < 3) {
            float a = 0.f;
            for (int i = 0; i < 8; ++i) a += red[tid][i];
            atomicAdd(p.dots + b * 8 + tid, (double)a);
        }
    }
}

// mask apply + ISTFT + clamp (+ loss dot products): spec [B][NBIN][T][2], mask as MaskIstftParams describes.
// MODE >= 0: the mask mode is a compile-time constant (the common modes: straight-line code), -1: p.mode at run time.
template <int NFFT, int MODE>
__global__ void __launch_bounds__(256, NFFT == 512 ? 3 : 2) mask_istft_fwd_w_kernel(const MaskIstftParams p) {
    using G = Geo<NFFT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IstftWSmem<NFFT>& S = *reinterpret_cast<IstftWSmem<NFFT>*>(smem_raw);
    const int b = blockIdx.y, c = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int f0 = WIHB * c, T = p.T;
    const int mode = MODE >= 0 ? MODE : p.mode;
    fftw::Twiddles<NFFT> tw;
    tw.init(lane);
    for (int i = tid; i < (WIF + 3) * G::HOP; i += 256) S.ola[i] = 0.f;
    init_window<NFFT>(S.win, S.coff);
    const float2* X = reinterpret_cast<const float2*>(p.spec) + (long long)b * G::NBIN * T;
    float2* s = S.buf + warp * fftw::Buf<NFFT>::LEN;
    // tile element (bin kq + 16 i, frame f): 16 lanes = the 128 contiguous bytes of one bin row
    const int f = tid & 15, kq = tid >> 4;
    constexpr int NI = (G::NBIN + 15) / 16;
    const long long xstep = 16LL * T, mstep = 16 * p.mF;

    for (int r = 0; r < WIF / WF; ++r) {
        const int la = WF * r + f, t = f0 + la;
        const bool tv = t < T;
        // ownership of the spectrum outputs: frames [f0 + 3, f0 + 3 + WIHB) plus frames 0..2 in the first CTA
        const bool own = p.out_real && tv && (la >= 3 ? la < 3 + WIHB : c == 0);
        const float2* xp = X + (long long)kq * T + t;
        const float* mp = p.mask + b * p.mB + (long long)(kq - 1) * p.mF + (long long)(t + p.m_tshift) * p.mT;
        float2* tp = S.buf + kq * TILE_PITCH + f;
        // the loads of LU elements are issued together (one element per iteration left the CTA waiting a DRAM latency
        // NI times a round)
        for (int i0 = 0; i0 < NI; i0 += LU, xp += LU * xstep, mp += LU * mstep, tp += LU * 16 * TILE_PITCH) {
            float2 x[LU], m[LU];
#pragma unroll
            for (int j = 0; j < LU; ++j) {
                const int k = kq + 16 * (i0 + j);
                x[j] = m[j] = make_float2(0.f, 0.f);
                if (tv && k < G::NBIN) {
                    x[j] = __ldg(xp + j * xstep);
                    if (mode != SEFD_MASK_NONE && k >= 1) {
                        if (mode == SEFD_MASK_MAG) m[j].x = __ldg(mp + j * mstep);
                        else m[j] = __ldg(reinterpret_cast<const float2*>(mp + j * mstep));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < LU; ++j) {
                const int k = kq + 16 * (i0 + j);
                if (k < G::NBIN) {
                    float2 sv = make_float2(0.f, 0.f);
                    if (tv && !(mode != SEFD_MASK_NONE && k == 0)) sv = apply_mask(mode, x[j], m[j]);
                    if (own) {
                        const long long o = ((long long)b * G::NBIN + k) * T + t;
                        if (mode == SEFD_MASK_MAG) {          // est_mags = tanh(mask) * |X| (models.py:521-522)
                            p.out_real[o] = tanhf(m[j].x) * sqrtf(x[j].x * x[j].x + x[j].y * x[j].y);
                        } else {
                            p.out_real[o] = sv.x;
                            p.out_imag[o] = sv.y;
                        }
                    }
                    tp[j * 16 * TILE_PITCH] = sv;
                }
            }
        }
        __syncthreads();
        // the tile is read into registers by every warp BEFORE any warp's transform reuses the memory
        float2 v[G::R1];
#pragma unroll
        for (int n1 = 0; n1 < G::R1; ++n1) {
            const int i = 32 * n1 + lane, k = i <= NFFT / 2 ? i : NFFT - i;
            v[n1] = pack_hermitian<NFFT>(i, S.buf[k * TILE_PITCH + 2 * warp], S.buf[k * TILE_PITCH + 2 * warp + 1]);
        }
        __syncthreads();
        fftw::fft_warp<NFFT, true>(v, s, tw, lane);
        const float2 par = parity_means<NFFT>(s, lane);
        overlap_add_pair<NFFT>(s, par, S.win, S.ola, WF * r + 2 * warp, lane);
    }
    finish_chunk<NFFT>(p, S.ola, S.coff, S.red, b, f0);
}

// ------------------------------------------------------------------------------------------------
// BASELINE configs[4], fully fused form: wave -> STFT -> mask -> ISTFT -> wave in ONE kernel; the spectrum never leaves
// the SM (8 F + 8 hop algorithmic bytes per frame).  Complex mask [B][NBIN - 1][T][2] (DC bin zero).
// ------------------------------------------------------------------------------------------------
template <int NFFT>
struct FusedWSmem {
    float2 fft[8 * fftw::Buf<NFFT>::LEN];
    float2 tile[Geo<NFFT>::NBIN * TILE_PITCH];            // mask of a round
    float ola[(WIF + 3) * Geo<NFFT>::HOP];
    float win[Geo<NFFT>::WIN];
    float coff[Geo<NFFT>::HOP];
    float red[3][8];
};

template <int NFFT, int MODE>
__global__ void __launch_bounds__(256, 2) stft_mask_istft_w_kernel(const float* __restrict__ wav, const MaskIstftParams p) {
    using G = Geo<NFFT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedWSmem<NFFT>& S = *reinterpret_cast<FusedWSmem<NFFT>*>(smem_raw);
    const int b = blockIdx.y, c = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int f0 = WIHB * c, T = p.T, L = p.L;
    const int mode = MODE >= 0 ? MODE : p.mode;
    fftw::Twiddles<NFFT> tw;
    tw.init(lane);
    for (int i = tid; i < (WIF + 3) * G::HOP; i += 256) S.ola[i] = 0.f;
    init_window<NFFT>(S.win, S.coff);
    float2* s = S.fft + warp * fftw::Buf<NFFT>::LEN;
    const float* w = wav + (long long)b * L;
    const int f = tid & 15, kq = tid >> 4;
    constexpr int NI = (G::NBIN + 15) / 16;
    const long long mstep = 16 * p.mF;

    for (int r = 0; r < WIF / WF; ++r) {
        const int la = WF * r + 2 * warp, fa = f0 + la;
        float2 v[G::R1];
        load_pair<NFFT>(v, w, S.win, fa, T, L, lane);
        {
            const int t = f0 + WF * r + f;
            const float* mp = p.mask + b * p.mB + (long long)(kq - 1) * p.mF + (long long)(t + p.m_tshift) * p.mT;
            float2* tp = S.tile + kq * TILE_PITCH + f;
            for (int i0 = 0; i0 < NI; i0 += LU, mp += LU * mstep, tp += LU * 16 * TILE_PITCH) {
                float2 m[LU];
#pragma unroll
                for (int j = 0; j < LU; ++j) {
                    const int k = kq + 16 * (i0 + j);
                    m[j] = make_float2(0.f, 0.f);
                    if (t < T && k >= 1 && k < G::NBIN) m[j] = __ldg(reinterpret_cast<const float2*>(mp + j * mstep));
                }
#pragma unroll
                for (int j = 0; j < LU; ++j)
                    if (kq + 16 * (i0 + j) < G::NBIN) tp[j * 16 * TILE_PITCH] = m[j];
            }
        }
        fftw::fft_warp<NFFT, false>(v, s, tw, lane);
        __syncthreads();                                   // the mask tile is complete
        // masked Hermitian-packed spectrum straight into the registers of the inverse transform
#pragma unroll
        for (int n1 = 0; n1 < G::R1; ++n1) {
            const int i = 32 * n1 + lane, k = i <= NFFT / 2 ? i : NFFT - i;
            const float2 z = s[k], zc = s[(NFFT - k) & (NFFT - 1)];
            const float2 xa = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
            const float2 xb = make_float2(0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));
            const float2 sa = apply_mask(mode, xa, S.tile[k * TILE_PITCH + 2 * warp]);
            const float2 sb = apply_mask(mode, xb, S.tile[k * TILE_PITCH + 2 * warp + 1]);
            v[n1] = pack_hermitian<NFFT>(i, sa, sb);
        }
        fftw::fft_warp<NFFT, true>(v, s, tw, lane);
        overlap_add_pair<NFFT>(s, parity_means<NFFT>(s, lane), S.win, S.ola, la, lane);
    }
    finish_chunk<NFFT>(p, S.ola, S.coff, S.red, b, f0);
}

// ------------------------------------------------------------------------------------------------
// adjoint: d wav -> d mask   (ISTFT^T, then the mask-apply Jacobian; no gradient to the noisy spectrum).
// ISTFT^T of a frame = forward transform of w[n] / (N/2) * (g[n] - parity means of it), g = d wav / coff gated by the
// clamp: the same analysis structure as the STFT kernel (one warp per frame pair, staging tile aliasing the buffers).
// ------------------------------------------------------------------------------------------------
template <int NFFT>
struct IstftBwdWSmem {
    static constexpr int FFT_LEN = 8 * fftw::Buf<NFFT>::LEN, OUT_LEN = Geo<NFFT>::NBIN * OUT_PITCH;
    float2 buf[FFT_LEN > OUT_LEN ? FFT_LEN : OUT_LEN];
    float win[Geo<NFFT>::WIN];
    float coff[Geo<NFFT>::HOP];
};

template <int NFFT>
__global__ void __launch_bounds__(256, 2) mask_istft_bwd_w_kernel(const MaskIstftBwdParams p) {
    using G = Geo<NFFT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IstftBwdWSmem<NFFT>& S = *reinterpret_cast<IstftBwdWSmem<NFFT>*>(smem_raw);
    const int b = blockIdx.y, t0 = blockIdx.x * WF;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T, L = p.L;
    fftw::Twiddles<NFFT> tw;
    tw.init(lane);
    init_window<NFFT>(S.win, S.coff);
    float2* s = S.buf + warp * fftw::Buf<NFFT>::LEN;
    const int fa = t0 + 2 * warp;
    // g[n] of frame f: d wav at sample f hop + n - pad (zero outside the signal or where the clamp was active) / coff
    auto grad_at = [&](int P) {
        const int n = P - G::PAD;
        float v = 0.f;
        if (n >= 0 && n < L) {
            const long long o = (long long)b * L + n;
            v = p.dwav ? __ldg(p.dwav + o) : 0.f;
            if (p.raw_wav) {
                const float raw = __ldg(p.raw_wav + o);
                if (!(raw >= -1.f && raw <= 1.f)) v = 0.f;
            }
            v /= S.coff[P % G::HOP];
        }
        return v;
    };
    float2 v[G::R1];
    float ea = 0.f, eb = 0.f;
#pragma unroll
    for (int n1 = 0; n1 < G::R1; ++n1) {
        const int n = 32 * n1 + lane;
        v[n1] = make_float2(0.f, 0.f);
        if (32 * n1 < G::WIN && n < G::WIN) {
            const float wn = S.win[n] * G::INV_HALF;
            v[n1] = make_float2(wn * grad_at(fa * G::HOP + n), wn * grad_at((fa + 1) * G::HOP + n));
            ea += v[n1].x; eb += v[n1].y;
        }
    }
#pragma unroll
    for (int o = 16; o >= 2; o >>= 1) {                     // sums over the lanes (= samples) of this lane's parity
        ea += __shfl_xor_sync(0xffffffffu, ea, o);
        eb += __shfl_xor_sync(0xffffffffu, eb, o);
    }
#pragma unroll
    for (int n1 = 0; n1 < G::R1; ++n1) {
        const int n = 32 * n1 + lane;
        if (32 * n1 < G::WIN && n < G::WIN) { v[n1].x -= ea * G::INV_PARITY; v[n1].y -= eb * G::INV_PARITY; }
    }
    fftw::fft_warp<NFFT, false>(v, s, tw, lane);
    float4 o[NFFT / 64 + 1];
    unpack_pair<NFFT>(s, o, lane);
    __syncthreads();
    stage_pair<NFFT>(S.buf, o, 2 * warp, lane);
    __syncthreads();

    const float2* X = reinterpret_cast<const float2*>(p.spec) + (long long)b * G::NBIN * T;
    const int k_lo = (p.mode == SEFD_MASK_NONE) ? 0 : 1;
    for (int e = tid; e < G::NBIN * WF; e += 256) {
        const int k = e >> 4, f = e & 15, t = t0 + f;
        if (t >= T || k < k_lo) continue;
        float2 ds = S.buf[k * OUT_PITCH + f];
        float dmag = 0.f;                 // SEFD_MASK_MAG: gradient arriving at est_mags = tanh(mask) |X|
        if (p.dreal) {
            const long long o2 = ((long long)b * G::NBIN + k) * T + t;
            if (p.mode == SEFD_MASK_MAG) {
                dmag = __ldg(p.dreal + o2);
            } else {
                ds.x += __ldg(p.dreal + o2);
                ds.y += __ldg(p.dimag + o2);
            }
        }
        float2 x = make_float2(0.f, 0.f), m = x;
        if (p.mode != SEFD_MASK_NONE) {
            x = __ldg(X + (long long)k * T + t);
            const float* mq = p.mask + b * p.mB + (long long)(k - 1) * p.mF + (long long)(t + p.m_tshift) * p.mT;
            if (p.mode == SEFD_MASK_E) m = __ldg(reinterpret_cast<const float2*>(mq));
            if (p.mode == SEFD_MASK_MAG) m.x = __ldg(mq);
        }
        float2 dm = mask_bwd(p.mode, x, m, ds);
        if (p.mode == SEFD_MASK_MAG && p.dreal) {
            const float th = tanhf(m.x);
            dm.x += dmag * (1.f - th * th) * sqrtf(x.x * x.x + x.y * x.y);
        }
        float* dq = p.dmask + b * p.mB + (long long)(k - k_lo) * p.mF + (long long)(t + p.m_tshift) * p.mT;
        if (p.mode == SEFD_MASK_MAG) *dq = dm.x;
        else *reinterpret_cast<float2*>(dq) = dm;
    }
    // frames in front of the shift (the decoder's dropped look-ahead frame) get zero gradient
    if (blockIdx.x == 0 && p.m_tshift > 0) {
        for (int e = tid; e < (G::NBIN - k_lo) * p.m_tshift; e += 256) {
            const int k = e / p.m_tshift, t = e % p.m_tshift;
            float* dq = p.dmask + b * p.mB + (long long)k * p.mF + (long long)t * p.mT;
            if (p.mode == SEFD_MASK_MAG) *dq = 0.f;
            else *reinterpret_cast<float2*>(dq) = make_float2(0.f, 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// losses (tools_for_loss.py:29-94; selection models.py:315-323).  dots[b][8] (double):
//   0 <E,G>  1 <G,G>  2 <E,E>  3 sum res^2  4 sum res*G      res = E - c_b G
// coef[b][2]: d loss / d E[b,n] = coef[b][0] * E[b,n] + coef[b][1] * G[b,n]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_dots_kernel(const float* __restrict__ E, const float* __restrict__ G,
                                                        int L, double* dots) {
    __shared__ float red[3][8];
    const int b = blockIdx.y;
    const float* e = E + (long long)b * L;
    const float* g = G + (long long)b * L;
    float d12 = 0.f, d22 = 0.f, d11 = 0.f;
    for (int n = blockIdx.x * 256 + threadIdx.x; n < L; n += gridDim.x * 256) {
        const float a = e[n], c = g[n];
        d12 += a * c; d22 += c * c; d11 += a * a;
    }
    d12 = warp_sum(d12); d22 = warp_sum(d22); d11 = warp_sum(d11);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d12; red[1][threadIdx.x >> 5] = d22; red[2][threadIdx.x >> 5] = d11; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += red[threadIdx.x][i];
        atomicAdd(dots + b * 8 + threadIdx.x, (double)a);
    }
}

__device__ __forceinline__ double loss_scale(int kind, const double* d) {
    const double eps = 1e-8;
    if (kind == SEFD_LOSS_SISNR) return d[0] / (d[1] + eps);
    if (kind == SEFD_LOSS_SISDR) return d[0] / d[1] + eps;
    return 1.0;
}

__global__ void __launch_bounds__(256) loss_resid_kernel(const float* __restrict__ E, const float* __restrict__ G,
                                                         int L, int kind, double* dots) {
    __shared__ float red[2][8];
    const int b = blockIdx.y;
    const float* e = E + (long long)b * L;
    const float* g = G + (long long)b * L;
    const float c = (float)loss_scale(kind, dots + b * 8);
    float rr = 0.f, rg = 0.f;
    for (int n = blockIdx.x * 256 + threadIdx.x; n < L; n += gridDim.x * 256) {
        const float gg = g[n];
        const float r = e[n] - c * gg;
        rr += r * r; rg += r * gg;
    }
    rr = warp_sum(rr); rg = warp_sum(rg);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = rr; red[1][threadIdx.x >> 5] = rg; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += red[threadIdx.x][i];
        atomicAdd(dots + b * 8 + 3 + threadIdx.x, (double)a);
    }
}

__global__ void loss_finalize_kernel(const double* dots, int B, int L, int kind, float* loss, float* coef) {
    // single CTA; B is small (<= a few hundred)
    __shared__ double acc[256];
    const double eps = 1e-8, kappa = 10.0 / log(10.0);
    double local = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const double* d = dots + b * 8;
        const double s12 = d[0], s22 = d[1], rr = d[3], rg = d[4];
        double term = 0.0, cA = 0.0, cB = 0.0;
        if (kind == SEFD_LOSS_MSE) {
            term = rr / ((double)B * L);
            cA = 2.0 / ((double)B * L);
            cB = -cA;
        } else if (kind == SEFD_LOSS_SDR) {
            term = -kappa * log(s22 * s22 / (rr * rr + eps)) / B;
            cA = 4.0 * kappa * rr / (B * (rr * rr + eps));
            cB = -cA;
        } else if (kind == SEFD_LOSS_SISNR) {
            const double al = s12 / (s22 + eps);
            const double tt = al * al * s22;
            const double r = tt / (rr + eps) + eps;
            term = -kappa * log(r) / B;
            const double pre = kappa / (B * r);
            cA = pre * 2.0 * tt / ((rr + eps) * (rr + eps));
            cB = -pre * (2.0 * al * s22 / ((s22 + eps) * (rr + eps)) +
                         2.0 * tt * (al + rg / (s22 + eps)) / ((rr + eps) * (rr + eps)));
        } else {   // SI-SDR: needs the batch mean of the ratios first
            const double a = s12 / s22 + eps;
            term = (a * a * s22 / rr + eps) / B;
        }
        local += term;
        if (kind != SEFD_LOSS_SISDR) { coef[2 * b] = (float)cA; coef[2 * b + 1] = (float)cB; }
    }
    acc[threadIdx.x] = local;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) acc[threadIdx.x] += acc[threadIdx.x + o];
        __syncthreads();
    }
    const double total = acc[0];
    if (kind == SEFD_LOSS_SISDR) {
        const double R = total;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const double* d = dots + b * 8;
            const double s12 = d[0], s22 = d[1], nn = d[3], ng = d[4];
            const double a = s12 / s22 + eps, pp = a * a * s22;
            const double pre = kappa / ((R + eps) * B);
            coef[2 * b] = (float)(pre * 2.0 * pp / (nn * nn));
            coef[2 * b + 1] = (float)(-pre * (2.0 * a / nn + 2.0 * pp * a / (nn * nn) + 2.0 * pp * ng / (nn * nn * s22)));
        }
        if (threadIdx.x == 0) loss[0] = (float)(-kappa * log(R + eps));
    } else if (threadIdx.x == 0) {
        loss[0] = (float)total;
    }
}

__global__ void loss_bwd_kernel(const float* __restrict__ E, const float* __restrict__ G, const float* __restrict__ coef,
                                const float* __restrict__ gout, float* __restrict__ dE, int B, int L) {
    const float go = gout ? gout[0] : 1.f;
    const long long total = (long long)B * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / L);
        dE[i] = go * (coef[2 * b] * E[i] + coef[2 * b + 1] * G[i]);
    }
}

__global__ void spec_mag_kernel(const float2* __restrict__ spec, float* __restrict__ mag, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 x = __ldg(spec + i);
        mag[i] = sqrtf(x.x * x.x + x.y * x.y);
    }
}

}  // namespace

int sefd_spec_mag_launch(const float* spec, float* mag, long long n, cudaStream_t st) {
    long long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    SefdProfScope prof(SEFD_PROF_STFT, 0, 12.0 * n, st);
    spec_mag_kernel<<<(int)g, 256, 0, st>>>(reinterpret_cast<const float2*>(spec), mag, n);
    return sefd_check_launch("spec_mag");
}

template <int NFFT>
static int stft_launch_w(const float* wav, float* spec, int B, int L, cudaStream_t st) {
    using G = Geo<NFFT>;
    SEFD_REQUIRE(L > 0 && L % G::HOP == 0, "stft: L=%d must be a positive multiple of the hop %d", L, G::HOP);
    const int T = L / G::HOP + 3;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(stft_fwd_w_kernel<NFFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftWSmem<NFFT>));
        attr = true;
    }
    const long long units = (long long)B * ((T + WF - 1) / WF);
    const int grid = (int)(units < 148 * 8 ? units : 148 * 8);
    SefdProfScope prof(SEFD_PROF_STFT, 0, 4.0 * B * L + 8.0 * B * G::NBIN * T, st);
    stft_fwd_w_kernel<NFFT><<<grid, 256, sizeof(StftWSmem<NFFT>), st>>>(wav, spec, B, L, T);
    return sefd_check_launch("stft_fwd");
}

template <int NFFT, int MODE>
static int mask_istft_launch_wm(const MaskIstftParams& p, const float* fused_wav, cudaStream_t st) {
    using G = Geo<NFFT>;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(mask_istft_fwd_w_kernel<NFFT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IstftWSmem<NFFT>));
        cudaFuncSetAttribute(stft_mask_istft_w_kernel<NFFT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedWSmem<NFFT>));
        attr = true;
    }
    dim3 grid((p.L / G::HOP + WIHB - 1) / WIHB, p.B);
    if (fused_wav) {
        SefdProfScope prof(SEFD_PROF_STFT, 0, 8.0 * p.B * (G::NBIN - 1) * p.T + 8.0 * p.B * p.L, st);
        stft_mask_istft_w_kernel<NFFT, MODE><<<grid, 256, sizeof(FusedWSmem<NFFT>), st>>>(fused_wav, p);
        return sefd_check_launch("stft_mask_istft");
    }
    SefdProfScope prof(SEFD_PROF_STFT, 0, 8.0 * p.B * G::NBIN * p.T * (p.mode != SEFD_MASK_NONE ? 2 : 1) +
                       4.0 * p.B * p.L * (p.target ? 3 : 2) + (p.out_real ? 8.0 * p.B * G::NBIN * p.T : 0.0), st);
    mask_istft_fwd_w_kernel<NFFT, MODE><<<grid, 256, sizeof(IstftWSmem<NFFT>), st>>>(p);
    return sefd_check_launch("mask_istft_fwd");
}

template <int NFFT>
static int mask_istft_launch_w(const MaskIstftParams& p, const float* fused_wav, cudaStream_t st) {
    using G = Geo<NFFT>;
    SEFD_REQUIRE(p.L > 0 && p.L % G::HOP == 0 && p.T == p.L / G::HOP + 3, "istft: L=%d / T=%d inconsistent (hop %d)", p.L, p.T, G::HOP);
    if (p.target) cudaMemsetAsync(p.dots, 0, sizeof(double) * 8 * p.B, st);
    if (p.mode == SEFD_MASK_C) return mask_istft_launch_wm<NFFT, SEFD_MASK_C>(p, fused_wav, st);
    if (p.mode == SEFD_MASK_NONE) return mask_istft_launch_wm<NFFT, SEFD_MASK_NONE>(p, fused_wav, st);
    return mask_istft_launch_wm<NFFT, -1>(p, fused_wav, st);
}

int sefd_stft_launch(const float* wav, float* spec, int B, int L, int T, cudaStream_t st) {
    SEFD_REQUIRE(L % HOP == 0 && T == L / HOP + 3, "stft: L=%d must be a multiple of %d and T=%d == L/hop+3", L, HOP, T);
    return stft_launch_w<512>(wav, spec, B, L, st);
}

int sefd_stft_launch_n(const float* wav, float* spec, int B, int L, int nfft, cudaStream_t st) {
    SEFD_REQUIRE(nfft == 512 || nfft == 1024, "stft: fft length %d is not built (512: win 400 / hop 100, 1024: win 800 / hop 200)", nfft);
    return nfft == 512 ? stft_launch_w<512>(wav, spec, B, L, st) : stft_launch_w<1024>(wav, spec, B, L, st);
}

int sefd_mask_istft_launch_n(const MaskIstftParams& p, const float* fused_wav, int nfft, cudaStream_t st) {
    SEFD_REQUIRE(nfft == 512 || nfft == 1024, "istft: fft length %d is not built (512: win 400 / hop 100, 1024: win 800 / hop 200)", nfft);
    return nfft == 512 ? mask_istft_launch_w<512>(p, fused_wav, st) : mask_istft_launch_w<1024>(p, fused_wav, st);
}

int sefd_mask_istft_launch(const MaskIstftParams& p, cudaStream_t st) { return mask_istft_launch_w<512>(p, nullptr, st); }

int sefd_mask_istft_bwd_launch(const MaskIstftBwdParams& p, cudaStream_t st) {
    SEFD_REQUIRE(p.L > 0 && p.L % HOP == 0 && p.T == p.L / HOP + 3, "istft backward: L=%d / T=%d inconsistent", p.L, p.T);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(mask_istft_bwd_w_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IstftBwdWSmem<512>));
        attr = true;
    }
    dim3 grid((p.T + WF - 1) / WF, p.B);
    SefdProfScope prof(SEFD_PROF_STFT, 0, 8.0 * p.B * NBIN * p.T * 2 + 8.0 * p.B * p.L, st);
    mask_istft_bwd_w_kernel<512><<<grid, 256, sizeof(IstftBwdWSmem<512>), st>>>(p);
    return sefd_check_launch("mask_istft_bwd");
}

int sefd_loss_fwd_launch(const float* est, const float* tgt, int B, int L, int kind, double* dots, int dots_ready,
                         float* loss, float* coef, cudaStream_t st) {
    SEFD_REQUIRE(kind >= 0 && kind <= 3, "loss: unknown kind %d", kind);
    dim3 grid(8, B);
    SefdProfScope prof(SEFD_PROF_STFT, 0, 8.0 * B * L * (dots_ready ? 1 : 2), st);
    if (!dots_ready) {
        cudaMemsetAsync(dots, 0, sizeof(double) * 8 * B, st);
        loss_dots_kernel<<<grid, 256, 0, st>>>(est, tgt, L, dots);
        SEFD_TRY(sefd_check_launch("loss_dots"));
    }
    cudaMemset2DAsync(dots + 3, 8 * sizeof(double), 0, 2 * sizeof(double), B, st);
    loss_resid_kernel<<<grid, 256, 0, st>>>(est, tgt, L, kind, dots);
    SEFD_TRY(sefd_check_launch("loss_resid"));
    loss_finalize_kernel<<<1, 256, 0, st>>>(dots, B, L, kind, loss, coef);
    return sefd_check_launch("loss_finalize");
}

int sefd_loss_bwd_launch(const float* est, const float* tgt, const float* coef, const float* gout, float* dest,
                         int B, int L, cudaStream_t st) {
    long long n = (long long)B * L;
    int g = (int)((n + 255) / 256);
    if (g > 148 * 8) g = 148 * 8;
    SefdProfScope prof(SEFD_PROF_STFT, 0, 12.0 * B * L, st);
    loss_bwd_kernel<<<g, 256, 0, st>>>(est, tgt, coef, gout, dest, B, L);
    return sefd_check_launch("loss_bwd");
}
