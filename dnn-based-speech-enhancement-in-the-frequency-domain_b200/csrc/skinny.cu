// Kernels for the two-channel ends of the network (encoder 0 reads the 2-channel spectrum, decoder 5 writes the
// 2-channel mask).  These contractions have K = 2 or N = 2 per tap: no tensor-core shape fits and the work is
// HBM-bound (intensity ~9 FLOP/B), so they are CUDA-core kernels organised around ONE pass
// over the wide (32/64-channel) tensor with the 2-channel tensor read at the tap offsets.  Same TapGemmParams /
// WgradParams contract as the generic engines.
#include <string.h>

#include "common.cuh"
#include "prof.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// T1: K = 2 per tap -> N in {32, 64}   (encoder-0 forward, decoder-5 data gradient)
// one thread = one output position; the <= 20 input scalars come straight from global (coalesced float2).
// ---------------------------------------------------------------------------------------------------
__constant__ float c_smallkW[10 * 2 * 64];     // [tap][k][n] of the current launch (stream-ordered upload)

template <int N>
__global__ void __launch_bounds__(128) smallk_conv_kernel(const TapGemmParams p) {
    __shared__ float stg[128 * (N + 1)];
    const int tid = threadIdx.x;
    const int ttiles = (p.Tout + 127) / 128;
    const int t0 = (blockIdx.x % ttiles) * 128;
    const int row = blockIdx.x / ttiles;
    const int b = row / p.J, j = row % p.J;
    float acc[N];
#pragma unroll
    for (int n = 0; n < N; ++n) acc[n] = p.bias ? __ldg(p.bias + n) : 0.f;
    const int t = t0 + tid;
    // the 10 taps are unrolled: the weights are constant-bank FFMA operands (no shared-memory traffic); the <= 20
    // input scalars come straight from global (coalesced float2)
    float2 x[10];
#pragma unroll
    for (int tap = 0; tap < 10; ++tap) {
        const int fi = j * p.fi_mul + p.df[tap];
        const int tin = t + p.dt[tap];
        x[tap] = make_float2(0.f, 0.f);
        if (fi >= 0 && fi < p.Fin && tin >= 0 && tin < p.Tin)
            x[tap] = __ldg(reinterpret_cast<const float2*>(p.a[0].p + b * p.a[0].sB + fi * p.a[0].sF + (long long)tin * p.a[0].sT));
    }
#pragma unroll
    for (int tap = 0; tap < 10; ++tap)
#pragma unroll
        for (int n = 0; n < N; ++n)
            acc[n] = fmaf(x[tap].x, c_smallkW[(tap * 2 + 0) * N + n], fmaf(x[tap].y, c_smallkW[(tap * 2 + 1) * N + n], acc[n]));
#pragma unroll
    for (int n = 0; n < N; ++n) stg[tid * (N + 1) + n] = acc[n];
    __syncthreads();
    // coalesced stores: N/4 float4 per row
    const int fo = j * p.fo_mul + p.fo_off;
    const int N0 = p.o[0].N;
    constexpr int Q = N / 4;
    for (int idx = tid; idx < 128 * Q; idx += 128) {
        const int r = idx / Q, c4 = (idx % Q) * 4;
        const int tt = t0 + r;
        if (tt >= p.Tout) continue;
        const int d = c4 < N0 ? 0 : 1;
        const int nn = d ? c4 - N0 : c4;
        float* dst = p.o[d].p + b * p.o[d].sB + fo * p.o[d].sF + (long long)tt * p.o[d].sT + nn;
        const float* sp = stg + r * (N + 1) + c4;
        float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
        if (p.accum[d]) {
            const float4 o = *reinterpret_cast<const float4*>(dst);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *reinterpret_cast<float4*>(dst) = v;
    }
    if (p.stats) {
        // 128 threads: column c = tid % N, row part = tid / N.  The partial sums are combined in a FIXED order
        // (no floating-point atomics): the BatchNorm statistics - and with them every PReLU branch decision
        // downstream - must not depend on the scheduling of this CTA's warps.
        constexpr int PARTS = 128 / N;
        const int c = tid % N, part = tid / N;
        float s1 = 0.f, s2 = 0.f;
        for (int r = part; r < 128; r += PARTS) {
            if (t0 + r < p.Tout) {
                const float x = stg[r * (N + 1) + c];
                s1 += x;
                s2 += x * x;
            }
        }
        __syncthreads();                       // all reads of stg are done: reuse it for the partials
        stg[part * 2 * N + c] = s1;
        stg[part * 2 * N + N + c] = s2;
        __syncthreads();
        if (tid < 2 * N) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < PARTS; ++q) s += stg[q * 2 * N + tid];
            atomicAdd(p.stats + tid, (double)s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// T3: N = 2   (decoder-5 forward): one thread = one output position, serial dot products over K channels
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) smalln_conv_kernel(const TapGemmParams p) {
    extern __shared__ __align__(16) float Wsm[];      // [ntaps][K][2]
    const int tid = threadIdx.x;
    const int C0 = p.a[0].C, C1 = p.a[1].C, K = C0 + C1;
    const int ttiles = (p.Tout + 127) / 128;
    const int t0 = (blockIdx.x % ttiles) * 128;
    const int row = blockIdx.x / ttiles;
    const int b = row / p.J, j = row % p.J;
    for (int i = tid; i < p.ntaps * K * 2; i += 128) {
        const int tap = i / (K * 2), r = i % (K * 2);
        Wsm[i] = p.W[(long long)p.wslab[tap] * K * 2 + r];
    }
    __syncthreads();
    const int t = t0 + tid;
    if (t >= p.Tout) return;
    float a0 = p.bias ? __ldg(p.bias) : 0.f, a1 = p.bias ? __ldg(p.bias + 1) : 0.f;
    for (int tap = 0; tap < p.ntaps; ++tap) {
        const int fi = j * p.fi_mul + p.df[tap];
        if (fi < 0 || fi >= p.Fin) continue;
        const int tin = t + p.dt[tap];
        if (tin < 0 || tin >= p.Tin) continue;
        const float2* w = reinterpret_cast<const float2*>(Wsm + tap * K * 2);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int C = s ? C1 : C0;
            if (!C) continue;
            const float4* x = reinterpret_cast<const float4*>(p.a[s].p + b * p.a[s].sB + fi * p.a[s].sF + (long long)tin * p.a[s].sT);
            const float2* ws = w + (s ? C0 : 0);
#pragma unroll 4
            for (int k4 = 0; k4 < C / 4; ++k4) {
                const float4 v = __ldg(x + k4);
                const float2 w0 = ws[4 * k4], w1 = ws[4 * k4 + 1], w2 = ws[4 * k4 + 2], w3 = ws[4 * k4 + 3];
                a0 = fmaf(v.x, w0.x, fmaf(v.y, w1.x, fmaf(v.z, w2.x, fmaf(v.w, w3.x, a0))));
                a1 = fmaf(v.x, w0.y, fmaf(v.y, w1.y, fmaf(v.z, w2.y, fmaf(v.w, w3.y, a1))));
            }
        }
    }
    const int fo = j * p.fo_mul + p.fo_off;
    float* dst = p.o[0].p + b * p.o[0].sB + fo * p.o[0].sF + (long long)t * p.o[0].sT;
    if (p.accum[0]) { a0 += dst[0]; a1 += dst[1]; }
    *reinterpret_cast<float2*>(dst) = make_float2(a0, a1);
}

// ---------------------------------------------------------------------------------------------------
// T2: weight gradient with a 2-channel side.  One pass over the wide tensor (lane = wide channel); the
// 2-channel tensor's rows are staged in shared memory and read at the tap offsets (broadcast).
//   wide_is_g = 1 (encoder 0): wide = G[b, j, t, n<32],  small = A[b, j*a_mul+a_off, t+dt, c<2]   -> dW[tap][c][n]
//   wide_is_g = 0 (decoder 5): wide = A[b, j, u, k<64],  small = G[b, j*g_mul+g_off, u-dt, c<2]   -> dW[tap][k][c]
// (for the second form the loop variable is the wide tensor's own time u = t + dt)
// ---------------------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------------------
// T4: decoder-5 forward = ConvTranspose (5,2)/(2,1) with N = 2 outputs from two 32-channel sources.
// Scatter form: y[b, 2j+kf-2, t+kt, o] += sum_c x_s[b, j, t, c] * W[kf*2+kt][s*32+c][o].
// One thread owns one (b, t) column of a strip of input rows and walks j; its 32 input channels sit in
// registers, the 1280 weights are FFMA constant-bank operands (no shared-memory traffic at all), and the five
// output rows an input row touches live in a sliding register window: after step j the rows 2j-2 and 2j-1
// are complete and leave through one shuffle (the kt = 1 half belongs to the neighbouring frame).
// Every input element is loaded exactly once (the gather form re-read it for each of its 10 taps).
// ---------------------------------------------------------------------------------------------------
__constant__ float c_upW[10 * 64 * 2];     // [slab = kf*2+kt][k = s*32+c][o]

struct UpN2Params {
    const float *x0, *x1;   // [B][F][T][32]
    const float* bias;      // [2]
    float* y;               // [B][2F][T+1][2]
    int B, F, T, JC;
};

template <int S>
__device__ __forceinline__ void up_n2_fma(const float4 (&x)[8], float (&acc)[5][2][2]) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        const float xv[4] = {x[c4].x, x[c4].y, x[c4].z, x[c4].w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int kf = 0; kf < 5; ++kf)
#pragma unroll
                for (int kt = 0; kt < 2; ++kt)
#pragma unroll
                    for (int o = 0; o < 2; ++o)
                        acc[kf][kt][o] = fmaf(xv[i], c_upW[((kf * 2 + kt) * 64 + S * 32 + c4 * 4 + i) * 2 + o], acc[kf][kt][o]);
    }
}

__global__ void __launch_bounds__(128) up_n2_kernel(const UpN2Params p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * 4 + warp;
    const int t = strip * 31 - 1 + lane;           // input frame of this lane (lane 0 = halo of the strip)
    const int b = blockIdx.z;
    const int j0 = blockIdx.y * p.JC, j1 = min(p.F, j0 + p.JC);
    if (strip * 31 > p.T) return;                  // whole warp beyond the T+1 output frames
    const bool tin = t >= 0 && t < p.T;
    const float b0 = p.bias ? __ldg(p.bias) : 0.f, b1 = p.bias ? __ldg(p.bias + 1) : 0.f;
    float acc[5][2][2];
#pragma unroll
    for (int a = 0; a < 5; ++a) acc[a][0][0] = acc[a][0][1] = acc[a][1][0] = acc[a][1][1] = 0.f;

    auto load = [&](const float* x, int j, float4 (&v)[8]) {
        if (tin && j >= 0 && j < p.F) {
            const float4* q = reinterpret_cast<const float4*>(x + (((long long)b * p.F + j) * p.T + t) * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(q + i);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    float4 xa[8], xb[8];
    load(p.x0, j0 - 1, xa);
    for (int j = j0 - 1; j <= j1; ++j) {
        load(p.x1, j, xb);                         // in flight while source 0 is consumed
        up_n2_fma<0>(xa, acc);
        load(p.x0, j + 1, xa);                     // in flight while source 1 is consumed
        up_n2_fma<1>(xb, acc);
        // rows 2j-2 (window 0) and 2j-1 (window 1) are complete
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const float n0 = __shfl_up_sync(0xffffffffu, acc[w][1][0], 1);
            const float n1 = __shfl_up_sync(0xffffffffu, acc[w][1][1], 1);
            const int row = 2 * j - 2 + w;
            if (lane >= 1 && t <= p.T && row >= 2 * j0 && row < 2 * j1)
                *reinterpret_cast<float2*>(p.y + (((long long)b * 2 * p.F + row) * (p.T + 1) + t) * 2) =
                    make_float2(acc[w][0][0] + n0 + b0, acc[w][0][1] + n1 + b1);
        }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                acc[0][kt][o] = acc[2][kt][o];
                acc[1][kt][o] = acc[3][kt][o];
                acc[2][kt][o] = acc[4][kt][o];
                acc[3][kt][o] = 0.f;
                acc[4][kt][o] = 0.f;
            }
    }
}

// distinct rows of the 2-channel operand touched by the taps (host-side dedupe of a_off / g_off)
struct SkinnyAux {
    int nslots;
    int slot_off[SEFD_MAX_TAPS];     // row offset of slot s: row = j * mul + slot_off[s]
    int tap_slot[SEFD_MAX_TAPS];
    int tap_dt[SEFD_MAX_TAPS];       // time of the 2-channel operand relative to the wide operand's time
};

constexpr int SW_CH = 128;           // wide-operand positions per work item

template <int PER, bool WIDE_IS_G, int NT>     // PER = wide channels per lane (WIDE / 32); NT = taps (compile time: no guards)
__global__ void __launch_bounds__(256, 2) smallside_wgrad_kernel(const WgradParams p, const SkinnyAux aux, int chunks) {
    constexpr int WIDE = 32 * PER;
    constexpr int SLD = SW_CH + 4;                       // staged times [u_lo - 2, u_lo + SW_CH + 2)
    constexpr int NPRE = (SEFD_MAX_TAPS * SLD + 255) / 256;
    __shared__ float2 srow[SEFD_MAX_TAPS][SLD];
    __shared__ float sred[SEFD_MAX_TAPS * 2 * WIDE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TapSrc& small = WIDE_IS_G ? p.a[0] : p.g;
    const int Fs = WIDE_IS_G ? p.Fa : p.Fg, Ts = WIDE_IS_G ? p.Ta : p.Tg;
    const int Tw = WIDE_IS_G ? p.Tg : p.Ta;              // wide tensor time extent
    const int smul = WIDE_IS_G ? p.a_mul : p.g_mul;
    float acc[NT][2][PER];
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int q = 0; q < PER; ++q) acc[a][c][q] = 0.f;

    // work item = (b, j) row x chunk of SW_CH positions.  The rows of the 2-channel operand the taps touch are
    // staged in shared memory (prefetched into registers one item ahead), the wide operand is streamed once with
    // UN x PER 128-byte requests in flight per warp.
    const long long items = (long long)p.B * p.J * chunks;
    const int nstage = aux.nslots * SLD;
    float2 pre[NPRE];
    auto fetch = [&](long long item) {
        const int r = (int)(item / chunks), ch = (int)(item % chunks);
        const int b = r / p.J, j = r % p.J;
        const int u_lo = ch * SW_CH;
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const int idx = tid + q * 256;
            pre[q] = make_float2(0.f, 0.f);
            if (idx < nstage) {
                const int slot = idx / SLD, i = idx - slot * SLD;
                const int fs = j * smul + aux.slot_off[slot], ts = u_lo - 2 + i;
                if (fs >= 0 && fs < Fs && ts >= 0 && ts < Ts)
                    pre[q] = __ldg(reinterpret_cast<const float2*>(small.p + b * small.sB + fs * small.sF + (long long)ts * small.sT));
            }
        }
    };
    constexpr int PW = SW_CH / 8;                            // positions per warp and item
    // staged-row element of tap `tap` for this warp's position i: sbase[soff[tap] + i]  (time u <-> index u - u_lo + 2)
    const float2* sbase = &srow[0][0];
    int soff[NT];
#pragma unroll
    for (int tap = 0; tap < NT; ++tap) soff[tap] = aux.tap_slot[tap] * SLD + warp * PW + 2 + aux.tap_dt[tap];
    long long item = blockIdx.x;
    if (item < items) fetch(item);
    for (; item < items; item += gridDim.x) {
        const int r = (int)(item / chunks), ch = (int)(item % chunks);
        const int b = r / p.J, j = r % p.J;
        const int u_lo = ch * SW_CH, u_hi = min(Tw, u_lo + SW_CH);
        const int fw = WIDE_IS_G ? j * p.g_mul + p.g_off[0] : j * p.a_mul + p.a_off[0];
        const float* wbase[PER];
        if (WIDE_IS_G) {
#pragma unroll
            for (int q = 0; q < PER; ++q) wbase[q] = p.g.p + b * p.g.sB + fw * p.g.sF + lane + 32 * q;
        } else {
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int k = lane + 32 * q;
                wbase[q] = k < p.a[0].C ? p.a[0].p + b * p.a[0].sB + fw * p.a[0].sF + k
                                        : p.a[1].p + b * p.a[1].sB + fw * p.a[1].sF + (k - p.a[0].C);
            }
        }
        const long long wsT = WIDE_IS_G ? p.g.sT : p.a[0].sT;    // both A sources share the time stride here
        const int uw = u_lo + warp * PW;
        // all of this warp's wide-operand loads for the item are issued BEFORE the staging barriers: PW x PER
        // 128-byte requests per warp are in flight while the 2-channel rows are written to shared memory
        float wv[PW][PER];
#pragma unroll
        for (int i = 0; i < PW; ++i)
#pragma unroll
            for (int q = 0; q < PER; ++q)
                wv[i][q] = uw + i < u_hi ? __ldg(wbase[q] + (long long)(uw + i) * wsT) : 0.f;
        __syncthreads();                                 // the previous item's readers are done
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const int idx = tid + q * 256;
            if (idx < nstage) (&srow[0][0])[idx] = pre[q];
        }
        __syncthreads();
        if (item + gridDim.x < items) fetch(item + gridDim.x);
#pragma unroll
        for (int i = 0; i < PW; ++i) {
            float2 sv[NT];
#pragma unroll
            for (int tap = 0; tap < NT; ++tap) sv[tap] = sbase[soff[tap] + i];     // LDS.64 [reg + immediate], all independent
#pragma unroll
            for (int tap = 0; tap < NT; ++tap)
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    acc[tap][0][q] = fmaf(sv[tap].x, wv[i][q], acc[tap][0][q]);
                    acc[tap][1][q] = fmaf(sv[tap].y, wv[i][q], acc[tap][1][q]);
                }
        }
    }
    // reduce the 8 warps through shared memory, then one atomic per output element and CTA
    __syncthreads();
    for (int i = tid; i < p.ntaps * 2 * WIDE; i += 256) sred[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < NT; ++tap)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int q = 0; q < PER; ++q) atomicAdd(&sred[(tap * 2 + c) * WIDE + lane + 32 * q], acc[tap][c][q]);
    __syncthreads();
    for (int i = tid; i < p.ntaps * 2 * WIDE; i += 256) {
        const int tap = i / (2 * WIDE), c = (i / WIDE) % 2, w = i % WIDE;
        // dW[slab][k][n]: encoder 0 -> k = c (2), n = w (WIDE) ; decoder 5 -> k = w (WIDE), n = c (2)
        const long long off = WIDE_IS_G ? ((long long)p.wslab[tap] * 2 + c) * WIDE + w
                                        : ((long long)p.wslab[tap] * WIDE + w) * 2 + c;
        atomicAdd(p.dW + off, sred[i]);
    }
}

}  // namespace

bool sefd_skinny_conv_eligible(const TapGemmParams& p) {
    const int N = p.o[0].N + p.o[1].N, K = p.a[0].C + p.a[1].C;
    if (p.wJ != 0 || p.bJ != 0) return false;
    if (K == 2 && p.a[1].C == 0 && (N == 32 || N == 64) && p.o[0].N % 4 == 0 && p.a[0].sT % 2 == 0 && p.a[0].sF % 2 == 0 &&
        p.a[0].sB % 2 == 0 && p.ntaps == 10) {
        for (int i = 0; i < 10; ++i)
            if (p.wslab[i] != i) return false;           // the constant-bank kernel takes the 10 slabs in order
        return true;
    }
    if (N == 2 && p.o[1].N == 0 && !p.stats && p.a[0].C % 4 == 0 && p.a[1].C % 4 == 0 && K <= 128) return true;
    return false;
}

int sefd_skinny_conv(const TapGemmParams& p, cudaStream_t st) {
    const int N = p.o[0].N + p.o[1].N, K = p.a[0].C + p.a[1].C;
    const long long blocks = (long long)((p.Tout + 127) / 128) * p.B * p.J;
    const double pos = (double)p.B * p.J * p.Tout;
    sefd_prof_label("skinny_conv K%d N%d taps%d J%d Tout%d", K, N, p.ntaps, p.J, p.Tout);
    SefdProfScope prof(SEFD_PROF_SKINNY, 2.0 * pos * N * K * p.ntaps,
                       4.0 * ((double)p.B * p.J * (p.fi_mul > 1 ? p.fi_mul : 1) * p.Tin * K + pos * N), st);
    if (K == 2) {
        cudaError_t e = cudaMemcpyToSymbolAsync(c_smallkW, p.W, sizeof(float) * 10 * 2 * N, 0, cudaMemcpyDeviceToDevice, st);
        SEFD_REQUIRE(e == cudaSuccess, "skinny_conv: constant upload failed: %s", cudaGetErrorString(e));
        if (N == 32) smallk_conv_kernel<32><<<(unsigned)blocks, 128, 0, st>>>(p);
        else smallk_conv_kernel<64><<<(unsigned)blocks, 128, 0, st>>>(p);
    } else {
        smalln_conv_kernel<<<(unsigned)blocks, 128, sizeof(float) * p.ntaps * K * 2, st>>>(p);
    }
    return sefd_check_launch("skinny_conv");
}


// decoder-5 forward (both output-row phases in one launch).  W is the packed [10][64][2] operand (device memory);
// it is copied into the constant bank in stream order.
bool sefd_skinny_up_n2_eligible(int Ch, int Cout) { return Ch == 32 && Cout == 2; }

int sefd_skinny_up_n2(const float* x0, const float* x1, const float* W, const float* bias, float* y, int B, int F, int T,
                      cudaStream_t st) {
    SEFD_REQUIRE((((uintptr_t)x0 | (uintptr_t)x1) & 15) == 0 && ((uintptr_t)y & 7) == 0, "skinny_up_n2: misaligned tensors");
    sefd_prof_label("skinny_up_n2 K64 N2 taps10 J%d Tout%d", F, T + 1);
    const double pos = (double)B * F * T;
    SefdProfScope prof(SEFD_PROF_SKINNY, 2.0 * pos * 64 * 2 * 10, 4.0 * (pos * 64 + (double)B * 2 * F * (T + 1) * 2), st);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_upW, W, sizeof(float) * 10 * 64 * 2, 0, cudaMemcpyDeviceToDevice, st);
    SEFD_REQUIRE(e == cudaSuccess, "skinny_up_n2: constant upload failed: %s", cudaGetErrorString(e));
    UpN2Params p;
    p.x0 = x0; p.x1 = x1; p.bias = bias; p.y = y; p.B = B; p.F = F; p.T = T;
    p.JC = F >= 64 ? 16 : (F >= 8 ? 8 : F);
    const int strips = (T + 1 + 30) / 31;
    dim3 grid((strips + 3) / 4, (F + p.JC - 1) / p.JC, B);
    up_n2_kernel<<<grid, 128, 0, st>>>(p);
    return sefd_check_launch("skinny_up_n2");
}

bool sefd_skinny_wgrad_eligible(const WgradParams& p) {
    const int K = p.a[0].C + p.a[1].C, N = p.g.C;
    if (p.ntaps != 10) return false;              // the kernel is instantiated for the 10-tap (5,2) conv
    if (K == 2 && p.a[1].C == 0 && N == 32) {          // wide = G: all taps must share the G row and time
        for (int i = 0; i < p.ntaps; ++i)
            if (p.g_off[i] != p.g_off[0]) return false;
        return true;
    }
    if (N == 2 && (K == 64 || K == 32) && p.a[0].C % 32 == 0) {   // wide = A: all taps must share the A row
        for (int i = 0; i < p.ntaps; ++i)
            if (p.a_off[i] != p.a_off[0]) return false;
        return true;
    }
    return false;
}

int sefd_skinny_wgrad(const WgradParams& p, cudaStream_t st) {
    const int K = p.a[0].C + p.a[1].C, N = p.g.C;
    const int wide_is_g = K == 2;
    const int Tw = wide_is_g ? p.Tg : p.Ta;
    const int chunks = (Tw + SW_CH - 1) / SW_CH;
    const long long items = (long long)p.B * p.J * chunks;
    long long ctas = 148 * 2;
    if (ctas > items) ctas = items;
    SkinnyAux aux;
    memset(&aux, 0, sizeof(aux));
    for (int t = 0; t < p.ntaps; ++t) {
        const int off = wide_is_g ? p.a_off[t] : p.g_off[t];
        int s = -1;
        for (int q = 0; q < aux.nslots; ++q)
            if (aux.slot_off[q] == off) s = q;
        if (s < 0) { s = aux.nslots++; aux.slot_off[s] = off; }
        aux.tap_slot[t] = s;
        // 2-channel operand time: A at t+dt when the wide operand is G (t = u); G at u-dt when the wide operand is A
        aux.tap_dt[t] = wide_is_g ? p.dt[t] : -p.dt[t];
        SEFD_REQUIRE(aux.tap_dt[t] >= -2 && aux.tap_dt[t] <= 2, "skinny_wgrad: time shift %d beyond the staged halo", aux.tap_dt[t]);
    }
    const double pos = (double)p.B * p.J * p.Tg;
    sefd_prof_label("skinny_wgrad K%d N%d taps%d J%d", K, N, p.ntaps, p.J);
    SefdProfScope prof(SEFD_PROF_SKINNY, 2.0 * pos * K * N * p.ntaps,
                       4.0 * ((double)p.B * p.J * (p.a_mul > 1 ? p.a_mul : 1) * p.Ta * K +
                              (double)p.B * p.J * (p.g_mul > 1 ? p.g_mul : 1) * p.Tg * N), st);
    if (wide_is_g) smallside_wgrad_kernel<1, true, 10><<<(unsigned)ctas, 256, 0, st>>>(p, aux, chunks);
    else if (K == 64) smallside_wgrad_kernel<2, false, 10><<<(unsigned)ctas, 256, 0, st>>>(p, aux, chunks);
    else smallside_wgrad_kernel<1, false, 10><<<(unsigned)ctas, 256, 0, st>>>(p, aux, chunks);
    return sefd_check_launch("skinny_wgrad");
}
