// Generic fp32 "tap-GEMM" on CUDA cores: the exact-precision implicit-GEMM that every
// convolution-like contraction on the DCCRN path can be expressed as (conv / convT forward,
// their data gradients, LSTM input projections, the output Linear) plus the matching weight
// gradient.  It is im2col-free: the K loop walks (tap, channel) and reads shifted rows of the
// channels-last activation directly.  This is the reference-precision engine; the tensor-core
// (tcgen05, TF32) engine in tapgemm_tc.cu implements the same contract for the wide layers.
#include "common.cuh"
#include "prof.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int AS_LD = BM + 4;

__global__ void __launch_bounds__(NT) tapgemm_simt_kernel(const TapGemmParams p) {
    __shared__ __align__(16) float As[BK][AS_LD];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ float red[16][BN];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ttiles = (p.Tout + BM - 1) / BM;
    const int t0 = (blockIdx.x % ttiles) * BM;
    const int row = blockIdx.x / ttiles;
    const int b = row / p.J, j = row % p.J;
    const int n0 = blockIdx.y * BN;
    const int C0 = p.a[0].C, C1 = p.a[1].C;
    const int K = C0 + C1;
    const int N0 = p.o[0].N;
    const int N = N0 + p.o[1].N;
    const bool vecA = (C0 % 4 == 0) && (C1 % 4 == 0);
    const bool vecN = (N % 4 == 0) && (N0 % 4 == 0);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

    for (int tap = 0; tap < p.ntaps; ++tap) {
        const int fi = j * p.fi_mul + p.df[tap];
        if (fi < 0 || fi >= p.Fin) continue;   // CTA-uniform
        const int dt = p.dt[tap];
        const float* Wt = p.W + p.wJ * j + (long long)p.wslab[tap] * K * N;
        const float* a0 = p.a[0].p + b * p.a[0].sB + fi * p.a[0].sF;
        const float* a1 = C1 ? p.a[1].p + b * p.a[1].sB + fi * p.a[1].sF : nullptr;
        for (int k0 = 0; k0 < K; k0 += BK) {
            if (vecA) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int idx = tid + i * NT;
                    const int r = idx >> 2, kq = idx & 3;
                    const int k = k0 + kq * 4;
                    const int tin = t0 + r + dt;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k < K && tin >= 0 && tin < p.Tin) {
                        const float* src = (k < C0) ? a0 + (long long)tin * p.a[0].sT + k
                                                    : a1 + (long long)tin * p.a[1].sT + (k - C0);
                        v = __ldg(reinterpret_cast<const float4*>(src));
                    }
                    As[kq * 4 + 0][r] = v.x;
                    As[kq * 4 + 1][r] = v.y;
                    As[kq * 4 + 2][r] = v.z;
                    As[kq * 4 + 3][r] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = tid + i * NT;
                    const int r = idx >> 4, kk = idx & 15;
                    const int k = k0 + kk;
                    const int tin = t0 + r + dt;
                    float v = 0.f;
                    if (k < K && tin >= 0 && tin < p.Tin) {
                        v = (k < C0) ? __ldg(a0 + (long long)tin * p.a[0].sT + k)
                                     : __ldg(a1 + (long long)tin * p.a[1].sT + (k - C0));
                    }
                    As[kk][r] = v;
                }
            }
            {
                const int r = tid >> 4, c4 = (tid & 15) * 4;
                const int k = k0 + r, n = n0 + c4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < K) {
                    const float* src = Wt + (long long)k * N + n;
                    if (vecN && n + 3 < N) {
                        v = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        if (n + 0 < N) v.x = __ldg(src + 0);
                        if (n + 1 < N) v.y = __ldg(src + 1);
                        if (n + 2 < N) v.z = __ldg(src + 2);
                        if (n + 3 < N) v.w = __ldg(src + 3);
                    }
                }
                *reinterpret_cast<float4*>(&Bs[r][c4]) = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 alo = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
                const float4 ahi = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float a[8] = {alo.x, alo.y, alo.z, alo.w, ahi.x, ahi.y, ahi.z, ahi.w};
                const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(a[i], bb[c], acc[i][c]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue: bias, optional accumulate, store, optional per-channel statistics ----
    const int fo = j * p.fo_mul + p.fo_off;
    const int n = n0 + tx * 4;
    float bsv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (n + c < N) bsv[c] = __ldg(p.bias + p.bJ * j + n + c);
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + ty * 8 + i;
        if (t >= p.Tout) continue;
        if (vecN && n + 3 < N) {
            const int d = (n < N0) ? 0 : 1;
            const int nn = d ? n - N0 : n;
            float* dst = p.o[d].p + b * p.o[d].sB + fo * p.o[d].sF + (long long)t * p.o[d].sT + nn;
            float4 v = make_float4(acc[i][0] + bsv[0], acc[i][1] + bsv[1], acc[i][2] + bsv[2], acc[i][3] + bsv[3]);
            if (p.accum[d]) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            if (p.round_out[d]) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }
            *reinterpret_cast<float4*>(dst) = v;
            s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
            s2[0] += v.x * v.x; s2[1] += v.y * v.y; s2[2] += v.z * v.z; s2[3] += v.w * v.w;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (n + c >= N) continue;
                const int d = (n + c < N0) ? 0 : 1;
                const int nn = d ? n + c - N0 : n + c;
                float* dst = p.o[d].p + b * p.o[d].sB + fo * p.o[d].sF + (long long)t * p.o[d].sT + nn;
                float v = acc[i][c] + bsv[c];
                if (p.accum[d]) v += *dst;
                *dst = v;
                s1[c] += v;
                s2[c] += v * v;
            }
        }
    }
    if (p.stats) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[ty][tx * 4 + c] = s1[c];
        __syncthreads();
        if (tid < BN) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) s += red[r][tid];
            if (n0 + tid < N) atomicAdd(p.stats + n0 + tid, (double)s);
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 4; ++c) red[ty][tx * 4 + c] = s2[c];
        __syncthreads();
        if (tid < BN) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) s += red[r][tid];
            if (n0 + tid < N) atomicAdd(p.stats + N + n0 + tid, (double)s);
        }
    }
}

// ---------------------------------------------------------------------------------------
constexpr int WK = 64, WN = 64, WP = 16;

__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradParams p) {
    __shared__ __align__(16) float As[WP][WK];
    __shared__ __align__(16) float Gs[WP][WN];

    const int tid = threadIdx.x, tk = tid >> 4, tn = tid & 15;
    const int C0 = p.a[0].C, C1 = p.a[1].C;
    const int K = C0 + C1, N = p.g.C;
    const int nnt = (N + WN - 1) / WN;
    const int k0 = (blockIdx.x / nnt) * WK, n0 = (blockIdx.x % nnt) * WN;
    const int tap = blockIdx.y;
    const int r0 = blockIdx.z * p.rows_per_cta;
    const int r1 = min(p.B * p.J, r0 + p.rows_per_cta);
    const bool vecA = (C0 % 4 == 0) && (C1 % 4 == 0);
    const bool vecN = (N % 4 == 0);
    const int dt = p.dt[tap];

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

    const int pos = tid >> 4, q = (tid & 15) * 4;
    for (int r = r0; r < r1; ++r) {
        const int b = r / p.J, j = r % p.J;
        const int fa = j * p.a_mul + p.a_off[tap];
        const int fg = j * p.g_mul + p.g_off[tap];
        if (fa < 0 || fa >= p.Fa || fg < 0 || fg >= p.Fg) continue;
        const float* a0 = p.a[0].p + b * p.a[0].sB + fa * p.a[0].sF;
        const float* a1 = C1 ? p.a[1].p + b * p.a[1].sB + fa * p.a[1].sF : nullptr;
        const float* g0 = p.g.p + b * p.g.sB + fg * p.g.sF;
        for (int tt = 0; tt < p.Tg; tt += WP) {
            const int t = tt + pos, ta = t + dt;
            {
                const int k = k0 + q;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < p.Tg && ta >= 0 && ta < p.Ta && k < K) {
                    if (vecA) {
                        const float* src = (k < C0) ? a0 + (long long)ta * p.a[0].sT + k
                                                    : a1 + (long long)ta * p.a[1].sT + (k - C0);
                        v = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int kc = k + c;
                            if (kc < K)
                                e[c] = (kc < C0) ? __ldg(a0 + (long long)ta * p.a[0].sT + kc)
                                                 : __ldg(a1 + (long long)ta * p.a[1].sT + (kc - C0));
                        }
                        v = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
                *reinterpret_cast<float4*>(&As[pos][q]) = v;
            }
            {
                const int n = n0 + q;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < p.Tg && n < N) {
                    const float* src = g0 + (long long)t * p.g.sT + n;
                    if (vecN) {
                        v = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        if (n + 0 < N) v.x = __ldg(src + 0);
                        if (n + 1 < N) v.y = __ldg(src + 1);
                        if (n + 2 < N) v.z = __ldg(src + 2);
                        if (n + 3 < N) v.w = __ldg(src + 3);
                    }
                }
                *reinterpret_cast<float4*>(&Gs[pos][q]) = v;
            }
            __syncthreads();
#pragma unroll
            for (int pp = 0; pp < WP; ++pp) {
                const float4 av = *reinterpret_cast<const float4*>(&As[pp][tk * 4]);
                const float4 gv = *reinterpret_cast<const float4*>(&Gs[pp][tn * 4]);
                const float a[4] = {av.x, av.y, av.z, av.w};
                const float g[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(a[i], g[c], acc[i][c]);
            }
            __syncthreads();
        }
    }
    float* dW = p.dW + (long long)p.wslab[tap] * K * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + tk * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int n = n0 + tn * 4 + c;
            if (n < N) atomicAdd(dW + (long long)k * N + n, acc[i][c]);
        }
    }
}

}  // namespace

int sefd_tapgemm_simt(const TapGemmParams& p, cudaStream_t st) {
    const int N = p.o[0].N + p.o[1].N;
    SEFD_REQUIRE(p.ntaps >= 1 && p.ntaps <= SEFD_MAX_TAPS, "tapgemm: bad tap count %d", p.ntaps);
    SEFD_REQUIRE(p.B > 0 && p.J > 0 && p.Tout > 0 && N > 0, "tapgemm: empty problem");
    const long long ttiles = (p.Tout + BM - 1) / BM;
    const long long gx = ttiles * p.B * p.J;
    SEFD_REQUIRE(gx < (1ll << 31), "tapgemm: grid too large");
    dim3 grid((unsigned)gx, (unsigned)((N + BN - 1) / BN));
    const double K = p.a[0].C + p.a[1].C, pos = (double)p.B * p.J * p.Tout;
    sefd_prof_label("tapgemm_simt K%d N%d taps%d J%d Tout%d", (int)K, N, p.ntaps, p.J, p.Tout);
    SefdProfScope prof(SEFD_PROF_TAPGEMM, 2.0 * pos * N * K * p.ntaps,
                       4.0 * ((double)p.B * p.J * (p.fi_mul > 1 ? p.fi_mul : 1) * p.Tin * K + pos * N), st);
    tapgemm_simt_kernel<<<grid, NT, 0, st>>>(p);
    return sefd_check_launch("tapgemm_simt");
}

int sefd_wgrad_simt(const WgradParams& p_in, cudaStream_t st) {
    WgradParams p = p_in;
    const int K = p.a[0].C + p.a[1].C, N = p.g.C;
    SEFD_REQUIRE(p.ntaps >= 1 && p.ntaps <= SEFD_MAX_TAPS, "wgrad: bad tap count %d", p.ntaps);
    const int tiles = ((K + WK - 1) / WK) * ((N + WN - 1) / WN);
    const int rows = p.B * p.J;
    int splits = (148 * 8 + tiles * p.ntaps - 1) / (tiles * p.ntaps);
    if (splits < 1) splits = 1;
    if (splits > rows) splits = rows;
    p.rows_per_cta = (rows + splits - 1) / splits;
    splits = (rows + p.rows_per_cta - 1) / p.rows_per_cta;
    SEFD_REQUIRE(splits <= 65535, "wgrad: too many splits");
    dim3 grid(tiles, p.ntaps, splits);
    const double pos = (double)p.B * p.J * p.Tg;
    sefd_prof_label("wgrad_simt K%d N%d taps%d J%d splits%d", K, N, p.ntaps, p.J, splits);
    SefdProfScope prof(SEFD_PROF_WGRAD, 2.0 * pos * K * N * p.ntaps,
                       4.0 * ((double)p.B * p.J * (p.a_mul > 1 ? p.a_mul : 1) * p.Ta * K +
                              (double)p.B * p.J * (p.g_mul > 1 ? p.g_mul : 1) * p.Tg * N), st);
    wgrad_simt_kernel<<<grid, 256, 0, st>>>(p);
    return sefd_check_launch("wgrad_simt");
}
