// HBM-bound passes of the DCCRN step: BatchNorm(train/eval)+PReLU forward and backward,
// packing of the complex conv weights into the block-real GEMM operand (and folding the
// block-real weight gradient back), complex-LSTM output combination, fused Adam.
#include "elementwise.cuh"
#include "prof.cuh"

namespace {

constexpr int MAXC = 512;

// ------------------------------------------------------------------------------------
// BatchNorm2d + PReLU forward  (reference: models.py:76-78; nn.BatchNorm2d eps 1e-5, momentum 0.1)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_prelu_fwd_kernel(const BnPreluFwdParams p) {
    __shared__ float s_scale[MAXC], s_shift[MAXC];
    const int C = p.C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, invstd;
        if (p.use_running) {
            mean = p.running_mean[c];
            invstd = (float)(1.0 / sqrt((double)p.running_var[c] + (double)p.eps));
        } else {
            const double m = p.stats[c] / p.n_stat;
            double var = p.stats[C + c] / p.n_stat - m * m;
            if (var < 0) var = 0;
            mean = (float)m;
            invstd = (float)(1.0 / sqrt(var + (double)p.eps));
            if (blockIdx.x == 0) {
                p.save[c] = mean;
                p.save[C + c] = invstd;
                if (p.running_mean) {
                    const double unb = var * p.n_stat / fmax(p.n_stat - 1.0, 1.0);
                    p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
                    p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unb;
                }
            }
        }
        const float sc = p.gamma[c] * invstd;
        s_scale[c] = sc;
        s_shift[c] = p.beta[c] - mean * sc;
    }
    __syncthreads();
    const float alpha = p.alpha[0];
    const int C4 = C >> 2;
    const long long total = (long long)p.BF * p.T * C4;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(e % C4);
        const long long row = e / C4;
        const int t = (int)(row % p.T);
        const long long bf = row / p.T;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.y + ((bf * p.Ty + t + p.tshift) * C) + c4 * 4));
        float4 o;
        const int c = c4 * 4;
        o.x = fmaf(v.x, s_scale[c + 0], s_shift[c + 0]);
        o.y = fmaf(v.y, s_scale[c + 1], s_shift[c + 1]);
        o.z = fmaf(v.z, s_scale[c + 2], s_shift[c + 2]);
        o.w = fmaf(v.w, s_scale[c + 3], s_shift[c + 3]);
        o.x = o.x > 0.f ? o.x : alpha * o.x;
        o.y = o.y > 0.f ? o.y : alpha * o.y;
        o.z = o.z > 0.f ? o.z : alpha * o.z;
        o.w = o.w > 0.f ? o.w : alpha * o.w;
        if (p.round_tf32) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
        *reinterpret_cast<float4*>(p.z + e * 4) = o;
    }
}

// ------------------------------------------------------------------------------------
// backward, pass 1: per-channel sums of g' and g'*xhat, scalar d(alpha)
//   u = gamma*xhat+beta ; z = prelu(u) ; g' = dz * (u>0 ? 1 : alpha)
// red layout (double): [0..C) sum g', [C..2C) sum g'*xhat, [2C] d alpha
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) bn_prelu_bwd_reduce_kernel(const BnPreluBwdParams p) {
    __shared__ double s_red[3][256][4];          // [which][thread][4]; sums cancel heavily -> double throughout
    const int C = p.C, C4 = C >> 2;
    const int lanes = 256 / C4;                  // rows processed per block iteration
    const int c4 = threadIdx.x % C4;
    const int rl = threadIdx.x / C4;
    const int c = c4 * 4;
    float mean[4], istd[4], gam[4], bet[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mean[i] = p.save[c + i];
        istd[i] = p.save[C + c + i];
        gam[i] = p.gamma[c + i];
        bet[i] = p.beta[c + i];
    }
    const float alpha = p.alpha[0];
    double sg[4] = {0, 0, 0, 0}, sgx[4] = {0, 0, 0, 0}, sa = 0.0;
    const long long rows = (long long)p.BF * p.T;
    if (rl < lanes) {
        // two adjacent rows per iteration, all loads issued before the first use: the pass is bound by the bytes a thread keeps
        // in flight (64 registers -> 4 CTAs per SM; one row per iteration left 32 KB per SM in flight and ran at 3.9 TB/s)
        for (long long row0 = 2 * ((long long)blockIdx.x * lanes + rl); row0 < rows; row0 += 2 * (long long)gridDim.x * lanes) {
            float4 yv[2], dv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const long long row = row0 + j;
                yv[j] = dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < rows) {
                    const int t = (int)(row % p.T);
                    const long long bf = row / p.T;
                    yv[j] = __ldg(reinterpret_cast<const float4*>(p.y + ((bf * p.Ty + t + p.tshift) * C) + c));
                    dv[j] = __ldg(reinterpret_cast<const float4*>(p.dz + row * C + c));
                    if (p.dz2) {
                        const float4 d2 = __ldg(reinterpret_cast<const float4*>(p.dz2 + row * C + c));
                        dv[j].x += d2.x; dv[j].y += d2.y; dv[j].z += d2.z; dv[j].w += d2.w;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float y4[4] = {yv[j].x, yv[j].y, yv[j].z, yv[j].w};
                const float d4[4] = {dv[j].x, dv[j].y, dv[j].z, dv[j].w};      // zero past the end: contributes nothing
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xh = (y4[i] - mean[i]) * istd[i];
                    const float u = fmaf(gam[i], xh, bet[i]);
                    const float g = u > 0.f ? d4[i] : alpha * d4[i];
                    sg[i] += (double)g;
                    sgx[i] += (double)(g * xh);
                    sa += u > 0.f ? 0.0 : (double)(d4[i] * u);
                }
            }
        }
    }
    // block reduction over the `lanes` row-lanes that share a channel quad
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s_red[0][threadIdx.x][i] = sg[i];
        s_red[1][threadIdx.x][i] = sgx[i];
    }
    s_red[2][threadIdx.x][0] = sa;
    __syncthreads();
    if (threadIdx.x < C4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double a = 0.0, b = 0.0;
            for (int l = 0; l < lanes; ++l) {
                a += s_red[0][l * C4 + threadIdx.x][i];
                b += s_red[1][l * C4 + threadIdx.x][i];
            }
            atomicAdd(p.red + threadIdx.x * 4 + i, a);
            atomicAdd(p.red + C + threadIdx.x * 4 + i, b);
        }
    }
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int l = 0; l < lanes * C4; ++l) a += s_red[2][l][0];
        atomicAdd(p.red + 2 * C, a);
    }
}

// backward, pass 2: dy = gamma*invstd*(g' - mean(g') - xhat*mean(g'*xhat)) over all Ty frames
// (frames outside the kept window carry g' = 0 but still receive the statistics terms);
// block 0 also emits d gamma, d beta, d alpha.
__global__ void __launch_bounds__(256) bn_prelu_bwd_apply_kernel(const BnPreluBwdParams p) {
    __shared__ float s_mg[MAXC], s_mgx[MAXC];
    const int C = p.C, C4 = C >> 2;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        s_mg[c] = (float)(p.red[c] / p.n_stat);
        s_mgx[c] = (float)(p.red[C + c] / p.n_stat);
        if (blockIdx.x == 0) {
            p.dgamma[c] = (float)p.red[C + c];
            p.dbeta[c] = (float)p.red[c];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.dalpha[0] = (float)p.red[2 * C];
    __syncthreads();
    const float alpha = p.alpha[0];
    const long long total = (long long)p.BF * p.Ty * C4;
    // the grid stride (gridDim * 256) is a multiple of C4 (C4 divides 256), so a thread keeps its channel quad for the whole
    // loop: the per-channel constants live in registers instead of being re-read for every element
    const int c = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % C4) * 4;
    float k_istd[4], k_mean[4], k_gam[4], k_bet[4], k_mg[4], k_mgx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        k_istd[i] = p.save[C + c + i];
        k_mean[i] = p.save[c + i];
        k_gam[i] = p.gamma[c + i];
        k_bet[i] = p.beta[c + i];
        k_mg[i] = s_mg[c + i];
        k_mgx[i] = s_mgx[c + i];
    }
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / C4;
        const int ty = (int)(row % p.Ty);
        const long long bf = row / p.Ty;
        const int t = ty - p.tshift;
        const float4 yv = __ldg(reinterpret_cast<const float4*>(p.y + e * 4));
        float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < p.T) {
            dv = __ldg(reinterpret_cast<const float4*>(p.dz + ((bf * p.T + t) * C) + c));
            if (p.dz2) {
                const float4 d2 = __ldg(reinterpret_cast<const float4*>(p.dz2 + ((bf * p.T + t) * C) + c));
                dv.x += d2.x; dv.y += d2.y; dv.z += d2.z; dv.w += d2.w;
            }
        }
        const float y4[4] = {yv.x, yv.y, yv.z, yv.w};
        const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float istd = k_istd[i];
            const float xh = (y4[i] - k_mean[i]) * istd;
            const float gam = k_gam[i];
            const float u = fmaf(gam, xh, k_bet[i]);
            const float g = u > 0.f ? d4[i] : alpha * d4[i];
            o[i] = gam * istd * (g - k_mg[i] - xh * k_mgx[i]);
            if (p.round_tf32) o[i] = tf32_rn(o[i]);
        }
        *reinterpret_cast<float4*>(p.dy + e * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------
// complex conv weights -> block-real GEMM operand  (tools_for_model.py:259-266, 328-335)
//   real_out = Wr*real_in - Wi*imag_in ; imag_out = Wi*real_in + Wr*imag_in
// K index order: one source [real Ci2 | imag Ci2]; two sources (skip concat, models.py:224 /
// tools_for_model.py:184-193): [src0: real h | imag h][src1: real h | imag h], h = Ci2/2,
// which corresponds to reference input channel  src*h + idx  inside each complex half.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void kmap(int k, int Ci2, int two_src, int& part, int& idx) {
    if (!two_src) {
        part = k / Ci2;
        idx = k % Ci2;
    } else {
        const int h = Ci2 >> 1;
        const int src = k / Ci2, kk = k % Ci2;
        part = kk / h;
        idx = src * h + kk % h;
    }
}
__device__ __forceinline__ int kinv(int part, int idx, int Ci2, int two_src) {
    if (!two_src) return part * Ci2 + idx;
    const int h = Ci2 >> 1;
    return (idx / h) * Ci2 + part * h + idx % h;
}

__global__ void pack_cconv_kernel(const CconvPackParams p) {
    const int K = 2 * p.Ci2, N = 2 * p.Co2;
    const long long total = 10ll * K * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e % N);
        const int k = (int)((e / N) % K);
        const int slab = (int)(e / ((long long)N * K));   // kf*2+kt
        int kp, ki;
        kmap(k, p.Ci2, p.two_src, kp, ki);
        const int np = n / p.Co2, ni = n % p.Co2;
        const long long widx = p.transposed ? ((long long)ki * p.Co2 + ni) * 10 + slab
                                            : ((long long)ni * p.Ci2 + ki) * 10 + slab;
        float v;
        if (kp == np) v = p.wr[widx];
        else if (kp == 1) v = -p.wi[widx];   // imag in -> real out
        else v = p.wi[widx];                 // real in -> imag out
        if (p.round_tf32) v = tf32_rn(v);
        p.Wf[e] = v;
        p.Wt[((long long)slab * N + n) * K + k] = v;
    }
    if (blockIdx.x == 0) {
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            const int np = n / p.Co2, ni = n % p.Co2;
            p.bias[n] = np == 0 ? p.br[ni] - p.bi[ni] : p.br[ni] + p.bi[ni];
        }
    }
}

__global__ void fold_cconv_kernel(const CconvFoldParams p) {
    const int K = 2 * p.Ci2, N = 2 * p.Co2;
    const long long total = 10ll * p.Ci2 * p.Co2;
    // thread order (slab, ki, ni) with ni fastest: the nsplit x 4 partial-gradient reads are coalesced; only the single
    // write per element is strided (parameter layout [..][..][5][2] has the tap index innermost)
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int ni = (int)(t % p.Co2);
        const int ki = (int)((t / p.Co2) % p.Ci2);
        const int slab = (int)(t / ((long long)p.Co2 * p.Ci2));
        const int kr = kinv(0, ki, p.Ci2, p.two_src), kim = kinv(1, ki, p.Ci2, p.two_src);
        float rr = 0.f, ii = 0.f, ri = 0.f, ir = 0.f;
        for (int s = 0; s < p.nsplit; ++s) {
            const float* d = p.dWf + s * p.split_stride + (long long)slab * K * N;
            rr += d[(long long)kr * N + ni];
            ii += d[(long long)kim * N + p.Co2 + ni];
            ri += d[(long long)kr * N + p.Co2 + ni];
            ir += d[(long long)kim * N + ni];
        }
        const long long e = p.transposed ? ((long long)ki * p.Co2 + ni) * 10 + slab : ((long long)ni * p.Ci2 + ki) * 10 + slab;
        p.dwr[e] = rr + ii;
        p.dwi[e] = ri - ir;
    }
    if (blockIdx.x == 0) {
        for (int n = threadIdx.x; n < p.Co2; n += blockDim.x) {
            const float dr = p.dbias ? p.dbias[n] : 0.f, di = p.dbias ? p.dbias[p.Co2 + n] : 0.f;
            p.dbr[n] = dr + di;
            p.dbi[n] = di - dr;
        }
    }
}

// real conv weights: slab = kf*2+kt is the innermost index of both torch layouts
__global__ void pack_rconv_kernel(const RconvPackParams p) {
    const int K = p.Ci, N = p.Co;
    const long long total = 10ll * K * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e % N);
        const int k = (int)((e / N) % K);
        const int slab = (int)(e / ((long long)N * K));
        const long long widx = p.transposed ? ((long long)k * N + n) * 10 + slab : ((long long)n * K + k) * 10 + slab;
        float v = p.w[widx];
        if (p.round_tf32) v = tf32_rn(v);
        p.Wf[e] = v;
        p.Wt[((long long)slab * N + n) * K + k] = v;
    }
}

__global__ void fold_rconv_kernel(const RconvFoldParams p) {
    const int K = p.Ci, N = p.Co;
    const long long total = 10ll * K * N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {          // (slab, k, n) order: coalesced partial reads
        const int n = (int)(t % N);
        const int k = (int)((t / N) % K);
        const int slab = (int)(t / ((long long)N * K));
        float a = 0.f;
        for (int s = 0; s < p.nsplit; ++s) a += p.dWf[s * p.split_stride + t];
        const long long e = p.transposed ? ((long long)k * N + n) * 10 + slab : ((long long)n * K + k) * 10 + slab;
        p.dw[e] = a;
    }
    if (blockIdx.x == 0)
        for (int n = threadIdx.x; n < N; n += blockDim.x) p.db[n] = p.dbias ? p.dbias[n] : 0.f;
}

// ------------------------------------------------------------------------------------
// generic strided 2-term gather:  dst[i] = c0 * src0[map(i)] (+ c1 * src1[map(i)])
// used for the LSTM / Linear weight (un)permutations, expressed as a 3-d index transpose:
//   dst[(a*nb + b)*nc + c] = src[a*sa + b*sb + c*sc]
// ------------------------------------------------------------------------------------
__global__ void permute3_kernel(const Permute3Params q) {
    const long long total = (long long)q.na * q.nb * q.nc;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % q.nc);
        const int b = (int)((e / q.nc) % q.nb);
        const int a = (int)(e / ((long long)q.nc * q.nb));
        const long long so = a * q.sa + b * q.sb + c * q.sc;
        float v = 0.f;
        for (int s = 0; s < q.nsplit; ++s) v += q.src[s * q.split_stride + so];
        if (q.round_tf32) v = tf32_rn(v);
        float* d = q.dst + a * q.da + b * q.db + c * q.dc;
        *d = q.accumulate ? *d + v : v;
    }
}

__global__ void add2_kernel(const float* a, const float* b, float* o, long long n) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        o[e] = a[e] + b[e];
}

// column sums over rows (o, i) of x[o*sO + i*sI + c], c < C.  Thread = (column, row lane); four independent loads per
// iteration, block-level reduction over the row lanes in shared memory, then ONE double atomic per column and block
// (per-thread atomics onto C addresses serialised: 0.14 ms for the 2-column bias gradient of decoder 5).
__global__ void __launch_bounds__(512) colsum2_kernel(const float* __restrict__ x, int nO, long long sO, long long nI, long long sI,
                                                      int C, double* out) {
    __shared__ float s_part[512];
    const int c = threadIdx.x % C;
    const int lane = threadIdx.x / C, lanes = blockDim.x / C;
    const long long rows = (long long)nO * nI;
    const long long step = (long long)gridDim.x * lanes;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (lane < lanes) {
        long long r = (long long)blockIdx.x * lanes + lane;
        for (; r + 3 * step < rows; r += 4 * step) {
            const long long r1 = r + step, r2 = r + 2 * step, r3 = r + 3 * step;
            const float a = __ldg(x + (r / nI) * sO + (r % nI) * sI + c);
            const float b = __ldg(x + (r1 / nI) * sO + (r1 % nI) * sI + c);
            const float d = __ldg(x + (r2 / nI) * sO + (r2 % nI) * sI + c);
            const float e = __ldg(x + (r3 / nI) * sO + (r3 % nI) * sI + c);
            s0 += a; s1 += b; s2 += d; s3 += e;
        }
        for (; r < rows; r += step) s0 += __ldg(x + (r / nI) * sO + (r % nI) * sI + c);
    }
    s_part[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.x < C) {
        double a = 0.0;
        for (int l = 0; l < lanes; ++l) a += (double)s_part[l * C + threadIdx.x];
        atomicAdd(out + threadIdx.x, a);
    }
}

__global__ void d2f_kernel(const double* s, float* d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = (float)s[i];
}

// ------------------------------------------------------------------------------------
// complex LSTM combination (tools_for_model.py:171-172)
//   H layout [lstm p][part q][B][T][H]:  p = 0 real_lstm / 1 imag_lstm, q = 0 real input / 1 imag input
//   real = H[0][0] - H[1][1] ;  imag = H[0][1] + H[1][0]      -> X [part][B][T][H]
// backward: dH[0][0] = dreal, dH[1][1] = -dreal, dH[0][1] = dH[1][0] = dimag  (accumulated into dH)
// ------------------------------------------------------------------------------------
__global__ void clstm_combine_kernel(const float* __restrict__ Hh, float* __restrict__ X, long long n, int round_tf32) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float r = Hh[e] - Hh[3 * n + e], i = Hh[n + e] + Hh[2 * n + e];
        if (round_tf32) { r = tf32_rn(r); i = tf32_rn(i); }
        X[e] = r;
        X[n + e] = i;
    }
}
__global__ void clstm_combine_bwd_kernel(const float* __restrict__ dX, float* __restrict__ dH, long long n) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float dr = dX[e], di = dX[n + e];
        dH[e] = dr;
        dH[3 * n + e] = -dr;
        dH[n + e] = di;
        dH[2 * n + e] = di;
    }
}

// ------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults of train_interface.py:59: betas (0.9,0.999), eps 1e-8, no decay)
// ------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt, float gscale) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float gr = g[e] * gscale;
        const float mm = b1 * m[e] + (1.f - b1) * gr;
        const float vv = b2 * v[e] + (1.f - b2) * gr * gr;
        m[e] = mm;
        v[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        w[e] -= (lr / bc1) * (mm / denom);
    }
}

// device-resident step counter (CUDA-graph friendly: nothing host-computed changes between steps): one thread bumps the
// counter and derives the bias corrections in double, like the host path does
__global__ void adam_prep_kernel(int* step, float* bc, double b1, double b2) {
    const int t = *step + 1;
    *step = t;
    bc[0] = (float)(1.0 - pow(b1, (double)t));
    bc[1] = (float)sqrt(1.0 - pow(b2, (double)t));
}
__global__ void adam_dev_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                const float* __restrict__ bc, float gscale) {
    const float bc1 = bc[0], bc2_sqrt = bc[1];
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float gr = g[e] * gscale;
        const float mm = b1 * m[e] + (1.f - b1) * gr;
        const float vv = b2 * v[e] + (1.f - b2) * gr * gr;
        m[e] = mm;
        v[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        w[e] -= (lr / bc1) * (mm / denom);
    }
}

inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
    long long g = (n + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

int sefd_bn_prelu_fwd(const BnPreluFwdParams& p, cudaStream_t st) {
    SEFD_REQUIRE(p.C % 4 == 0 && p.C <= MAXC, "bn_prelu_fwd: C=%d unsupported", p.C);
    sefd_prof_label("bn_prelu_fwd C%d rows%lld", p.C, (long long)p.BF * p.T);
    SefdProfScope prof(SEFD_PROF_BN, 0, 8.0 * p.BF * (double)p.T * p.C, st);
    bn_prelu_fwd_kernel<<<grid_for((long long)p.BF * p.T * (p.C / 4)), 256, 0, st>>>(p);
    return sefd_check_launch("bn_prelu_fwd");
}

int sefd_bn_prelu_bwd(const BnPreluBwdParams& p, cudaStream_t st) {
    SEFD_REQUIRE(p.C % 4 == 0 && p.C <= MAXC && 256 % (p.C / 4) == 0, "bn_prelu_bwd: C=%d unsupported", p.C);
    sefd_prof_label("bn_prelu_bwd C%d rows%lld", p.C, (long long)p.BF * p.T);
    SefdProfScope prof(SEFD_PROF_BN, 0, 4.0 * p.BF * p.C * ((p.dz2 ? 6.0 : 4.0) * p.T + p.Ty), st);
    cudaMemsetAsync(p.red, 0, sizeof(double) * (2 * p.C + 1), st);
    const int lanes = 256 / (p.C / 4);
    long long g = (((long long)p.BF * p.T + 1) / 2 + lanes - 1) / lanes;      // two rows per thread and iteration
    if (g > 148 * 4) g = 148 * 4;                                              // 4 resident CTAs per SM (64 registers): one wave
    bn_prelu_bwd_reduce_kernel<<<(int)g, 256, 0, st>>>(p);
    SEFD_TRY(sefd_check_launch("bn_prelu_bwd_reduce"));
    bn_prelu_bwd_apply_kernel<<<grid_for((long long)p.BF * p.Ty * (p.C / 4)), 256, 0, st>>>(p);
    return sefd_check_launch("bn_prelu_bwd_apply");
}

int sefd_pack_cconv(const CconvPackParams& p, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    pack_cconv_kernel<<<grid_for(40ll * p.Ci2 * p.Co2), 256, 0, st>>>(p);
    return sefd_check_launch("pack_cconv");
}

int sefd_fold_cconv(const CconvFoldParams& p, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    fold_cconv_kernel<<<grid_for(10ll * p.Ci2 * p.Co2), 256, 0, st>>>(p);
    return sefd_check_launch("fold_cconv");
}

int sefd_pack_rconv(const RconvPackParams& p, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    pack_rconv_kernel<<<grid_for(10ll * p.Ci * p.Co), 256, 0, st>>>(p);
    return sefd_check_launch("pack_rconv");
}

int sefd_fold_rconv(const RconvFoldParams& p, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    fold_rconv_kernel<<<grid_for(10ll * p.Ci * p.Co), 256, 0, st>>>(p);
    return sefd_check_launch("fold_rconv");
}

int sefd_permute3p(const Permute3Params& q, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    permute3_kernel<<<grid_for((long long)q.na * q.nb * q.nc), 256, 0, st>>>(q);
    return sefd_check_launch("permute3");
}

int sefd_permute3(const float* src, float* dst, int na, int nb, int nc, long long sa, long long sb, long long sc,
                  int accumulate, cudaStream_t st) {
    Permute3Params q;
    q.src = src; q.dst = dst; q.na = na; q.nb = nb; q.nc = nc; q.sa = sa; q.sb = sb; q.sc = sc;
    q.da = (long long)nb * nc; q.db = nc; q.dc = 1;
    q.accumulate = accumulate; q.nsplit = 1; q.split_stride = 0; q.round_tf32 = 0;
    return sefd_permute3p(q, st);
}

int sefd_add2(const float* a, const float* b, float* o, long long n, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    add2_kernel<<<grid_for(n), 256, 0, st>>>(a, b, o, n);
    return sefd_check_launch("add2");
}

int sefd_colsum2(const float* x, int nO, long long sO, long long nI, long long sI, int C, double* scratch, float* out,
                 cudaStream_t st) {
    SEFD_REQUIRE(C >= 1 && C <= 512, "colsum: C=%d unsupported", C);
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    cudaMemsetAsync(scratch, 0, sizeof(double) * C, st);
    const int block = C > 256 ? 512 : 256;
    const int lanes = block / C;
    long long g = ((long long)nO * nI + lanes - 1) / lanes;
    if (g > 148 * 4) g = 148 * 4;
    colsum2_kernel<<<(int)g, block, 0, st>>>(x, nO, sO, nI, sI, C, scratch);
    SEFD_TRY(sefd_check_launch("colsum2"));
    d2f_kernel<<<(C + 255) / 256, 256, 0, st>>>(scratch, out, C);
    return sefd_check_launch("d2f");
}

int sefd_clstm_combine(const float* H, float* X, long long n, int round_tf32, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    clstm_combine_kernel<<<grid_for(n), 256, 0, st>>>(H, X, n, round_tf32);
    return sefd_check_launch("clstm_combine");
}

int sefd_clstm_combine_bwd(const float* dX, float* dH, long long n, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    clstm_combine_bwd_kernel<<<grid_for(n), 256, 0, st>>>(dX, dH, n);
    return sefd_check_launch("clstm_combine_bwd");
}

int sefd_adam_dev(float* w, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                  int* step_dev, float* bc_dev, float gscale, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    adam_prep_kernel<<<1, 1, 0, st>>>(step_dev, bc_dev, (double)b1, (double)b2);
    SEFD_TRY(sefd_check_launch("adam_prep"));
    adam_dev_kernel<<<grid_for(n), 256, 0, st>>>(w, g, m, v, n, lr, b1, b2, eps, bc_dev, gscale);
    return sefd_check_launch("adam_dev");
}

int sefd_adam(float* w, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
              int step, float gscale, cudaStream_t st) {
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    adam_kernel<<<grid_for(n), 256, 0, st>>>(w, g, m, v, n, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), gscale);
    return sefd_check_launch("adam");
}

// ---- small step glue that used to run as eager framework kernels ------------------------------------------------------
namespace {
__global__ void axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = fmaf(a, x[i], b * y[i]);
}
__global__ void counters_inc_kernel(long long* const* __restrict__ ptrs, int n, long long inc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) *ptrs[i] += inc;
}
}  // namespace

int sefd_axpby_launch(float* y, const float* x, float a, float b, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    long long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    sefd_absorb_stale_error();
    axpby_kernel<<<(int)g, 256, 0, st>>>(y, x, a, b, n);
    return sefd_check_launch("axpby");
}

int sefd_counters_inc_launch(long long* const* ptrs, int n, long long inc, cudaStream_t st) {
    if (n <= 0) return 0;
    sefd_absorb_stale_error();
    counters_inc_kernel<<<(n + 127) / 128, 128, 0, st>>>(ptrs, n, inc);
    return sefd_check_launch("counters_inc");
}
