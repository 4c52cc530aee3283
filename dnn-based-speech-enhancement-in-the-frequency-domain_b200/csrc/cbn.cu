// ComplexBatchNorm + PReLU (use_cbn = True: models.py:76, 120, 151; tools_for_model.py:430-603), forward and backward.
//
// Per complex feature k (h = C / 2 of them; channels-last rows hold the real parts in [0, h) and the imaginary parts in
// [h, 2h)):  xc = x - M,  V = E[xc xc^T] + eps I,  U = V^{-1/2} (closed form of the 2x2 inverse square root,
// tools_for_model.py:563-574),  Z = W U with W = [[Wrr, Wri], [Wri, Wii]],  y = Z xc + B,  z = prelu(y).
// Train mode uses the batch moments and lerps them (WITHOUT eps, biased) into RMr / RMi / RVrr / RVri / RVii with
// momentum 0.1; eval mode reads the running buffers.  The reference's torch.addcmul(Vrr * Vii, -1, Vri, Vri) is
// Vrr Vii - Vri^2 (legacy positional `value`).
//
// Backward (hand-derived, checked against autograd of the oracle in double to 1e-15):
//   g = dz * prelu'(y);  dB = sum g;  dZ = sum g xc^T;  dW = dZ U (symmetric parts), dU = W dZ;
//   dU -> dV through rst = 1 / (s t), t = sqrt(tau + 2 s), s = sqrt(delta), tau = Vrr + Vii, delta = Vrr Vii - Vri^2;
//   dx = Z^T (g - dB / N) + [[2 dVrr, dVri], [dVri, 2 dVii]] xc / N.
// Same pass structure as the real BatchNorm (elementwise.cu): statistics pass, apply pass; backward reduce pass, apply pass
// over all Ty frames (frames outside the kept window carry g = 0 but receive the statistics terms).
#include "elementwise.cuh"
#include "prof.cuh"

namespace {

constexpr int MAXH = 256;

inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
    long long g = (n + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// block reduction of NM per-thread double[4] accumulators over the `lanes` row-lanes that share a feature quad, three sets at a
// time (24 KB of shared memory), then one atomicAdd per (set, feature) into red[set * h + feature]
template <int NM>
__device__ __forceinline__ void reduce_sets(double (&a)[NM][4], double* red, int h, int H4, int lanes) {
    __shared__ double s_red[3][256][4];
    for (int m0 = 0; m0 < NM; m0 += 3) {
#pragma unroll
        for (int m = 0; m < 3; ++m)
            if (m0 + m < NM) {
#pragma unroll
                for (int i = 0; i < 4; ++i) s_red[m][threadIdx.x][i] = a[m0 + m < NM ? m0 + m : 0][i];
            }
        __syncthreads();
        if (threadIdx.x < H4) {
            for (int m = 0; m < 3 && m0 + m < NM; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double t = 0.0;
                    for (int l = 0; l < lanes; ++l) t += s_red[m][l * H4 + threadIdx.x][i];
                    atomicAdd(red + (m0 + m) * h + threadIdx.x * 4 + i, t);
                }
        }
        __syncthreads();
    }
}

// moments of y: red[0..5h) = sum xr, sum xi, sum xr^2, sum xr xi, sum xi^2 (double; raw moments, centred at finalize)
__global__ void __launch_bounds__(256) cbn_stats_kernel(const float* __restrict__ y, long long rows, int C, double* __restrict__ red) {
    const int h = C >> 1, H4 = h >> 2;
    const int lanes = 256 / H4;
    const int q = threadIdx.x % H4, rl = threadIdx.x / H4;
    double a[5][4];
#pragma unroll
    for (int m = 0; m < 5; ++m)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[m][i] = 0.0;
    if (rl < lanes) {
        for (long long row = (long long)blockIdx.x * lanes + rl; row < rows; row += (long long)gridDim.x * lanes) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(y + row * C + 4 * q));
            const float4 i4 = __ldg(reinterpret_cast<const float4*>(y + row * C + h + 4 * q));
            const float xr[4] = {r4.x, r4.y, r4.z, r4.w}, xi[4] = {i4.x, i4.y, i4.z, i4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[0][i] += (double)xr[i];
                a[1][i] += (double)xi[i];
                a[2][i] += (double)xr[i] * (double)xr[i];
                a[3][i] += (double)xr[i] * (double)xi[i];
                a[4][i] += (double)xi[i] * (double)xi[i];
            }
        }
    }
    reduce_sets<5>(a, red, h, H4, lanes);
}

// per feature: moments -> M, V (+ running buffers), U, Z; save[9][h] = Mr, Mi, Zrr, Zri, Zir, Zii, Vrr + eps, Vri, Vii + eps
__global__ void cbn_finalize_kernel(const CbnPreluFwdParams p) {
    const int h = p.C >> 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= h) return;
    double Mr, Mi, Vrr, Vri, Vii;
    if (p.use_running) {
        Mr = p.RM[0][k]; Mi = p.RM[1][k];
        Vrr = p.RV[0][k]; Vri = p.RV[1][k]; Vii = p.RV[2][k];
    } else {
        const double n = p.n_stat;
        Mr = p.stats[k] / n; Mi = p.stats[h + k] / n;
        Vrr = p.stats[2 * h + k] / n - Mr * Mr;
        Vri = p.stats[3 * h + k] / n - Mr * Mi;
        Vii = p.stats[4 * h + k] / n - Mi * Mi;
        if (Vrr < 0) Vrr = 0;
        if (Vii < 0) Vii = 0;
        if (p.RM[0]) {                                     // lerp_(batch value, momentum), tools_for_model.py:527-550
            const float f = p.momentum;
            p.RM[0][k] += f * ((float)Mr - p.RM[0][k]);
            p.RM[1][k] += f * ((float)Mi - p.RM[1][k]);
            p.RV[0][k] += f * ((float)Vrr - p.RV[0][k]);
            p.RV[1][k] += f * ((float)Vri - p.RV[1][k]);
            p.RV[2][k] += f * ((float)Vii - p.RV[2][k]);
        }
    }
    Vrr += (double)p.eps;
    Vii += (double)p.eps;
    const double tau = Vrr + Vii, delta = Vrr * Vii - Vri * Vri;
    const double s = sqrt(delta), t = sqrt(tau + 2.0 * s), rst = 1.0 / (s * t);
    const double Urr = (s + Vii) * rst, Uii = (s + Vrr) * rst, Uri = -Vri * rst;
    const double Wrr = p.W[0][k], Wri = p.W[1][k], Wii = p.W[2][k];
    p.save[0 * h + k] = (float)Mr;
    p.save[1 * h + k] = (float)Mi;
    p.save[2 * h + k] = (float)(Wrr * Urr + Wri * Uri);
    p.save[3 * h + k] = (float)(Wrr * Uri + Wri * Uii);
    p.save[4 * h + k] = (float)(Wri * Urr + Wii * Uri);
    p.save[5 * h + k] = (float)(Wri * Uri + Wii * Uii);
    p.save[6 * h + k] = (float)Vrr;
    p.save[7 * h + k] = (float)Vri;
    p.save[8 * h + k] = (float)Vii;
}

struct Quad {
    float mr[4], mi[4], zrr[4], zri[4], zir[4], zii[4], br[4], bi[4];
};
__device__ __forceinline__ void load_quad(Quad& c, const float* save, const float* Br, const float* Bi, int h, int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c.mr[i] = save[0 * h + k0 + i]; c.mi[i] = save[1 * h + k0 + i];
        c.zrr[i] = save[2 * h + k0 + i]; c.zri[i] = save[3 * h + k0 + i];
        c.zir[i] = save[4 * h + k0 + i]; c.zii[i] = save[5 * h + k0 + i];
        c.br[i] = Br[k0 + i]; c.bi[i] = Bi[k0 + i];
    }
}

// z[bf, t] = prelu(Z (y[bf, t + tshift] - M) + B); a thread owns one feature quad (real + imaginary float4) of one row
__global__ void __launch_bounds__(256) cbn_prelu_fwd_kernel(const CbnPreluFwdParams p) {
    const int C = p.C, h = C >> 1, H4 = h >> 2;
    const float alpha = p.alpha[0];
    const long long total = (long long)p.BF * p.T * H4;
    // the grid stride is a multiple of H4 (H4 divides 256): a thread keeps its feature quad, constants live in registers
    const int k0 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % H4) * 4;
    Quad c;
    load_quad(c, p.save, p.B2[0], p.B2[1], h, k0);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / H4;
        const int t = (int)(row % p.T);
        const long long bf = row / p.T;
        const float* src = p.y + (bf * p.Ty + t + p.tshift) * C + k0;
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(src)), i4 = __ldg(reinterpret_cast<const float4*>(src + h));
        const float xr[4] = {r4.x, r4.y, r4.z, r4.w}, xi[4] = {i4.x, i4.y, i4.z, i4.w};
        float orr[4], oi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = xr[i] - c.mr[i], b = xi[i] - c.mi[i];
            float yr = fmaf(c.zrr[i], a, fmaf(c.zri[i], b, c.br[i]));
            float yi = fmaf(c.zir[i], a, fmaf(c.zii[i], b, c.bi[i]));
            yr = yr > 0.f ? yr : alpha * yr;
            yi = yi > 0.f ? yi : alpha * yi;
            if (p.round_tf32) { yr = tf32_rn(yr); yi = tf32_rn(yi); }
            orr[i] = yr; oi[i] = yi;
        }
        float* dst = p.z + row * C + k0;
        *reinterpret_cast<float4*>(dst) = make_float4(orr[0], orr[1], orr[2], orr[3]);
        *reinterpret_cast<float4*>(dst + h) = make_float4(oi[0], oi[1], oi[2], oi[3]);
    }
}

// backward, pass 1: red[0..6h) = sum gr, sum gi, sum gr xcr, sum gr xci, sum gi xcr, sum gi xci; red[6h] = d alpha
__global__ void __launch_bounds__(256) cbn_bwd_reduce_kernel(const CbnPreluBwdParams p) {
    const int C = p.C, h = C >> 1, H4 = h >> 2;
    const int lanes = 256 / H4;
    const int q = threadIdx.x % H4, rl = threadIdx.x / H4, k0 = 4 * q;
    const float alpha = p.alpha[0];
    Quad c;
    load_quad(c, p.save, p.B2[0], p.B2[1], h, k0);
    double a[6][4], sa = 0.0;
#pragma unroll
    for (int m = 0; m < 6; ++m)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[m][i] = 0.0;
    const long long rows = (long long)p.BF * p.T;
    if (rl < lanes) {
        for (long long row = (long long)blockIdx.x * lanes + rl; row < rows; row += (long long)gridDim.x * lanes) {
            const int t = (int)(row % p.T);
            const long long bf = row / p.T;
            const float* src = p.y + (bf * p.Ty + t + p.tshift) * C + k0;
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(src)), i4 = __ldg(reinterpret_cast<const float4*>(src + h));
            float4 dr = __ldg(reinterpret_cast<const float4*>(p.dz + row * C + k0));
            float4 di = __ldg(reinterpret_cast<const float4*>(p.dz + row * C + h + k0));
            if (p.dz2) {
                const float4 er = __ldg(reinterpret_cast<const float4*>(p.dz2 + row * C + k0));
                const float4 ei = __ldg(reinterpret_cast<const float4*>(p.dz2 + row * C + h + k0));
                dr.x += er.x; dr.y += er.y; dr.z += er.z; dr.w += er.w;
                di.x += ei.x; di.y += ei.y; di.z += ei.z; di.w += ei.w;
            }
            const float xr[4] = {r4.x, r4.y, r4.z, r4.w}, xi[4] = {i4.x, i4.y, i4.z, i4.w};
            const float d_r[4] = {dr.x, dr.y, dr.z, dr.w}, d_i[4] = {di.x, di.y, di.z, di.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xa = xr[i] - c.mr[i], xb = xi[i] - c.mi[i];
                const float yr = fmaf(c.zrr[i], xa, fmaf(c.zri[i], xb, c.br[i]));
                const float yi = fmaf(c.zir[i], xa, fmaf(c.zii[i], xb, c.bi[i]));
                const float gr = yr > 0.f ? d_r[i] : alpha * d_r[i];
                const float gi = yi > 0.f ? d_i[i] : alpha * d_i[i];
                a[0][i] += (double)gr;
                a[1][i] += (double)gi;
                a[2][i] += (double)(gr * xa);
                a[3][i] += (double)(gr * xb);
                a[4][i] += (double)(gi * xa);
                a[5][i] += (double)(gi * xb);
                sa += (yr > 0.f ? 0.0 : (double)(d_r[i] * yr)) + (yi > 0.f ? 0.0 : (double)(d_i[i] * yi));
            }
        }
    }
    reduce_sets<6>(a, p.red, h, H4, lanes);
    // d alpha: one scalar per block (warp shuffle, then the first lanes of the warps through shared memory)
    __shared__ double s_a[8];
    for (int o = 16; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
    if ((threadIdx.x & 31) == 0) s_a[threadIdx.x >> 5] = sa;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_a[w];
        atomicAdd(p.red + 6 * h, t);
    }
}

// per feature: parameter gradients and the coefficients of pass 2; coef[9][h] = Zrr, Zri, Zir, Zii (dx = Z^T g ...),
// cr, ci (= Z^T dB / N), qrr, qri, qii (= [[2 dVrr, dVri], [dVri, 2 dVii]] / N)
__global__ void cbn_bwd_finalize_kernel(const CbnPreluBwdParams p) {
    const int h = p.C >> 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) p.dalpha[0] = (float)p.red[6 * h];
    if (k >= h) return;
    const double N = p.n_stat;
    const double dBr = p.red[k], dBi = p.red[h + k];
    const double dZrr = p.red[2 * h + k], dZri = p.red[3 * h + k], dZir = p.red[4 * h + k], dZii = p.red[5 * h + k];
    const double Vrr = p.save[6 * h + k], Vri = p.save[7 * h + k], Vii = p.save[8 * h + k];     // eps included
    const double Wrr = p.W[0][k], Wri = p.W[1][k], Wii = p.W[2][k];
    const double tau = Vrr + Vii, delta = Vrr * Vii - Vri * Vri;
    const double s = sqrt(delta), t = sqrt(tau + 2.0 * s), rst = 1.0 / (s * t);
    const double Urr = (s + Vii) * rst, Uii = (s + Vrr) * rst, Uri = -Vri * rst;
    p.dW[0][k] = (float)(dZrr * Urr + dZri * Uri);
    p.dW[1][k] = (float)(dZrr * Uri + dZri * Uii + dZir * Urr + dZii * Uri);
    p.dW[2][k] = (float)(dZir * Uri + dZii * Uii);
    p.dB2[0][k] = (float)dBr;
    p.dB2[1][k] = (float)dBi;
    const double dUrr = dZrr * Wrr + dZir * Wri;
    const double dUri = dZrr * Wri + dZri * Wrr + dZir * Wii + dZii * Wri;
    const double dUii = dZri * Wri + dZii * Wii;
    const double d_rst = dUrr * (s + Vii) + dUii * (s + Vrr) - dUri * Vri;
    double d_s = (dUrr + dUii) * rst - d_rst * rst / s;
    const double d_t = -d_rst * rst / t;
    const double d_tau = d_t / (2.0 * t);
    d_s += d_t / t;
    const double d_delta = d_s / (2.0 * s);
    const double dVrr = dUii * rst + d_tau + d_delta * Vii;
    const double dVii = dUrr * rst + d_tau + d_delta * Vrr;
    const double dVri = -dUri * rst - 2.0 * d_delta * Vri;
    const double Zrr = Wrr * Urr + Wri * Uri, Zri = Wrr * Uri + Wri * Uii, Zir = Wri * Urr + Wii * Uri, Zii = Wri * Uri + Wii * Uii;
    p.coef[0 * h + k] = (float)Zrr; p.coef[1 * h + k] = (float)Zri; p.coef[2 * h + k] = (float)Zir; p.coef[3 * h + k] = (float)Zii;
    p.coef[4 * h + k] = (float)((Zrr * dBr + Zir * dBi) / N);
    p.coef[5 * h + k] = (float)((Zri * dBr + Zii * dBi) / N);
    p.coef[6 * h + k] = (float)(2.0 * dVrr / N);
    p.coef[7 * h + k] = (float)(dVri / N);
    p.coef[8 * h + k] = (float)(2.0 * dVii / N);
}

// backward, pass 2: dy = Z^T g - c + Q xc over all Ty frames
__global__ void __launch_bounds__(256) cbn_bwd_apply_kernel(const CbnPreluBwdParams p) {
    const int C = p.C, h = C >> 1, H4 = h >> 2;
    const float alpha = p.alpha[0];
    const long long total = (long long)p.BF * p.Ty * H4;
    const int k0 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % H4) * 4;
    Quad c;
    load_quad(c, p.save, p.B2[0], p.B2[1], h, k0);
    float cr[4], ci[4], qrr[4], qri[4], qii[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        cr[i] = p.coef[4 * h + k0 + i]; ci[i] = p.coef[5 * h + k0 + i];
        qrr[i] = p.coef[6 * h + k0 + i]; qri[i] = p.coef[7 * h + k0 + i]; qii[i] = p.coef[8 * h + k0 + i];
    }
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / H4;
        const int ty = (int)(row % p.Ty);
        const long long bf = row / p.Ty;
        const int t = ty - p.tshift;
        const float* src = p.y + row * C + k0;
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(src)), i4 = __ldg(reinterpret_cast<const float4*>(src + h));
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f), di = dr;
        if (t >= 0 && t < p.T) {
            const long long o = (bf * p.T + t) * C + k0;
            dr = __ldg(reinterpret_cast<const float4*>(p.dz + o));
            di = __ldg(reinterpret_cast<const float4*>(p.dz + o + h));
            if (p.dz2) {
                const float4 er = __ldg(reinterpret_cast<const float4*>(p.dz2 + o)), ei = __ldg(reinterpret_cast<const float4*>(p.dz2 + o + h));
                dr.x += er.x; dr.y += er.y; dr.z += er.z; dr.w += er.w;
                di.x += ei.x; di.y += ei.y; di.z += ei.z; di.w += ei.w;
            }
        }
        const float xr[4] = {r4.x, r4.y, r4.z, r4.w}, xi[4] = {i4.x, i4.y, i4.z, i4.w};
        const float d_r[4] = {dr.x, dr.y, dr.z, dr.w}, d_i[4] = {di.x, di.y, di.z, di.w};
        float orr[4], oi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float xa = xr[i] - c.mr[i], xb = xi[i] - c.mi[i];
            const float yr = fmaf(c.zrr[i], xa, fmaf(c.zri[i], xb, c.br[i]));
            const float yi = fmaf(c.zir[i], xa, fmaf(c.zii[i], xb, c.bi[i]));
            const float gr = yr > 0.f ? d_r[i] : alpha * d_r[i];
            const float gi = yi > 0.f ? d_i[i] : alpha * d_i[i];
            float a = fmaf(c.zrr[i], gr, c.zir[i] * gi) - cr[i] + fmaf(qrr[i], xa, qri[i] * xb);
            float b = fmaf(c.zri[i], gr, c.zii[i] * gi) - ci[i] + fmaf(qri[i], xa, qii[i] * xb);
            if (p.round_tf32) { a = tf32_rn(a); b = tf32_rn(b); }
            orr[i] = a; oi[i] = b;
        }
        float* dst = p.dy + row * C + k0;
        *reinterpret_cast<float4*>(dst) = make_float4(orr[0], orr[1], orr[2], orr[3]);
        *reinterpret_cast<float4*>(dst + h) = make_float4(oi[0], oi[1], oi[2], oi[3]);
    }
}

}  // namespace

int sefd_cbn_prelu_fwd(const CbnPreluFwdParams& p, cudaStream_t st) {
    const int h = p.C / 2;
    SEFD_REQUIRE(p.C % 8 == 0 && h <= MAXH && 256 % (h / 4) == 0, "cbn_prelu_fwd: C=%d unsupported", p.C);
    sefd_prof_label("cbn_prelu_fwd C%d rows%lld", p.C, (long long)p.BF * p.T);
    SefdProfScope prof(SEFD_PROF_BN, 0, 4.0 * p.BF * p.C * ((p.use_running ? 0.0 : 1.0) * p.Ty + 2.0 * p.T), st);
    if (!p.use_running) {
        SEFD_REQUIRE(p.stats != nullptr, "cbn_prelu_fwd: train mode needs the moment scratch%s", "");
        cudaMemsetAsync(p.stats, 0, sizeof(double) * 5 * h, st);
        const long long rows = (long long)p.BF * p.Ty;
        const int lanes = 256 / (h / 4);
        long long g = (rows + lanes - 1) / lanes;
        if (g > 148 * 8) g = 148 * 8;
        cbn_stats_kernel<<<(int)g, 256, 0, st>>>(p.y, rows, p.C, p.stats);
        SEFD_TRY(sefd_check_launch("cbn_stats"));
    } else {
        SEFD_REQUIRE(p.RM[0] && p.RM[1] && p.RV[0] && p.RV[1] && p.RV[2], "cbn_prelu_fwd: eval mode needs the running buffers%s", "");
    }
    cbn_finalize_kernel<<<(h + 63) / 64, 64, 0, st>>>(p);
    SEFD_TRY(sefd_check_launch("cbn_finalize"));
    cbn_prelu_fwd_kernel<<<grid_for((long long)p.BF * p.T * (h / 4)), 256, 0, st>>>(p);
    return sefd_check_launch("cbn_prelu_fwd");
}

int sefd_cbn_prelu_bwd(const CbnPreluBwdParams& p, cudaStream_t st) {
    const int h = p.C / 2;
    SEFD_REQUIRE(p.C % 8 == 0 && h <= MAXH && 256 % (h / 4) == 0, "cbn_prelu_bwd: C=%d unsupported", p.C);
    sefd_prof_label("cbn_prelu_bwd C%d rows%lld", p.C, (long long)p.BF * p.T);
    SefdProfScope prof(SEFD_PROF_BN, 0, 4.0 * p.BF * p.C * ((p.dz2 ? 6.0 : 4.0) * p.T + p.Ty), st);
    cudaMemsetAsync(p.red, 0, sizeof(double) * (6 * h + 1), st);
    const int lanes = 256 / (h / 4);
    long long g = ((long long)p.BF * p.T + lanes - 1) / lanes;
    if (g > 148 * 8) g = 148 * 8;
    cbn_bwd_reduce_kernel<<<(int)g, 256, 0, st>>>(p);
    SEFD_TRY(sefd_check_launch("cbn_bwd_reduce"));
    cbn_bwd_finalize_kernel<<<(h + 63) / 64, 64, 0, st>>>(p);
    SEFD_TRY(sefd_check_launch("cbn_bwd_finalize"));
    cbn_bwd_apply_kernel<<<grid_for((long long)p.BF * p.Ty * (h / 4)), 256, 0, st>>>(p);
    return sefd_check_launch("cbn_bwd_apply");
}
