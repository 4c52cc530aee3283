// FullSubNet train-step orchestration (BASELINE configs[2]): FullSubNet.forward (models.py:626-672),
// SequenceModel (tools_for_model.py:726-795), BaseModel.unfold (:806-837), offline_laplace_norm (:997-1011) and the autograd
// graph torch derives from them.
//
// Data layout: every sequence tensor is TIME-MAJOR [T][rows][C] (lstm_seq.cuh); the full-band model runs B sequences
// (rows = B, LSTM 257 -> 512 -> 512, Linear 512 -> 257, ReLU), the sub-band model B * 257 sequences (rows = R,
// LSTM 32 -> 384 -> 384, Linear 384 -> 2).  T = noisy frames + 2 look-ahead zero frames (models.py:640).  The 257-wide
// full-band tensors are stored 288 wide (zero padding) so that they are tensor-core GEMM operands.
//   fb_in  [T][B][288]   = noisy_mag / (mean_b + 1e-5)                                  (models.py:645)
//   fb_lin [T][B][288]   = fc_output_layer(h1)  (ReLU is applied by its consumers)      (tools_for_model.py:787-789)
//   sb_in  [T][R][32]    = [31 reflect-unfolded noisy bins | fb_out] / (mean_b + 1e-5)  (models.py:649-658)
//   crm    [B][257][Tf][2] = sub-band mask without the look-ahead frames                (models.py:667-671)
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "fsnet.cuh"
#include "lstm_seq.cuh"
#include "plan.cuh"
#include "prof.cuh"
#include "seqstack.cuh"

namespace {

constexpr int FBINS = 257, FPAD = 288, SB_N = 15, SB_I = 32, LOOK = 2, FB_H = 512, SB_H = 384;
constexpr float NORM_EPS = 1e-5f;

}  // namespace

struct FsnExt {
    int B, Tf, T, R;
    SeqStack fb, sb;
    size_t rowsum /*double [B][257]*/, wsum /*double [B]*/, sum2 /*double [B]*/, Sred /*double [B]*/, inv /*float [2][B]*/;
    size_t fb_in, fb_lin, dfb_lin, Wl_nk, Wl_kn, bl, sb_in, dsb_in;
    size_t hpart, red /*double*/;
    mutable SeqScratch sc;      // scratch of the LSTM stacks + dropout state of the last forward (needed by the backward)
    int head_blocks;
    mutable const float *mask_fb = nullptr, *mask_sb = nullptr;
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// feature normalisation / sub-band unfolding
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_bin(int i) { return i < 0 ? -i : (i >= FBINS ? 2 * (FBINS - 1) - i : i); }

// rowsum[b][f] = sum_t x[b][f][t]; one warp per row
__global__ void fsn_rowsum_kernel(const float* __restrict__ x, int rows, int Tf, double* __restrict__ out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= rows) return;
    double s = 0.0;
    for (int t = lane; t < Tf; t += 32) s += (double)x[(long long)w * Tf + t];
    s = warp_sum_d(s);
    if (lane == 0) out[w] = s;
}

// per utterance: inv[0][b] = 1 / (mean over [257][T] + eps) (models.py:645) and wsum[b] = sum of the reflect-unfolded noisy
// part of the sub-band input = sum_f' cnt[f'] rowsum[b][f'] (cnt = how often bin f' appears in the 31-neighbour unfold)
__global__ void fsn_mu_kernel(const double* __restrict__ rowsum, int T, float* __restrict__ inv1, double* __restrict__ wsum) {
    __shared__ int cnt[FBINS];
    __shared__ double red[2][8];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < FBINS; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < FBINS * (2 * SB_N + 1); e += blockDim.x) {
        const int f = e / (2 * SB_N + 1), j = e % (2 * SB_N + 1);
        atomicAdd(&cnt[reflect_bin(f + j - SB_N)], 1);
    }
    __syncthreads();
    double s = 0.0, w = 0.0;
    for (int f = threadIdx.x; f < FBINS; f += blockDim.x) {
        const double r = rowsum[b * FBINS + f];
        s += r;
        w += r * cnt[f];
    }
    s = warp_sum_d(s);
    w = warp_sum_d(w);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = w; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; c += red[1][i]; }
        const float mu = (float)(a / ((double)FBINS * T));
        inv1[b] = 1.f / (mu + NORM_EPS);
        wsum[b] = c;
    }
}

// fb_in[t][b][f] = x[b][f][t] * inv1[b] (zero in the look-ahead frames and in the padding columns): tiled transpose
__global__ void fsn_fb_in_kernel(const float* __restrict__ x, const float* __restrict__ inv1, int B, int Tf, int T, float* __restrict__ out,
                                 int round_tf32) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int f = f0 + r, t = t0 + threadIdx.x;
        tile[r][threadIdx.x] = (f < FBINS && t < Tf) ? x[((long long)b * FBINS + f) * Tf + t] : 0.f;
    }
    __syncthreads();
    const float s = inv1[b];
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int t = t0 + r, f = f0 + threadIdx.x;
        if (t < T && f < FPAD) {
            float v = tile[threadIdx.x][r] * s;
            if (round_tf32) v = tf32_rn(v);
            out[((long long)t * B + b) * FPAD + f] = v;
        }
    }
}

// sum2[b] += sum_f relu(lin[t][b][f])  (grid = T x B)
__global__ void fsn_relusum_kernel(const float* __restrict__ lin, int B, double* __restrict__ sum2) {
    __shared__ double red[8];
    const int t = blockIdx.x, b = blockIdx.y;
    const float* p = lin + ((long long)t * B + b) * FPAD;
    double s = 0.0;
    for (int f = threadIdx.x; f < FBINS; f += blockDim.x) s += (double)fmaxf(p[f], 0.f);
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += red[i];
        atomicAdd(sum2 + b, a);
    }
}

// sb_in[t][b*257+f][j] = (j < 31 ? x[b][reflect(f + j - 15)][t] : relu(lin[t][b][f])) * inv2[b]   (models.py:649-658)
// one CTA = (16 frames, utterance b): the 257 x 16 spectrogram tile is staged in shared memory (coalesced over t), one warp
// writes one 128-byte sub-band row at a time
constexpr int UT = 16;
__global__ void __launch_bounds__(256) fsn_unfold_kernel(const float* __restrict__ x, const float* __restrict__ lin,
                                                         const double* __restrict__ wsum, const double* __restrict__ sum2, int B,
                                                         int Tf, int T, float* __restrict__ inv2, float* __restrict__ out,
                                                         int round_tf32) {
    __shared__ float tile[FBINS][UT + 1];
    const int b = blockIdx.y, t0 = blockIdx.x * UT;
    for (int e = threadIdx.x; e < FBINS * UT; e += blockDim.x) {
        const int f = e / UT, tt = e % UT, t = t0 + tt;
        tile[f][tt] = t < Tf ? x[((long long)b * FBINS + f) * Tf + t] : 0.f;
    }
    const float mu = (float)((wsum[b] + sum2[b]) / ((double)FBINS * SB_I * T));
    const float s = 1.f / (mu + NORM_EPS);
    if (blockIdx.x == 0 && threadIdx.x == 0) inv2[b] = s;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = warp; e < FBINS * UT; e += 8) {
        const int tt = e / FBINS, f = e % FBINS, t = t0 + tt;
        if (t >= T) break;
        float v;
        if (lane < 2 * SB_N + 1) v = tile[reflect_bin(f + lane - SB_N)][tt];
        else v = fmaxf(lin[((long long)t * B + b) * FPAD + f], 0.f);
        v *= s;
        if (round_tf32) v = tf32_rn(v);
        out[((long long)t * B * FBINS + (long long)b * FBINS + f) * SB_I + lane] = v;
    }
}

// S[b] += sum over the (t, b) slab [257][32] of dsb_in * sb_in   (grid = T x B)
__global__ void fsn_unfold_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ y, int B, double* __restrict__ S) {
    __shared__ double red[8];
    const int t = blockIdx.x, b = blockIdx.y;
    const long long base = ((long long)t * B + b) * FBINS * SB_I;
    const float4* a = reinterpret_cast<const float4*>(dy + base);
    const float4* c = reinterpret_cast<const float4*>(y + base);
    double s = 0.0;
    for (int e = threadIdx.x; e < FBINS * SB_I / 4; e += blockDim.x) {
        const float4 u = a[e], v = c[e];
        s += (double)(u.x * v.x + u.y * v.y) + (double)(u.z * v.z + u.w * v.w);
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
        atomicAdd(S + b, r);
    }
}

// d lin[t][b][f] = relu'(lin) * (dsb_in[t][b*257+f][31] - S[b] / N) * inv2[b];   padding columns -> 0
__global__ void fsn_unfold_bwd_kernel(const float* __restrict__ dsb, const float* __restrict__ lin, const double* __restrict__ S,
                                      const float* __restrict__ inv2, int B, int T, float* __restrict__ dlin, int round_tf32) {
    const long long n = (long long)T * B * FPAD;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(e % FPAD);
        const long long tb = e / FPAD;
        const int b = (int)(tb % B);
        float v = 0.f;
        if (f < FBINS && lin[e] > 0.f) {
            const float m = (float)(S[b] / ((double)FBINS * SB_I * T));
            v = (dsb[(tb * FBINS + f) * SB_I + (SB_I - 1)] - m) * inv2[b];
            if (round_tf32) v = tf32_rn(v);
        }
        dlin[e] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// sub-band head: Linear(384 -> 2) on every (t, r) row, look-ahead crop and the [B, F, T, 2] output permutation
// (tools_for_model.py:787, models.py:667-671): crm[r][t - 2][c] = W[c] . h1[t][r] + b[c]
// ---------------------------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256) sb_head_fwd_kernel(const float* __restrict__ h1, const float* __restrict__ W, const float* __restrict__ bias,
                                                          float* __restrict__ crm, int T, int R, int Tf) {
    constexpr int V = H / 128;          // float4 per lane
    const int lane = threadIdx.x & 31;
    float4 w0[V], w1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        w0[i] = reinterpret_cast<const float4*>(W)[i * 32 + lane];
        w1[i] = reinterpret_cast<const float4*>(W + H)[i * 32 + lane];
    }
    const float b0 = bias[0], b1 = bias[1];
    const long long rows = (long long)(T - LOOK) * R, wstep = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstep) {
        const long long t = row / R + LOOK, r = row % R;
        const float4* hp = reinterpret_cast<const float4*>(h1 + (t * R + r) * H);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float4 h = hp[i * 32 + lane];
            s0 += h.x * w0[i].x + h.y * w0[i].y + h.z * w0[i].z + h.w * w0[i].w;
            s1 += h.x * w1[i].x + h.y * w1[i].y + h.z * w1[i].z + h.w * w1[i].w;
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) *reinterpret_cast<float2*>(crm + (r * Tf + (t - LOOK)) * 2) = make_float2(s0 + b0, s1 + b1);
    }
}

// dh1[t][r][k] = d0 W[0][k] + d1 W[1][k] (zero in the look-ahead frames), partial sums of dW[c][k] = sum d_c h1[k], db[c]
template <int H>
__global__ void __launch_bounds__(256) sb_head_bwd_kernel(const float* __restrict__ h1, const float* __restrict__ W, const float* __restrict__ dcrm,
                                                          float* __restrict__ dh1, float* __restrict__ part, int T, int R, int Tf) {
    constexpr int V = H / 128;
    __shared__ __align__(16) float red[8][2 * H + 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 w0[V], w1[V], a0[V], a1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        w0[i] = reinterpret_cast<const float4*>(W)[i * 32 + lane];
        w1[i] = reinterpret_cast<const float4*>(W + H)[i * 32 + lane];
        a0[i] = a1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float sb0 = 0.f, sb1 = 0.f;
    const long long rows = (long long)T * R, wstep = (long long)gridDim.x * 8;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += wstep) {
        const long long t = row / R, r = row % R;
        float4* dp = reinterpret_cast<float4*>(dh1 + row * H);
        if (t < LOOK) {
#pragma unroll
            for (int i = 0; i < V; ++i) dp[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const float2 d = *reinterpret_cast<const float2*>(dcrm + (r * Tf + (t - LOOK)) * 2);
        const float4* hp = reinterpret_cast<const float4*>(h1 + row * H);
        sb0 += d.x;
        sb1 += d.y;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float4 h = hp[i * 32 + lane];
            a0[i].x = fmaf(d.x, h.x, a0[i].x); a0[i].y = fmaf(d.x, h.y, a0[i].y); a0[i].z = fmaf(d.x, h.z, a0[i].z); a0[i].w = fmaf(d.x, h.w, a0[i].w);
            a1[i].x = fmaf(d.y, h.x, a1[i].x); a1[i].y = fmaf(d.y, h.y, a1[i].y); a1[i].z = fmaf(d.y, h.z, a1[i].z); a1[i].w = fmaf(d.y, h.w, a1[i].w);
            dp[i * 32 + lane] = make_float4(d.x * w0[i].x + d.y * w1[i].x, d.x * w0[i].y + d.y * w1[i].y,
                                            d.x * w0[i].z + d.y * w1[i].z, d.x * w0[i].w + d.y * w1[i].w);
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
        reinterpret_cast<float4*>(&red[warp][0])[i * 32 + lane] = a0[i];
        reinterpret_cast<float4*>(&red[warp][H])[i * 32 + lane] = a1[i];
    }
    if (lane == 0) { red[warp][2 * H] = sb0; red[warp][2 * H + 1] = sb1; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * H + 2; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][i];
        part[(long long)blockIdx.x * (2 * H + 2) + i] = s;
    }
}

__global__ void sb_head_fold_kernel(const float* __restrict__ part, int nblk, int H, float* dW, float* db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * H + 2) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += (double)part[(long long)b * (2 * H + 2) + i];
    if (i < 2 * H) dW[i] = (float)s;
    else db[i - 2 * H] = (float)s;
}

// Linear weights [N_real][K] -> GEMM operands [Npad][K] (zero rows) and [K][Npad], bias [Npad]
__global__ void pack_linear_kernel(const float* __restrict__ w, const float* __restrict__ b, int N_real, int N, int K, float* nk, float* kn,
                                   float* bias, int round_tf32) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)N * K + N; e += (long long)gridDim.x * blockDim.x) {
        if (e < (long long)N * K) {
            const int n = (int)(e / K), k = (int)(e % K);
            float v = n < N_real ? w[(long long)n * K + k] : 0.f;
            if (round_tf32) v = tf32_rn(v);
            nk[e] = v;
            kn[(long long)k * N + n] = v;
        } else {
            const int n = (int)(e - (long long)N * K);
            bias[n] = n < N_real ? b[n] : 0.f;
        }
    }
}
// dW[n][k] = sum_s part[s][k][n] for n < N_real
__global__ void fold_linear_kernel(const float* __restrict__ part, int nsplit, long long stride, int K, int N, int N_real, float* __restrict__ dW) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, n = n0 + threadIdx.x;
        float s = 0.f;
        if (k < K && n < N)
            for (int sp = 0; sp < nsplit; ++sp) s += part[sp * stride + (long long)k * N + n];
        tile[r][threadIdx.x] = s;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int n = n0 + r, k = k0 + threadIdx.x;
        if (n < N_real && k < K) dW[(long long)n * K + k] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------------------------------
void add_stack_params(sefd_plan* P, SeqStack& S, const char* name, int I, int O, int H, long long& pc) {
    const std::string pre = std::string(name) + ".sequence_model.";
    for (int l = 0; l < 2; ++l) {
        SeqLayer& L = S.l[l];
        L.I_real = l == 0 ? I : H;
        L.I = (L.I_real + 31) / 32 * 32;
        L.H = H;
        const std::string sfx = "_l" + std::to_string(l);
        add_param(P, pre + "weight_ih" + sfx, pc, &L.w_ih, {4ll * H, L.I_real});
        add_param(P, pre + "weight_hh" + sfx, pc, &L.w_hh, {4ll * H, H});
        add_param(P, pre + "bias_ih" + sfx, pc, &L.b_ih, {4ll * H});
        add_param(P, pre + "bias_hh" + sfx, pc, &L.b_hh, {4ll * H});
    }
    add_param(P, std::string(name) + ".fc_output_layer.weight", pc, &S.fc_w, {O, H});
    add_param(P, std::string(name) + ".fc_output_layer.bias", pc, &S.fc_b, {O});
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
sefd_plan* sefd_fsn_plan_create_impl(int B, int Tf) {
    if (B <= 0 || Tf <= 0) {
        sefd_set_error("fsn plan: need B > 0 and frames > 0 (got B=%d frames=%d)", B, Tf);
        return nullptr;
    }
    sefd_plan* P = new sefd_plan();
    P->kind = 2;
    P->B = B; P->L = 0; P->T = Tf + LOOK; P->mask_mode = 0;
    FsnExt* E = new FsnExt();
    P->fsn = E;
    E->B = B; E->Tf = Tf; E->T = Tf + LOOK; E->R = B * FBINS;
    long long pc = 0;
    add_stack_params(P, E->fb, "fb_model", FBINS, FBINS, FB_H, pc);
    add_stack_params(P, E->sb, "sb_model", SB_I, 2, SB_H, pc);
    P->n_param_floats = pc;
    P->n_buffer_floats = 0;

    const int T = E->T, R = E->R;
    Carver w;
    E->rowsum = w.doubles((size_t)B * FBINS);
    E->wsum = w.doubles(B);
    E->sum2 = w.doubles(B);
    E->Sred = w.doubles(B);
    E->red = w.doubles(1024);
    E->inv = w.floats(2 * (size_t)B);
    E->fb_in = w.floats((size_t)T * B * FPAD);
    E->fb_lin = w.floats((size_t)T * B * FPAD);
    E->dfb_lin = w.floats((size_t)T * B * FPAD);
    E->Wl_nk = w.floats((size_t)FPAD * FB_H);
    E->Wl_kn = w.floats((size_t)FPAD * FB_H);
    E->bl = w.floats(FPAD);
    E->sb_in = w.floats((size_t)T * R * SB_I);
    E->dsb_in = w.floats((size_t)T * R * SB_I);
    carve_stack(E->fb, w, B, T);
    carve_stack(E->sb, w, R, T);
    const size_t state = std::max((size_t)(R + 127) / 128 * 128 * SB_H, (size_t)(B + 127) / 128 * 128 * FB_H);
    carve_seq_scratch(E->sc, w, state, R, FB_H, FB_H);
    E->head_blocks = 148 * 4;
    E->hpart = w.floats((size_t)E->head_blocks * (2 * SB_H + 2));
    P->ws_bytes = align_up(w.cur, 256);
    return P;
}

void sefd_fsn_plan_free_ext(sefd_plan* P) {
    delete P->fsn;
    P->fsn = nullptr;
}

int sefd_fsn_forward_impl(const sefd_plan* P, const float* prm, const float* noisy_mag, int train, float dropout_p,
                          const float* mask_fb, const float* mask_sb, unsigned long long seed, float* crm, void* wsv,
                          size_t ws_bytes, cudaStream_t st) {
    SEFD_REQUIRE(P->kind == 2 && P->fsn, "fsn_forward: not a FullSubNet plan");
    SEFD_REQUIRE(ws_bytes >= P->ws_bytes, "fsn_forward: workspace too small (%zu < %zu)", ws_bytes, P->ws_bytes);
    SEFD_REQUIRE(((uintptr_t)wsv & 255) == 0 && ((uintptr_t)prm & 15) == 0, "fsn_forward: workspace / params misaligned");
    SEFD_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "fsn_forward: dropout probability %f outside [0, 1)", dropout_p);
    const FsnExt& E = *P->fsn;
    float* ws = (float*)wsv;
    double* wsd = (double*)wsv;
    const int B = E.B, Tf = E.Tf, T = E.T, R = E.R;
    const int tf = sefd_get_engine_internal() == 1;
    E.sc.drop_on = train && dropout_p > 0.f;
    E.sc.drop_p = dropout_p; E.sc.seed = seed; E.mask_fb = mask_fb; E.mask_sb = mask_sb;

    // ---- packed operands ----
    SEFD_TRY(pack_stack(E.fb, prm, ws, tf, st));
    SEFD_TRY(pack_stack(E.sb, prm, ws, tf, st));
    sefd_absorb_stale_error();
    pack_linear_kernel<<<148, 256, 0, st>>>(prm + E.fb.fc_w, prm + E.fb.fc_b, FBINS, FPAD, FB_H, ws + E.Wl_nk, ws + E.Wl_kn, ws + E.bl, tf);
    SEFD_TRY(sefd_check_launch("fsn_pack_linear"));

    // ---- full-band model (models.py:645-646) ----
    {
        SefdProfScope prof(SEFD_PROF_STFT, 0, 4.0 * B * FBINS * (Tf + (double)T), st);
        fsn_rowsum_kernel<<<(B * FBINS * 32 + 255) / 256, 256, 0, st>>>(noisy_mag, B * FBINS, Tf, wsd + E.rowsum);
        SEFD_TRY(sefd_check_launch("fsn_rowsum"));
        fsn_mu_kernel<<<B, 256, 0, st>>>(wsd + E.rowsum, T, ws + E.inv, wsd + E.wsum);
        SEFD_TRY(sefd_check_launch("fsn_mu"));
        fsn_fb_in_kernel<<<dim3((T + 31) / 32, FPAD / 32, B), dim3(32, 8), 0, st>>>(noisy_mag, ws + E.inv, B, Tf, T, ws + E.fb_in, tf);
        SEFD_TRY(sefd_check_launch("fsn_fb_in"));
    }
    SEFD_TRY(stack_forward(E.sc, E.fb, ws, ws + E.fb_in, T, tf, mask_fb, 1u, st));
    SEFD_TRY(gemm_all_steps(ws + E.fb.l[1].h, FB_H, ws + E.fb_lin, FPAD, B, T, ws + E.Wl_kn, ws + E.Wl_nk, ws + E.bl, 0, st));

    // ---- sub-band input (models.py:649-664) ----
    {
        SefdProfScope prof(SEFD_PROF_STFT, 0, 4.0 * ((double)T * R * SB_I + 2.0 * B * FBINS * T), st);
        cudaMemsetAsync(wsd + E.sum2, 0, sizeof(double) * B, st);
        fsn_relusum_kernel<<<dim3(T, B), 128, 0, st>>>(ws + E.fb_lin, B, wsd + E.sum2);
        SEFD_TRY(sefd_check_launch("fsn_relusum"));
        fsn_unfold_kernel<<<dim3((T + UT - 1) / UT, B), 256, 0, st>>>(noisy_mag, ws + E.fb_lin, wsd + E.wsum, wsd + E.sum2, B, Tf, T,
                                                                     ws + E.inv + B, ws + E.sb_in, tf);
        SEFD_TRY(sefd_check_launch("fsn_unfold"));
    }
    // ---- sub-band model (models.py:667) ----
    SEFD_TRY(stack_forward(E.sc, E.sb, ws, ws + E.sb_in, T, tf, mask_sb, 2u, st));
    {
        SefdProfScope prof(SEFD_PROF_MISC, 4.0 * T * R * SB_H, 4.0 * T * R * SB_H, st);
        sb_head_fwd_kernel<SB_H><<<148 * 8, 256, 0, st>>>(ws + E.sb.l[1].h, prm + E.sb.fc_w, prm + E.sb.fc_b, crm, T, R, Tf);
        SEFD_TRY(sefd_check_launch("fsn_sb_head"));
    }
    return 0;
}

int sefd_fsn_backward_impl(const sefd_plan* P, const float* prm, const float* d_crm, float* grads, void* wsv, size_t ws_bytes,
                           cudaStream_t st) {
    SEFD_REQUIRE(P->kind == 2 && P->fsn, "fsn_backward: not a FullSubNet plan");
    SEFD_REQUIRE(ws_bytes >= P->ws_bytes, "fsn_backward: workspace too small");
    const FsnExt& E = *P->fsn;
    float* ws = (float*)wsv;
    double* wsd = (double*)wsv;
    const int B = E.B, Tf = E.Tf, T = E.T, R = E.R;
    const int tf = sefd_get_engine_internal() == 1;

    // ---- sub-band head ----
    {
        SefdProfScope prof(SEFD_PROF_MISC, 8.0 * T * R * SB_H, 8.0 * T * R * SB_H, st);
        sb_head_bwd_kernel<SB_H><<<E.head_blocks, 256, 0, st>>>(ws + E.sb.l[1].h, prm + E.sb.fc_w, d_crm, ws + E.sb.dh[1], ws + E.hpart, T, R, Tf);
        SEFD_TRY(sefd_check_launch("fsn_sb_head_bwd"));
        sb_head_fold_kernel<<<(2 * SB_H + 2 + 127) / 128, 128, 0, st>>>(ws + E.hpart, E.head_blocks, SB_H, grads + E.sb.fc_w, grads + E.sb.fc_b);
        SEFD_TRY(sefd_check_launch("fsn_sb_head_fold"));
    }
    // The full-band chain (unfold backward -> Linear data gradient -> two recurrences of ONE 128-row tile: 8 CTAs for ~9 ms,
    // latency-bound) depends only on the sub-band RECURRENCES; the sub-band weight gradients (~16 ms on all SMs) depend on
    // nothing after them.  With the tensor-core engine the chain therefore runs on a side stream BESIDE the sub-band weight
    // gradients, which leave 8 SMs to it; the full-band weight gradients (they share the partial buffer) follow the join.
    static const bool overlap_on = getenv("SEFD_FSN_OVERLAP") == nullptr || atoi(getenv("SEFD_FSN_OVERLAP")) != 0;
    // (not while per-launch profiling is on: event timings of kernels that share the device with another stream are meaningless)
    const bool overlap = overlap_on && !sefd_prof_on() && tf && E.sb.l[0].tiled && E.sb.l[1].tiled;
    if (overlap && !P->side) {
        SEFD_REQUIRE(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking) == cudaSuccess, "fsn_backward: side stream");
        cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&P->ev_join, cudaEventDisableTiming);
    }
    // ---- sub-band LSTMs: recurrences (and, without the overlap, their weight gradients) ----
    SEFD_TRY(stack_backward(E.sc, E.sb, ws, ws + E.sb_in, T, tf, E.mask_sb, 2u, ws + E.dsb_in, grads, st, overlap ? 1 : 3));
    cudaStream_t sc = st;                      // stream of the full-band chain
    if (overlap) {
        cudaEventRecord(P->ev_fork, st);
        cudaStreamWaitEvent(P->side, P->ev_fork, 0);
        sc = P->side;
    }
    // ---- normalisation / unfold backward: only the full-band output carries a gradient (models.py:649-658) ----
    {
        SefdProfScope prof(SEFD_PROF_STFT, 0, 8.0 * T * R * SB_I, sc);
        cudaMemsetAsync(wsd + E.Sred, 0, sizeof(double) * B, sc);
        fsn_unfold_bwd_reduce_kernel<<<dim3(T, B), 256, 0, sc>>>(ws + E.dsb_in, ws + E.sb_in, B, wsd + E.Sred);
        SEFD_TRY(sefd_check_launch("fsn_unfold_bwd_reduce"));
        fsn_unfold_bwd_kernel<<<148 * 4, 256, 0, sc>>>(ws + E.dsb_in, ws + E.fb_lin, wsd + E.Sred, ws + E.inv + B, B, T, ws + E.dfb_lin, tf);
        SEFD_TRY(sefd_check_launch("fsn_unfold_bwd"));
    }
    // ---- full-band Linear (+ ReLU, folded into d lin): data gradient, then the recurrences (the input is data: no dx) ----
    SEFD_TRY(gemm_all_steps(ws + E.dfb_lin, FPAD, ws + E.fb.dh[1], FB_H, B, T, ws + E.Wl_nk, ws + E.Wl_kn, nullptr, 0, sc));
    SEFD_TRY(stack_backward(E.sc, E.fb, ws, ws + E.fb_in, T, tf, E.mask_fb, 1u, nullptr, grads, sc, overlap ? 1 : 3));
    if (overlap) {
        cudaEventRecord(P->ev_join, sc);
        // meanwhile, on the main stream: the sub-band weight gradients on all SMs but the 8 of the full-band cluster
        sefd_wgrad_tc_set_cta_limit(148 - 8);
        const int rc = stack_backward(E.sc, E.sb, ws, ws + E.sb_in, T, tf, E.mask_sb, 2u, ws + E.dsb_in, grads, st, 2);
        sefd_wgrad_tc_set_cta_limit(0);
        SEFD_TRY(rc);
        cudaStreamWaitEvent(st, P->ev_join, 0);
        SEFD_TRY(stack_backward(E.sc, E.fb, ws, ws + E.fb_in, T, tf, E.mask_fb, 1u, nullptr, grads, st, 2));
    }
    // ---- full-band Linear: weight and bias gradients ----
    {
        int nsplit = 1;
        long long sstride = 0;
        float* part = ws + E.sc.wpart;
        SEFD_TRY(wgrad_all_steps(ws + E.fb.l[1].h, FB_H, ws + E.dfb_lin, FPAD, B, T, 0, part, E.sc.wpart_floats, &nsplit, &sstride, st));
        fold_linear_kernel<<<dim3(FPAD / 32, FB_H / 32), dim3(32, 8), 0, st>>>(part, nsplit, sstride, FB_H, FPAD, FBINS, grads + E.fb.fc_w);
        SEFD_TRY(sefd_check_launch("fsn_fold_linear"));
        SEFD_TRY(sefd_colsum2(ws + E.dfb_lin, 1, 0, (long long)T * B, FPAD, FPAD, wsd + E.red, ws + E.bl, st));   // bl is free after the forward
        cudaMemcpyAsync(grads + E.fb.fc_b, ws + E.bl, sizeof(float) * FBINS, cudaMemcpyDeviceToDevice, st);
    }
    return 0;
}

int sefd_fsn_tensor_info(const sefd_plan* P, const char* name, long long* off, int* ndim, long long shape[4]) {
    SEFD_REQUIRE(P->fsn, "tensor_info: not a FullSubNet plan");
    const FsnExt& E = *P->fsn;
    const long long B = E.B, T = E.T, R = E.R;
    auto set = [&](size_t o, long long a, long long b, long long c) {
        *off = (long long)o; *ndim = 3; shape[0] = a; shape[1] = b; shape[2] = c; shape[3] = 1;
        return 0;
    };
    const std::string n(name);
    if (n == "fb_in") return set(E.fb_in, T, B, FPAD);
    if (n == "fb_lin") return set(E.fb_lin, T, B, FPAD);
    if (n == "dfb_lin") return set(E.dfb_lin, T, B, FPAD);
    if (n == "sb_in") return set(E.sb_in, T, R, SB_I);
    if (n == "dsb_in") return set(E.dsb_in, T, R, SB_I);
    if (n.size() == 6 && (n.compare(0, 3, "fb.") == 0 || n.compare(0, 3, "sb.") == 0)) {
        const SeqStack& S = n[0] == 'f' ? E.fb : E.sb;
        const int l = n[5] - '0';
        SEFD_REQUIRE(l == 0 || l == 1, "tensor_info: bad name %s", name);
        const SeqLayer& L = S.l[l];
        if (n[3] == 'h' && n[4] == '.') return set(L.h, T, S.rows, L.H);          // "sb.h.1"
        if (n[3] == 'c' && n[4] == '.') return set(L.c, T, S.rows, L.H);
        if (n[3] == 'g' && n[4] == '.') return set(L.gates, T, S.rows, 4 * L.H);
        if (n[3] == 'd' && n[4] == '.') return set(S.dh[l], T, S.rows, L.H);      // "sb.d.1": gradient arriving at h of layer 1
    }
    sefd_set_error("tensor_info: unknown FullSubNet tensor '%s'", name);
    return -1;
}
