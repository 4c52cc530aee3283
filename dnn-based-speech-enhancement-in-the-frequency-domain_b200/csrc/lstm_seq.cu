// Time-major nn.LSTM layer (see lstm_seq.cuh): weight packing into the interleaved gate order, the per-step
// recurrence as [generic tap-GEMM + cell kernel] (every engine; the exact-precision path of the parity tests) and the
// dispatch to the fused tcgen05 step kernels of lstm_step_tc.cu, inter-layer dropout.
// Reference: nn.LSTM inside SequenceModel (tools_for_model.py:741-748, 785-786), gate order i, f, g, o, zero initial
// state, two bias vectors.
#include <stdlib.h>
#include <string.h>

#include "lstm_seq.cuh"
#include "lstm_step_tc.cuh"
#include "prof.cuh"

namespace {

constexpr int RB = 32;   // rows per CTA of the cell kernels (one bias-gradient slot per CTA)

__device__ __forceinline__ int interleave(int n, int H) {   // reference column n = gate * H + u -> stored column n'
    const int g = n / H, u = n - g * H;
    return (u >> 3) * 32 + g * 8 + (u & 7);
}
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) {
    // 1 - 2 / (exp(2x) + 1): full relative accuracy away from 0, absolute error ~1e-7 near 0
    return 1.f - 2.f / (__expf(2.f * x) + 1.f);
}

__global__ void pack_kernel(const SeqLstmPackParams p) {
    const int N = 4 * p.H;
    const long long nih = (long long)N * p.I, nhh = (long long)N * p.H;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nih + nhh + N; e += (long long)gridDim.x * blockDim.x) {
        if (e < nih) {
            const int n = (int)(e / p.I), k = (int)(e % p.I);
            const int ks = p.kd > 1 ? (k % (p.I_real / p.kd)) * p.kd + k / (p.I_real / p.kd) : k;     // reference column of stored column k
            float v = k < p.I_real ? p.w_ih[(long long)n * p.I_real + ks] : 0.f;
            if (p.round_tf32) v = tf32_rn(v);
            const int np = interleave(n, p.H);
            p.Wih_nk[(long long)np * p.I + k] = v;
            p.Wih_kn[(long long)k * N + np] = v;
            if (p.Wcat_nk) p.Wcat_nk[(long long)np * (p.I + p.H) + k] = v;
        } else if (e < nih + nhh) {
            const long long r = e - nih;
            const int n = (int)(r / p.H), k = (int)(r % p.H);
            float v = p.w_hh[r];
            if (p.round_tf32) v = tf32_rn(v);
            const int np = interleave(n, p.H);
            p.Whh_nk[(long long)np * p.H + k] = v;
            p.Whh_kn[(long long)k * N + np] = v;
            if (p.Wcat_nk) p.Wcat_nk[(long long)np * (p.I + p.H) + p.I + k] = v;
        } else {
            const int n = (int)(e - nih - nhh);
            p.bias[interleave(n, p.H)] = p.b_ih[n] + p.b_hh[n];
        }
    }
}

__global__ void fold_wgrad_kernel(const float* __restrict__ part, int nsplit, long long split_stride, int K, int K_real, int H,
                                  float* __restrict__ dW, int kd) {
    // tile transpose: read [k][n'] coalesced over n', write [n][k] coalesced over k
    __shared__ float tile[32][33];
    const int N = 4 * H;
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, np = n0 + threadIdx.x;
        float s = 0.f;
        if (k < K)
            for (int sp = 0; sp < nsplit; ++sp) s += part[sp * split_stride + (long long)k * N + np];
        tile[r][threadIdx.x] = s;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int np = n0 + r, k = k0 + threadIdx.x;
        // inverse of the interleave: n' = ub * 32 + g * 8 + i
        const int ub = np >> 5, g = (np >> 3) & 3, i8 = np & 7;
        const int n = g * H + ub * 8 + i8;
        const int ks = kd > 1 ? (k % (K_real / kd)) * kd + k / (K_real / kd) : k;
        if (k < K_real) dW[(long long)n * K_real + ks] = tile[threadIdx.x][r];
    }
}

__global__ void fold_bias_kernel(const float* __restrict__ part, int nblk, int H, float* db_ih, float* db_hh) {
    const int N = 4 * H;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int np = interleave(n, H);
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += (double)part[(long long)b * N + np];
    db_ih[n] = (float)s;
    db_hh[n] = (float)s;
}

// ---- cell kernels (generic path): one CTA = RB rows x all hidden units, thread = 4 consecutive units ----------------
struct CellFwd {
    float* gates;          // [rows][4H'] in: pre-activations (bias and both projections included), out: activated
    const float* c_prev;   // [rows][H] or null (t = 0)
    float *c, *h;
    int rows, H, round_h;
};
__global__ void cell_fwd_kernel(const CellFwd p) {
    const int H4 = p.H >> 2;
    const int u4 = threadIdx.x % H4, rl = threadIdx.x / H4, rstep = blockDim.x / H4;
    const int u = u4 * 4;
    const int col = (u >> 3) * 32 + (u & 7);      // gate g of these 4 units sits at col + 8 g
    const int r1 = min(p.rows, (int)(blockIdx.x + 1) * RB);
    for (int r = blockIdx.x * RB + rl; r < r1; r += rstep) {
        float* g = p.gates + (long long)r * 4 * p.H + col;
        float4 gi = *reinterpret_cast<float4*>(g), gf = *reinterpret_cast<float4*>(g + 8);
        float4 gg = *reinterpret_cast<float4*>(g + 16), go = *reinterpret_cast<float4*>(g + 24);
        float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.c_prev) cp = *reinterpret_cast<const float4*>(p.c_prev + (long long)r * p.H + u);
        float4 c, h;
#define CELL(x)                                                   \
        gi.x = sigmoid_(gi.x); gf.x = sigmoid_(gf.x); gg.x = tanh_(gg.x); go.x = sigmoid_(go.x); \
        c.x = fmaf(gf.x, cp.x, gi.x * gg.x);                      \
        h.x = go.x * tanh_(c.x);
        CELL(x) CELL(y) CELL(z) CELL(w)
#undef CELL
        if (p.round_h) { h.x = tf32_rn(h.x); h.y = tf32_rn(h.y); h.z = tf32_rn(h.z); h.w = tf32_rn(h.w); }
        *reinterpret_cast<float4*>(g) = gi; *reinterpret_cast<float4*>(g + 8) = gf;
        *reinterpret_cast<float4*>(g + 16) = gg; *reinterpret_cast<float4*>(g + 24) = go;
        *reinterpret_cast<float4*>(p.c + (long long)r * p.H + u) = c;
        *reinterpret_cast<float4*>(p.h + (long long)r * p.H + u) = h;
    }
}

struct CellBwd {
    float* gates;          // in: activated gates of step t; out: dG of step t
    const float *c, *c_prev, *dh_out, *dh_rec;   // c_prev null at t = 0; dh_rec null at t = T-1
    float* dc;             // [rows][H] carried (read unless first, written)
    float* bias_part;      // [nblk][4H']
    int rows, H, first, round_tf32;
};
__global__ void cell_bwd_kernel(const CellBwd p) {
    extern __shared__ float s_red[];    // [rstep][4H] column partials of the row lanes
    const int H4 = p.H >> 2;
    const int u4 = threadIdx.x % H4, rl = threadIdx.x / H4, rstep = blockDim.x / H4;
    const int u = u4 * 4;
    const int col = (u >> 3) * 32 + (u & 7);      // gate g of these 4 units sits at col + 8 g
    const int r1 = min(p.rows, (int)(blockIdx.x + 1) * RB);
    float4 si = make_float4(0.f, 0.f, 0.f, 0.f), sf = si, sg = si, so = si;
    for (int r = blockIdx.x * RB + rl; r < r1; r += rstep) {
        float* g = p.gates + (long long)r * 4 * p.H + col;
        const float4 gi = *reinterpret_cast<float4*>(g), gf = *reinterpret_cast<float4*>(g + 8);
        const float4 gg = *reinterpret_cast<float4*>(g + 16), go = *reinterpret_cast<float4*>(g + 24);
        const long long o = (long long)r * p.H + u;
        const float4 ct = *reinterpret_cast<const float4*>(p.c + o);
        float4 cp = make_float4(0.f, 0.f, 0.f, 0.f), dcn = cp;
        if (p.c_prev) cp = *reinterpret_cast<const float4*>(p.c_prev + o);
        float4 dh = *reinterpret_cast<const float4*>(p.dh_out + o);
        if (p.dh_rec) {
            const float4 d2 = *reinterpret_cast<const float4*>(p.dh_rec + o);
            dh.x += d2.x; dh.y += d2.y; dh.z += d2.z; dh.w += d2.w;
        }
        if (!p.first) dcn = *reinterpret_cast<const float4*>(p.dc + o);
        float4 di, df, dg, d_o, dcc;
#define CELLB(x)                                                          \
        {                                                                 \
            const float tc = tanh_(ct.x);                                 \
            const float dc = fmaf(dh.x * go.x, 1.f - tc * tc, dcn.x);     \
            di.x = dc * gg.x * gi.x * (1.f - gi.x);                       \
            df.x = dc * cp.x * gf.x * (1.f - gf.x);                       \
            dg.x = dc * gi.x * (1.f - gg.x * gg.x);                       \
            d_o.x = dh.x * tc * go.x * (1.f - go.x);                      \
            dcc.x = dc * gf.x;                                            \
        }
        CELLB(x) CELLB(y) CELLB(z) CELLB(w)
#undef CELLB
        si.x += di.x; si.y += di.y; si.z += di.z; si.w += di.w;
        sf.x += df.x; sf.y += df.y; sf.z += df.z; sf.w += df.w;
        sg.x += dg.x; sg.y += dg.y; sg.z += dg.z; sg.w += dg.w;
        so.x += d_o.x; so.y += d_o.y; so.z += d_o.z; so.w += d_o.w;
        if (p.round_tf32) {
#define RND(v) v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w);
            RND(di) RND(df) RND(dg) RND(d_o)
#undef RND
        }
        *reinterpret_cast<float4*>(g) = di; *reinterpret_cast<float4*>(g + 8) = df;
        *reinterpret_cast<float4*>(g + 16) = dg; *reinterpret_cast<float4*>(g + 24) = d_o;
        *reinterpret_cast<float4*>(p.dc + o) = dcc;
    }
    // column sums of this CTA's rows -> its own slot (fixed order: deterministic, no atomics)
    float* mine = s_red + (long long)rl * 4 * p.H + col;
    *reinterpret_cast<float4*>(mine) = si; *reinterpret_cast<float4*>(mine + 8) = sf;
    *reinterpret_cast<float4*>(mine + 16) = sg; *reinterpret_cast<float4*>(mine + 24) = so;
    __syncthreads();
    for (int n = threadIdx.x; n < 4 * p.H; n += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < rstep; ++l) s += s_red[l * 4 * p.H + n];
        float* slot = p.bias_part + (long long)blockIdx.x * 4 * p.H + n;
        *slot = p.first ? s : *slot + s;
    }
}

// ---- Philox4x32-10 (Salmon et al., SC'11): counter = element index / 4, key = seed ------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__global__ void dropout_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4, float p, const float* __restrict__ mask,
                               unsigned long long seed, unsigned int stream_id, int round_tf32) {
    const float scale = 1.f / (1.f - p);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(src)[e];
        float4 m;
        if (mask) {
            m = reinterpret_cast<const float4*>(mask)[e];
        } else {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32), stream_id, 0u),
                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            // keep with probability 1 - p: u = r * 2^-32 in [0, 1)
            const float t = p * 4294967296.f;
            m.x = (float)r.x >= t ? scale : 0.f; m.y = (float)r.y >= t ? scale : 0.f;
            m.z = (float)r.z >= t ? scale : 0.f; m.w = (float)r.w >= t ? scale : 0.f;
        }
        v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
        if (round_tf32) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }
        reinterpret_cast<float4*>(dst)[e] = v;
    }
}

int cell_threads(int H) {   // (H / 4) * row lanes, <= 512
    const int H4 = H / 4;
    int lanes = 512 / H4;
    if (lanes < 1) lanes = 1;
    if (lanes > 4) lanes = 4;
    return H4 * lanes;
}

TapGemmParams step_gemm(const float* a, int K, float* out, int N, int rows, const float* Wkn, const float* Wnk, int accum,
                        int round_out) {
    TapGemmParams g;
    memset(&g, 0, sizeof(g));
    g.a[0].p = a; g.a[0].sT = K; g.a[0].sF = (long long)rows * K; g.a[0].sB = (long long)rows * K; g.a[0].C = K;
    g.o[0].p = out; g.o[0].sT = N; g.o[0].sF = (long long)rows * N; g.o[0].sB = (long long)rows * N; g.o[0].N = N;
    g.W = Wkn; g.Wnk = Wnk; g.nslabs = 1;
    g.B = 1; g.J = 1; g.Tout = rows; g.Fin = 1; g.Tin = rows;
    g.fi_mul = 0; g.fo_mul = 1; g.fo_off = 0;
    g.ntaps = 1;
    g.accum[0] = accum;
    g.round_out[0] = round_out;
    return g;
}

}  // namespace

int sefd_seqlstm_fused_enabled() {
    static const int on = getenv("SEFD_LSTM_FUSED") == nullptr || atoi(getenv("SEFD_LSTM_FUSED")) != 0;
    return on;
}

int sefd_seqlstm_pack(const SeqLstmPackParams& p, cudaStream_t st) {
    SEFD_REQUIRE(p.H % 64 == 0 && p.I % 4 == 0 && p.I >= p.I_real, "seqlstm_pack: H=%d I=%d unsupported", p.H, p.I);
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    const long long n = 4ll * p.H * (p.I + p.H + 1);
    long long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    pack_kernel<<<(int)g, 256, 0, st>>>(p);
    return sefd_check_launch("seqlstm_pack");
}

int sefd_seqlstm_fold_wgrad(const float* part, int nsplit, long long split_stride, int K, int K_real, int H, float* dW,
                            cudaStream_t st, int kd) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    dim3 grid(4 * H / 32, (K + 31) / 32), block(32, 8);
    fold_wgrad_kernel<<<grid, block, 0, st>>>(part, nsplit, split_stride, K, K_real, H, dW, kd);
    return sefd_check_launch("seqlstm_fold_wgrad");
}

int sefd_seqlstm_fold_bias(const float* part, int nblk, int H, float* db_ih, float* db_hh, cudaStream_t st) {
    SefdProfScope prof(SEFD_PROF_MISC, 0, 0, st);
    fold_bias_kernel<<<(4 * H + 127) / 128, 128, 0, st>>>(part, nblk, H, db_ih, db_hh);
    return sefd_check_launch("seqlstm_fold_bias");
}

int sefd_seqlstm_bias_blocks(int rows) { return (rows + RB - 1) / RB + sefd_lstm_step_bias_blocks(rows); }

int sefd_seqlstm_forward(const SeqLstmFwdParams& p0, cudaStream_t st) {
    SeqLstmFwdParams& p = const_cast<SeqLstmFwdParams&>(p0);
    p.tiled = 0;
    const int H = p.w.H, N = 4 * H, I = p.w.I;
    SEFD_REQUIRE(H % 64 == 0 && H <= 512 && p.rows > 0 && p.T > 0, "seqlstm_forward: H=%d rows=%d T=%d unsupported", H, p.rows, p.T);
    const bool tc = sefd_get_engine_internal() == 1;
    if (tc && sefd_seqlstm_fused_enabled() && p.w.Wcat_nk && p.h_zero_slot && p.w.Wih_kn == p.w.Whh_kn + (long long)H * N &&
        sefd_lstm_step_tc_eligible(I, H)) {
        p.tiled = 1;
        return sefd_lstm_step_tc_forward(p, st);
    }
    // ---- generic path: input projections of all steps as ONE GEMM, then per step [recurrent GEMM (+=) ; cell kernel] ----
    {
        TapGemmParams g = step_gemm(p.x, I, p.gates, N, p.rows, p.w.Wih_kn, p.w.Wih_nk, 0, 0);
        g.J = p.T; g.Fin = p.T; g.fi_mul = 1;
        g.bias = p.w.bias;
        SEFD_TRY(sefd_tapgemm(g, st));
    }
    const int nblk = (p.rows + RB - 1) / RB, nthr = cell_threads(H);
    for (int t = 0; t < p.T; ++t) {
        float* gt = p.gates + (long long)t * p.rows * N;
        if (t > 0) {
            TapGemmParams g = step_gemm(p.h + (long long)(t - 1) * p.rows * H, H, gt, N, p.rows, p.w.Whh_kn, p.w.Whh_nk, 1, 0);
            SEFD_TRY(sefd_tapgemm(g, st));
        }
        CellFwd c;
        c.gates = gt;
        c.c_prev = t > 0 ? p.c + (long long)(t - 1) * p.rows * H : nullptr;
        c.c = p.c + (long long)t * p.rows * H;
        c.h = p.h + (long long)t * p.rows * H;
        c.rows = p.rows; c.H = H; c.round_h = p.round_h;
        SefdProfScope prof(SEFD_PROF_LSTM, 0, 4.0 * p.rows * (2.0 * N + 3.0 * H), st);
        cell_fwd_kernel<<<nblk, nthr, 0, st>>>(c);
        SEFD_TRY(sefd_check_launch("seqlstm_cell_fwd"));
    }
    return 0;
}

int sefd_seqlstm_cell_bwd_step(SeqLstmBwdParams& p, int t, int* nblk_out, cudaStream_t st) {
    const int H = p.w.H, N = 4 * H;
    const int nblk = (p.rows + RB - 1) / RB, nthr = cell_threads(H);
    const size_t smem = sizeof(float) * (nthr / (H / 4)) * N;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(cell_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr = true;
    }
    CellBwd c;
    c.gates = p.gates + (long long)t * p.rows * N;
    c.c = p.c + (long long)t * p.rows * H;
    c.c_prev = t > 0 ? p.c + (long long)(t - 1) * p.rows * H : nullptr;
    c.dh_out = p.dh_out + (long long)t * p.rows * H;
    c.dh_rec = t == p.T - 1 ? nullptr : p.dh_rec;
    c.dc = p.dc;
    c.bias_part = p.bias_part;
    c.rows = p.rows; c.H = H; c.first = t == p.T - 1; c.round_tf32 = p.round_tf32;
    SefdProfScope prof(SEFD_PROF_LSTM, 0, 4.0 * p.rows * (2.0 * N + 5.0 * H), st);
    cell_bwd_kernel<<<nblk, nthr, smem, st>>>(c);
    if (nblk_out) *nblk_out = nblk;
    return sefd_check_launch("seqlstm_cell_bwd");
}

int sefd_seqlstm_backward(const SeqLstmBwdParams& p0, cudaStream_t st) {
    SeqLstmBwdParams& p = const_cast<SeqLstmBwdParams&>(p0);
    p.dx_done = 0;
    const int H = p.w.H, N = 4 * H;
    SEFD_REQUIRE(H % 64 == 0 && H <= 512 && p.rows > 0 && p.T > 0, "seqlstm_backward: H=%d rows=%d T=%d unsupported", H, p.rows, p.T);
    if (p.tiled) {
        SEFD_REQUIRE(sefd_get_engine_internal() == 1 && p.w.Wih_kn == p.w.Whh_kn + (long long)H * N && sefd_lstm_step_tc_eligible(p.w.I, H),
                     "seqlstm_backward: the forward ran fused (tile-major state) but the fused backward cannot run");
        return sefd_lstm_step_tc_backward(p, st);
    }
    for (int t = p.T - 1; t >= 0; --t) {
        SEFD_TRY(sefd_seqlstm_cell_bwd_step(p, t, &p.bias_blocks, st));
        if (t > 0) {   // dh_rec = dG_t W_hh  (contraction over the 4H' gate columns)
            TapGemmParams g = step_gemm(p.gates + (long long)t * p.rows * N, N, p.dh_rec, H, p.rows, p.w.Whh_nk, p.w.Whh_kn, 0, 0);
            SEFD_TRY(sefd_tapgemm(g, st));
        }
    }
    return 0;
}

int sefd_dropout_apply(const float* src, float* dst, long long n, float p, const float* mask, unsigned long long seed,
                       unsigned int stream_id, int round_tf32, cudaStream_t st) {
    SEFD_REQUIRE(n % 4 == 0 && p >= 0.f && p < 1.f, "dropout: n=%lld p=%f unsupported", n, p);
    SefdProfScope prof(SEFD_PROF_MISC, 0, 8.0 * n, st);
    long long g = (n / 4 + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    dropout_kernel<<<(int)g, 256, 0, st>>>(src, dst, n / 4, p, mask, seed, stream_id, round_tf32);
    return sefd_check_launch("dropout");
}
