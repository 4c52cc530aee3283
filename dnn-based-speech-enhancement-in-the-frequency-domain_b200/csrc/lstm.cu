// Recurrent part of the complex LSTM (tools_for_model.py:141-181; nn.LSTM, H = 128, gate order i,f,g,o).
// The four passes of NavieComplexLSTM (real/imag LSTM x real/imag input) are two LSTMs run on the
// batch-concatenated rows [real inputs ; imag inputs]; the input projections x W_ih^T + b are computed for
// all time steps by the tap-GEMM beforehand, so these kernels only do the T sequential steps.
//
// One CTA (512 threads) owns R rows of one LSTM for the whole sequence.  W_hh (512 x 128 fp32 = 256 KB)
// does not fit in shared memory, so every thread keeps half of its weight row in registers and the other
// half in shared memory (128 KB); h / gate exchange goes through shared memory with two barriers a step.
#include "lstm.cuh"
#include "prof.cuh"

namespace {

constexpr int H = 128, G4 = 512;   // KR of a thread's 128 weights live in registers, H-KR in shared memory

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int R, int KR>
__global__ void __launch_bounds__(512, 1) lstm_fwd_kernel(const LstmFwdParams p) {
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                    // [H-KR][512]
    float* hs = Ws + (H - KR) * G4;    // [R][H]
    float* gs = hs + R * H;            // [R][512]
    const int g = threadIdx.x, lstm = blockIdx.y, row0 = blockIdx.x * R;
    const float* W = p.Whh + ((size_t)lstm * G4 + g) * H;
    float wreg[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) wreg[k] = W[k];
    for (int k = 0; k < H - KR; ++k) Ws[k * G4 + g] = W[KR + k];
    for (int i = g; i < R * H; i += 512) hs[i] = 0.f;

    const int cr = g >> 7, cj = g & 127;
    float* Gr[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        valid[r] = row0 + r < p.rows;
        Gr[r] = p.G + ((size_t)lstm * p.rows + (valid[r] ? row0 + r : 0)) * p.T * G4;
    }
    const bool cvalid = cr < R && row0 + cr < p.rows;
    const size_t hbase = ((size_t)lstm * p.rows + (cvalid ? row0 + cr : 0)) * p.T * H + cj;
    float c = 0.f;
    // the input projections of the next PF steps are prefetched
    constexpr int PF = 1;      // deeper prefetch costs registers (spills at R = 1) and did not pay: the step is shared-memory bound
    float pre[PF][R];
#pragma unroll
    for (int d = 0; d < PF; ++d)
#pragma unroll
        for (int r = 0; r < R; ++r) pre[d][r] = (valid[r] && d < p.T) ? Gr[r][(size_t)d * G4 + g] : 0.f;
    const int gate = g >> 7;
    __syncthreads();

    for (int t0 = 0; t0 < p.T; t0 += PF) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        const int t = t0 + d;
        if (t >= p.T) break;
        float acc[R], acc2[R];   // two independent FMA chains per row
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = pre[d][r]; acc2[r] = 0.f; }
        if (t + PF < p.T) {
#pragma unroll
            for (int r = 0; r < R; ++r) pre[d][r] = valid[r] ? Gr[r][(size_t)(t + PF) * G4 + g] : 0.f;
        }
#pragma unroll
        for (int k4 = 0; k4 < KR / 4; ++k4) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 h4 = *reinterpret_cast<const float4*>(&hs[r * H + 4 * k4]);
                acc[r] = fmaf(wreg[4 * k4 + 0], h4.x, acc[r]);
                acc2[r] = fmaf(wreg[4 * k4 + 1], h4.y, acc2[r]);
                acc[r] = fmaf(wreg[4 * k4 + 2], h4.z, acc[r]);
                acc2[r] = fmaf(wreg[4 * k4 + 3], h4.w, acc2[r]);
            }
        }
#pragma unroll 4
        for (int k4 = 0; k4 < (H - KR) / 4; ++k4) {
            const float w0 = Ws[(4 * k4 + 0) * G4 + g], w1 = Ws[(4 * k4 + 1) * G4 + g];
            const float w2 = Ws[(4 * k4 + 2) * G4 + g], w3 = Ws[(4 * k4 + 3) * G4 + g];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 h4 = *reinterpret_cast<const float4*>(&hs[r * H + KR + 4 * k4]);
                acc[r] = fmaf(w0, h4.x, acc[r]);
                acc2[r] = fmaf(w1, h4.y, acc2[r]);
                acc[r] = fmaf(w2, h4.z, acc[r]);
                acc2[r] = fmaf(w3, h4.w, acc2[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float pa = acc[r] + acc2[r];
            const float a = gate == 2 ? tanhf(pa) : sigmoidf_(pa);
            gs[r * G4 + g] = a;
            if (valid[r]) Gr[r][(size_t)t * G4 + g] = a;
        }
        __syncthreads();
        if (cr < R) {
            const float ig = gs[cr * G4 + cj], fg = gs[cr * G4 + H + cj];
            const float gg = gs[cr * G4 + 2 * H + cj], og = gs[cr * G4 + 3 * H + cj];
            c = fmaf(fg, c, ig * gg);
            const float h = og * tanhf(c);
            hs[cr * H + cj] = h;
            if (cvalid) {
                p.Hh[hbase + (size_t)t * H] = h;
                p.Cc[hbase + (size_t)t * H] = c;
            }
        }
        __syncthreads();
      }
    }
}


template <int R, int KR>
__global__ void __launch_bounds__(512, 1) lstm_bwd_kernel(const LstmBwdParams p) {
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                     // [H-KR... as gate index][512 threads]
    float* dgs = Ws + (H - KR) * G4;    // [R][512]
    float* part = dgs + R * G4;         // [4][R][H]
    const int tid = threadIdx.x, lstm = blockIdx.y, row0 = blockIdx.x * R;
    const int q = tid >> 7, k = tid & 127;          // mat-vec role: gate block q, hidden unit k
    const int cr = q, cj = k;                       // cell role: row cr, hidden unit cj
    const float* W = p.Whh + (size_t)lstm * G4 * H;
    float wreg[KR];
#pragma unroll
    for (int gg = 0; gg < KR; ++gg) wreg[gg] = W[(size_t)(q * H + gg) * H + k];
    for (int gg = 0; gg < H - KR; ++gg) Ws[gg * G4 + tid] = W[(size_t)(q * H + KR + gg) * H + k];

    const bool cvalid = cr < R && row0 + cr < p.rows;
    const size_t row = (size_t)lstm * p.rows + (cvalid ? row0 + cr : 0);
    const float* Gp = p.G + row * p.T * G4 + cj;
    const float* Cp = p.Cc + row * p.T * H + cj;
    const float* dHp = p.dH + row * p.T * H + cj;
    float* dGp = p.dG + row * p.T * G4 + cj;

    float dc_next = 0.f, dh_rec = 0.f;
    // software prefetch, PF steps deep (a step is shorter than a DRAM round trip): slot d holds the operands of the
    // step that runs (d) iterations from now; c_t chains through n_cprev
    constexpr int PF = 3;
    float n_i[PF], n_f[PF], n_g[PF], n_o[PF], n_c[PF], n_cprev[PF], n_dh[PF];
#pragma unroll
    for (int d = 0; d < PF; ++d) {
        n_i[d] = n_f[d] = n_g[d] = n_o[d] = n_c[d] = n_cprev[d] = n_dh[d] = 0.f;
        const int t = p.T - 1 - d;
        if (cvalid && t >= 0) {
            n_i[d] = Gp[(size_t)t * G4]; n_f[d] = Gp[(size_t)t * G4 + H]; n_g[d] = Gp[(size_t)t * G4 + 2 * H]; n_o[d] = Gp[(size_t)t * G4 + 3 * H];
            n_c[d] = Cp[(size_t)t * H];
            n_cprev[d] = t > 0 ? Cp[(size_t)(t - 1) * H] : 0.f;
            n_dh[d] = dHp[(size_t)t * H];
        }
    }
    __syncthreads();

    for (int t0 = p.T - 1; t0 >= 0; t0 -= PF) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        const int t = t0 - d;
        if (t < 0) break;
        if (cr < R) {
            const float ig = n_i[d], fg = n_f[d], gg = n_g[d], og = n_o[d], cprev = n_cprev[d], dho = n_dh[d];
            const float ct = n_c[d];
            {
                const int u = t - PF;
                if (cvalid && u >= 0) {
                    n_i[d] = Gp[(size_t)u * G4]; n_f[d] = Gp[(size_t)u * G4 + H]; n_g[d] = Gp[(size_t)u * G4 + 2 * H]; n_o[d] = Gp[(size_t)u * G4 + 3 * H];
                    n_c[d] = Cp[(size_t)u * H];
                    n_cprev[d] = u > 0 ? Cp[(size_t)(u - 1) * H] : 0.f;
                    n_dh[d] = dHp[(size_t)u * H];
                }
            }
            const float dh = dho + dh_rec;
            const float tc = tanhf(ct);
            const float dc = fmaf(dh * og, 1.f - tc * tc, dc_next);
            float dpi = dc * gg * ig * (1.f - ig);
            float dpf = dc * cprev * fg * (1.f - fg);
            float dpg = dc * ig * (1.f - gg * gg);
            float dpo = dh * tc * og * (1.f - og);
            dc_next = dc * fg;
            dgs[cr * G4 + cj] = dpi;
            dgs[cr * G4 + H + cj] = dpf;
            dgs[cr * G4 + 2 * H + cj] = dpg;
            dgs[cr * G4 + 3 * H + cj] = dpo;
            if (cvalid) {
                if (p.round_tf32) {
                    dGp[(size_t)t * G4] = tf32_rn(dpi); dGp[(size_t)t * G4 + H] = tf32_rn(dpf);
                    dGp[(size_t)t * G4 + 2 * H] = tf32_rn(dpg); dGp[(size_t)t * G4 + 3 * H] = tf32_rn(dpo);
                } else {
                    dGp[(size_t)t * G4] = dpi; dGp[(size_t)t * G4 + H] = dpf;
                    dGp[(size_t)t * G4 + 2 * H] = dpg; dGp[(size_t)t * G4 + 3 * H] = dpo;
                }
            }
        }
        __syncthreads();
        float acc[R], acc2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = acc2[r] = 0.f;
#pragma unroll
        for (int g4 = 0; g4 < KR / 4; ++g4) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 d4 = *reinterpret_cast<const float4*>(&dgs[r * G4 + q * H + 4 * g4]);
                acc[r] = fmaf(wreg[4 * g4 + 0], d4.x, acc[r]);
                acc2[r] = fmaf(wreg[4 * g4 + 1], d4.y, acc2[r]);
                acc[r] = fmaf(wreg[4 * g4 + 2], d4.z, acc[r]);
                acc2[r] = fmaf(wreg[4 * g4 + 3], d4.w, acc2[r]);
            }
        }
#pragma unroll 4
        for (int g4 = 0; g4 < (H - KR) / 4; ++g4) {
            const float w0 = Ws[(4 * g4 + 0) * G4 + tid], w1 = Ws[(4 * g4 + 1) * G4 + tid];
            const float w2 = Ws[(4 * g4 + 2) * G4 + tid], w3 = Ws[(4 * g4 + 3) * G4 + tid];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 d4 = *reinterpret_cast<const float4*>(&dgs[r * G4 + q * H + KR + 4 * g4]);
                acc[r] = fmaf(w0, d4.x, acc[r]);
                acc2[r] = fmaf(w1, d4.y, acc2[r]);
                acc[r] = fmaf(w2, d4.z, acc[r]);
                acc2[r] = fmaf(w3, d4.w, acc2[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) part[(q * R + r) * H + k] = acc[r] + acc2[r];
        __syncthreads();
        if (cr < R)
            dh_rec = part[(0 * R + cr) * H + cj] + part[(1 * R + cr) * H + cj] + part[(2 * R + cr) * H + cj] +
                     part[(3 * R + cr) * H + cj];
      }
    }
}

// rows per CTA: as few as keeps the grid within one wave of SMs (more CTAs = shorter per-step critical path)
template <int R, int KR>
int launch_fwd(const LstmFwdParams& p, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((H - KR) * G4 + R * H + R * G4);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(lstm_fwd_kernel<R, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid((p.rows + R - 1) / R, p.nl ? p.nl : 2);
    sefd_absorb_stale_error();
    lstm_fwd_kernel<R, KR><<<grid, 512, smem, st>>>(p);
    return sefd_check_launch("lstm_fwd");
}
template <int R, int KR>
int launch_bwd(const LstmBwdParams& p, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((H - KR) * G4 + R * G4 + 4 * R * H);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(lstm_bwd_kernel<R, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid((p.rows + R - 1) / R, p.nl ? p.nl : 2);
    sefd_absorb_stale_error();
    lstm_bwd_kernel<R, KR><<<grid, 512, smem, st>>>(p);
    return sefd_check_launch("lstm_bwd");
}
int pick_rows(int rows2) {
    const int rows = (rows2 + 1) / 2;     // callers pass rows * nl; the table below is written for two LSTMs
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (2 * rows <= sms) return 1;
    if (rows <= sms) return 2;
    return 4;
}

}  // namespace

int sefd_lstm_fwd_launch(const LstmFwdParams& p, cudaStream_t st) {
    if (const int RC = sefd_lstm_cluster_rows(p.rows, p.nl ? p.nl : 2)) {
        sefd_prof_label("lstm_cluster_fwd rows%d T%d R%d", p.rows, p.T, RC);
        SefdProfScope prof(SEFD_PROF_LSTM, 4.0 * p.rows * p.T * 512.0 * 128.0, 4.0 * 2 * p.rows * p.T * (2 * 512.0 + 256.0), st);
        return sefd_lstm_cluster_fwd(p, RC, st);
    }
    const int R = pick_rows(p.rows * (p.nl ? p.nl : 2));
    sefd_prof_label("lstm_fwd rows%d T%d R%d", p.rows, p.T, R);
    SefdProfScope prof(SEFD_PROF_LSTM, 4.0 * p.rows * p.T * 512.0 * 128.0, 4.0 * 2 * p.rows * p.T * (2 * 512.0 + 256.0), st);
    if (R == 1) return launch_fwd<1, 88>(p, st);
    if (R == 2) return launch_fwd<2, 80>(p, st);
    return launch_fwd<4, 64>(p, st);
}

int sefd_lstm_bwd_launch(const LstmBwdParams& p, cudaStream_t st) {
    if (const int RC = sefd_lstm_cluster_rows(p.rows, p.nl ? p.nl : 2)) {
        sefd_prof_label("lstm_cluster_bwd rows%d T%d R%d", p.rows, p.T, RC);
        SefdProfScope prof(SEFD_PROF_LSTM, 4.0 * p.rows * p.T * 512.0 * 128.0, 4.0 * 2 * p.rows * p.T * (2 * 512.0 + 384.0), st);
        return sefd_lstm_cluster_bwd(p, RC, st);
    }
    const int R = pick_rows(p.rows * (p.nl ? p.nl : 2));
    sefd_prof_label("lstm_bwd rows%d T%d R%d", p.rows, p.T, R);
    SefdProfScope prof(SEFD_PROF_LSTM, 4.0 * p.rows * p.T * 512.0 * 128.0, 4.0 * 2 * p.rows * p.T * (2 * 512.0 + 384.0), st);
    if (R == 1) return launch_bwd<1, 76>(p, st);
    if (R == 2) return launch_bwd<2, 80>(p, st);
    return launch_bwd<4, 64>(p, st);
}
