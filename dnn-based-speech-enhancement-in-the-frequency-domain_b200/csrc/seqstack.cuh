// Two-layer nn.LSTM stacks on the time-major layer engine (lstm_seq.cuh), shared by the FullSubNet orchestration (fsnet.cu:
// SequenceModel, tools_for_model.py:726-795) and DCCRN's cfg.lstm = 'real' branch (dccrn.cu: models.py:96-105): parameter /
// workspace bookkeeping, packing, forward, backward with weight gradients.
#pragma once
#include <string.h>

#include "lstm_seq.cuh"
#include "plan.cuh"

namespace {

struct SeqLayer {
    int I_real, I, H;
    int kd = 1;                                          // input-column permutation (lstm_seq.cuh), layer 0 of DCCRN's real LSTM
    long long w_ih, w_hh, b_ih, b_hh;                    // parameter offsets
    size_t Wih_nk, Wih_kn, Whh_nk, Whh_kn, bias, Wcat;   // packed operands (workspace, floats); Whh_kn follows Wih_kn directly
    size_t gates, h, c;                                  // gates / c sized for rows rounded up to 128 (tile-major in the fused path)
    mutable int tiled = 0;                               // layout the last forward left gates / c in (lstm_seq.cuh)
};
struct SeqStack {
    SeqLayer l[2];
    int rows;
    size_t h0d;                                          // inter-layer dropout output [T][rows][H]
    size_t dh[2];                                        // gradient arriving at h of layer l from above [T][rows][H]
    long long fc_w, fc_b;
};

// scratch shared by the stacks of one plan + the dropout state of the last forward
struct SeqScratch {
    size_t dh_rec, dc, bias_part, wpart;      // workspace offsets (floats)
    long long wpart_floats;
    int drop_on = 0;
    float drop_p = 0.f;
    unsigned long long seed = 0;
};

SeqLstmWeights weights_of(const SeqLayer& L, const float* ws) {
    SeqLstmWeights w;
    w.Wih_nk = ws + L.Wih_nk; w.Wih_kn = ws + L.Wih_kn; w.Whh_nk = ws + L.Whh_nk; w.Whh_kn = ws + L.Whh_kn; w.bias = ws + L.bias;
    w.Wcat_nk = ws + L.Wcat;
    w.I = L.I; w.H = L.H;
    return w;
}

TapSrc tm_src(const float* p, int rows, int T, int C) {      // time-major [T][rows][C] as [B = 1][F = T]["T" = rows][C]
    TapSrc s;
    s.p = p; s.sT = C; s.sF = (long long)rows * C; s.sB = (long long)T * rows * C; s.C = C;
    return s;
}
TapDst tm_dst(float* p, int rows, int T, int N) {
    TapDst d;
    d.p = p; d.sT = N; d.sF = (long long)rows * N; d.sB = (long long)T * rows * N; d.N = N;
    return d;
}
// out[t][r][:] = a[t][r][:] W + bias over all steps
int gemm_all_steps(const float* a, int K, float* out, int N, int rows, int T, const float* Wkn, const float* Wnk, const float* bias,
                   int round_out, cudaStream_t st) {
    TapGemmParams g;
    memset(&g, 0, sizeof(g));
    g.a[0] = tm_src(a, rows, T, K);
    g.o[0] = tm_dst(out, rows, T, N);
    g.W = Wkn; g.Wnk = Wnk; g.nslabs = 1; g.bias = bias;
    g.B = 1; g.J = T; g.Tout = rows; g.Fin = T; g.Tin = rows;
    g.fi_mul = 1; g.fo_mul = 1; g.fo_off = 0; g.ntaps = 1;
    g.round_out[0] = round_out;
    return sefd_tapgemm(g, st);
}
// partial[s][k][n] = sum over steps j < J and rows of a[j][r][k] g[j][r][n]
int wgrad_all_steps(const float* a, int K, const float* g, int N, int rows, int J, int g_tiled, float* part, long long cap, int* nsplit,
                    long long* sstride, cudaStream_t st) {
    WgradParams w;
    memset(&w, 0, sizeof(w));
    w.a[0] = tm_src(a, rows, J, K);
    w.g = tm_src(g, rows, J, N);
    w.g_tiled = g_tiled;
    w.B = 1; w.J = J; w.Tg = rows; w.Fa = J; w.Ta = rows; w.Fg = J;
    w.a_mul = 1; w.g_mul = 1; w.ntaps = 1;
    w.rows_per_cta = 1;
    return sefd_wgrad(w, part, cap, 1, nsplit, sstride, st);
}

int pack_stack(const SeqStack& S, const float* prm, float* ws, int tf, cudaStream_t st) {
    for (int l = 0; l < 2; ++l) {
        const SeqLayer& L = S.l[l];
        SeqLstmPackParams p;
        p.w_ih = prm + L.w_ih; p.w_hh = prm + L.w_hh; p.b_ih = prm + L.b_ih; p.b_hh = prm + L.b_hh;
        p.I_real = L.I_real; p.I = L.I; p.H = L.H;
        p.Wih_nk = ws + L.Wih_nk; p.Wih_kn = ws + L.Wih_kn; p.Whh_nk = ws + L.Whh_nk; p.Whh_kn = ws + L.Whh_kn; p.bias = ws + L.bias;
        p.Wcat_nk = ws + L.Wcat;
        p.round_tf32 = tf;
        p.kd = L.kd;
        SEFD_TRY(sefd_seqlstm_pack(p, st));
    }
    return 0;
}

int stack_forward(const SeqScratch& E, const SeqStack& S, float* ws, const float* x, int T, int tf, const float* mask, unsigned int stream_id,
                  cudaStream_t st) {
    for (int l = 0; l < 2; ++l) {
        const SeqLayer& L = S.l[l];
        SeqLstmFwdParams p;
        p.x = l == 0 ? x : (E.drop_on ? ws + S.h0d : ws + S.l[0].h);
        p.w = weights_of(L, ws);
        p.gates = ws + L.gates; p.h = ws + L.h; p.c = ws + L.c;
        p.rows = S.rows; p.T = T; p.round_h = tf; p.h_zero_slot = 1;
        cudaMemsetAsync(ws + L.h - (size_t)S.rows * L.H, 0, sizeof(float) * S.rows * L.H, st);     // h_{-1} = 0
        SEFD_TRY(sefd_seqlstm_forward(p, st));
        L.tiled = p.tiled;
        if (l == 0 && E.drop_on)
            SEFD_TRY(sefd_dropout_apply(ws + L.h, ws + S.h0d, (long long)T * S.rows * L.H, E.drop_p, mask, E.seed, stream_id, tf, st));
    }
    return 0;
}

// backward through the two layers; dh[1] holds the gradient arriving at h1.  dx0 (gradient w.r.t. the stack input) is
// written when non-null.
// phases: 1 = the recurrences (LSTM backward kernels of both layers, bias folds, input gradients, dropout backward), 2 = the
// weight gradients (two wgrad calls + folds per layer; they only read what phase 1 left: dG in the gates buffers, the layer
// inputs and h), 3 = both in the reference order.  Phase 2 may run later and on another stream than phase 1 (fsnet.cu
// overlaps the sub-band weight gradients with the full-band recurrences); it owns E.wpart, phase 1 owns dh_rec / dc / bias_part.
int stack_backward(const SeqScratch& E, const SeqStack& S, float* ws, const float* x, int T, int tf, const float* mask, unsigned int stream_id,
                   float* dx0, float* grads, cudaStream_t st, int phases = 3) {
    const int rows = S.rows;
    for (int l = 1; l >= 0; --l) {
        const SeqLayer& L = S.l[l];
        const int N = 4 * L.H;
        if (phases & 1) {
            SeqLstmBwdParams p;
            p.w = weights_of(L, ws);
            p.gates = ws + L.gates; p.c = ws + L.c; p.dh_out = ws + S.dh[l];
            p.dh_rec = ws + E.dh_rec; p.dc = ws + E.dc; p.bias_part = ws + E.bias_part;
            p.rows = rows; p.T = T; p.round_tf32 = tf;
            p.dx = l == 1 ? ws + S.dh[0] : dx0;
            p.dx_done = 0;
            p.tiled = L.tiled;
            SEFD_TRY(sefd_seqlstm_backward(p, st));
            SEFD_TRY(sefd_seqlstm_fold_bias(ws + E.bias_part, p.bias_blocks, L.H, grads + L.b_ih, grads + L.b_hh, st));
            float* dx = l == 1 ? ws + S.dh[0] : dx0;
            if (dx) {
                if (!p.dx_done) {
                    SEFD_REQUIRE(!L.tiled, "fsn backward: the input gradient must come from the fused step kernel when dG is tile-major");
                    SEFD_TRY(gemm_all_steps(ws + L.gates, N, dx, L.I, rows, T, L.Wih_nk + ws, L.Wih_kn + ws, nullptr, 0, st));
                }
                if (l == 1 && E.drop_on)
                    SEFD_TRY(sefd_dropout_apply(dx, dx, (long long)T * rows * L.I, E.drop_p, mask, E.seed, stream_id, 0, st));
            }
        }
        if (phases & 2) {
            const long long step_stride = L.tiled ? (long long)((rows + 127) / 128) * 128 * N : (long long)rows * N;
            const float* dG = ws + L.gates;
            const float* xin = l == 0 ? x : (E.drop_on ? ws + S.h0d : ws + S.l[0].h);
            int nsplit = 1;
            long long sstride = 0;
            float* part = ws + E.wpart;
            SEFD_TRY(wgrad_all_steps(xin, L.I, dG, N, rows, T, L.tiled, part, E.wpart_floats, &nsplit, &sstride, st));
            SEFD_TRY(sefd_seqlstm_fold_wgrad(part, nsplit, sstride, L.I, L.I_real, L.H, grads + L.w_ih, st, L.kd));
            if (T > 1) {
                SEFD_TRY(wgrad_all_steps(ws + L.h, L.H, dG + step_stride, N, rows, T - 1, L.tiled, part, E.wpart_floats, &nsplit, &sstride, st));
                SEFD_TRY(sefd_seqlstm_fold_wgrad(part, nsplit, sstride, L.H, L.H, L.H, grads + L.w_hh, st));
            } else {
                cudaMemsetAsync(grads + L.w_hh, 0, sizeof(float) * N * L.H, st);
            }
        }
    }
    return 0;
}

void carve_stack(SeqStack& S, Carver& w, int rows, int T) {
    S.rows = rows;
    for (int l = 0; l < 2; ++l) {
        SeqLayer& L = S.l[l];
        const size_t N = 4 * (size_t)L.H;
        L.Wih_nk = w.floats(N * L.I);
        L.Whh_nk = w.floats(N * L.H);
        L.Whh_kn = w.floats(N * (L.I + L.H));        // [H][4H'] immediately followed by [I][4H'] = the backward's [W_hh^T ; W_ih^T]
        L.Wih_kn = L.Whh_kn + N * L.H;
        L.Wcat = w.floats(N * (L.I + L.H));
        L.bias = w.floats(N);
        const size_t rpad = (size_t)(rows + 127) / 128 * 128;
        L.gates = w.floats((size_t)T * rpad * N);
        L.h = w.floats((size_t)(T + 1) * rows * L.H) + (size_t)rows * L.H;     // one zero step in front (h_{-1})
        L.c = w.floats((size_t)T * rpad * L.H);
        S.dh[l] = w.floats((size_t)T * rows * L.H);
    }
    S.h0d = w.floats((size_t)T * rows * S.l[0].H);
}


// scratch sizes for stacks whose largest state is `state_floats` = max over stacks of roundup(rows, 128) * H and whose
// widest layer has `maxH` hidden units / `maxK` input + hidden columns
inline void carve_seq_scratch(SeqScratch& E, Carver& w, size_t state_floats, int max_rows, int maxH, int maxK) {
    E.dh_rec = w.floats(state_floats);
    E.dc = w.floats(state_floats);
    E.bias_part = w.floats((size_t)sefd_seqlstm_bias_blocks(max_rows) * 4 * maxH);
    E.wpart_floats = 32ll * 4 * maxH * maxK;
    E.wpart = w.floats((size_t)E.wpart_floats);
}

}  // namespace
