// DCCRN train-step orchestration: parameter layout, workspace carving, forward and backward
// sequencing of the kernels.  Mirrors DCCRN.forward (models.py:176-284) and the autograd graph the
// reference gets from torch; see DESIGN.md for the dataflow and SURVEY.md appendix B for the
// backward obligations.
#include <stdlib.h>

#include "fsnet.cuh"
#include "plan.cuh"
#include "prof.cuh"
#include "seqstack.cuh"
#include "taps.cuh"

// cfg.lstm = 'real' (models.py:96-105, 213-218): self.enhance = nn.LSTM(1024 -> 256, 2 layers), self.tranform = Linear(256 -> 1024)
// on the [T, B, C * D] view of the encoder output.  Runs on the time-major layer engine (lstm_seq.cuh; B rows <= 128: the
// chunks of the one row tile are split over a cluster).  The LSTM input is stored as D = 4 blocks of C = 256 channels
// (k' = d * 256 + c instead of the reference's c * 4 + d): the packed W_ih columns are permuted to match.
struct RealLstmExt {
    SeqStack st;
    SeqScratch sc;
    long long w_tr, b_tr;                  // tranform.weight [1024][256], tranform.bias [1024]
    size_t x_tm, dx_tm;                    // [T][B][1024]
    size_t Wtrp, WtrT, btrp;               // [d][k][c], [d][c][k], [d][c]
};

namespace {
constexpr int RL_I = 1024, RL_H = 256, RL_C = 256, RL_D = 4;

// x_tm[t][b][d * 256 + c] = z[b][d][t][c]  (forward) / the inverse scatter (backward)
__global__ void real_lstm_gather_kernel(const float* __restrict__ z, float* __restrict__ x, int B, int T, int scatter) {
    const long long n4 = (long long)B * RL_D * T * RL_C / 4;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(e % (RL_C / 4));
        long long r = e / (RL_C / 4);
        const int t = (int)(r % T);
        r /= T;
        const int d = (int)(r % RL_D), b = (int)(r / RL_D);
        const long long xi = (((long long)t * B + b) * RL_D + d) * (RL_C / 4) + c4;
        if (scatter) reinterpret_cast<float4*>(const_cast<float*>(z))[e] = reinterpret_cast<const float4*>(x)[xi];
        else reinterpret_cast<float4*>(x)[xi] = reinterpret_cast<const float4*>(z)[e];
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------------
sefd_plan* sefd_plan_create_impl(int B, int L, int mask_mode, int flags) {
    if (B <= 0 || L <= 0 || L % HOP != 0) {
        sefd_set_error("plan: need B > 0 and L a positive multiple of %d (got B=%d L=%d)", HOP, B, L);
        return nullptr;
    }
    if (!((mask_mode >= SEFD_MASK_E && mask_mode <= SEFD_MASK_R) || mask_mode == SEFD_MASK_DIRECT)) {
        sefd_set_error("plan: masking mode %d unsupported (E=1, C=2, R=3, Direct=5)", mask_mode);
        return nullptr;
    }
    sefd_plan* P = new sefd_plan();
    P->kind = 0;
    P->B = B;
    P->L = L;
    P->T = L / HOP + 3;
    P->mask_mode = mask_mode;
    P->skip = (flags & SEFD_PLAN_NO_SKIP) ? 0 : 1;
    P->cbn = (flags & SEFD_PLAN_CBN) ? 1 : 0;
    // BatchNorm2d(C): weight, bias + running_mean, running_var; ComplexBatchNorm(C): Wrr, Wri, Wii, Br, Bi + RMr, RMi, RVrr, RVri,
    // RVii over C / 2 complex features (tools_for_model.py:444-470), in the reference's registration order
    auto add_norm = [&](ConvLayer& c, const std::string& pre, long long& pc, long long& bc) {
        if (!P->cbn) {
            add_param(P, pre + ".1.weight", pc, &c.gamma, {c.Cout});
            add_param(P, pre + ".1.bias", pc, &c.beta, {c.Cout});
            add_param(P, pre + ".2.weight", pc, &c.alpha, {1});
            add_buffer(P, pre + ".1.running_mean", bc, &c.rmean, c.Cout);
            add_buffer(P, pre + ".1.running_var", bc, &c.rvar, c.Cout);
            return;
        }
        const int h = c.Cout / 2;
        add_param(P, pre + ".1.Wrr", pc, &c.gamma, {h});
        add_param(P, pre + ".1.Wri", pc, &c.wri, {h});
        add_param(P, pre + ".1.Wii", pc, &c.wii, {h});
        add_param(P, pre + ".1.Br", pc, &c.beta, {h});
        add_param(P, pre + ".1.Bi", pc, &c.bi2, {h});
        add_param(P, pre + ".2.weight", pc, &c.alpha, {1});
        add_buffer(P, pre + ".1.RMr", bc, &c.rmean, h);
        add_buffer(P, pre + ".1.RMi", bc, &c.rmi, h);
        add_buffer(P, pre + ".1.RVrr", bc, &c.rvar, h);
        add_buffer(P, pre + ".1.RVri", bc, &c.rvri, h);
        add_buffer(P, pre + ".1.RVii", bc, &c.rvii, h);
    };
    const int kn[NL + 1] = {2, 32, 64, 128, 256, 256, 256};
    for (int i = 0; i <= NL; ++i) {
        P->ch[i] = kn[i];
        P->Fe[i] = 256 >> i;
    }
    long long pc = 0, bc = 0;
    for (int i = 0; i < NL; ++i) {
        ConvLayer& c = P->enc[i];
        c.Cin = kn[i];
        c.Cout = kn[i + 1];
        c.Fin = P->Fe[i];
        c.Fout = P->Fe[i + 1];
        const std::string pre = "encoder." + std::to_string(i);
        add_param(P, pre + ".0.real_conv.weight", pc, &c.wr, {c.Cout / 2, c.Cin / 2, 5, 2});
        add_param(P, pre + ".0.real_conv.bias", pc, &c.br, {c.Cout / 2});
        add_param(P, pre + ".0.imag_conv.weight", pc, &c.wi, {c.Cout / 2, c.Cin / 2, 5, 2});
        add_param(P, pre + ".0.imag_conv.bias", pc, &c.bi, {c.Cout / 2});
        add_norm(c, pre, pc, bc);
    }
    for (int j = 0; j < NL; ++j) {
        ConvLayer& c = P->dec[j];
        const int idx = NL - j;
        c.Cin = (P->skip ? 2 : 1) * kn[idx];
        c.Cout = kn[idx - 1];
        c.Fin = P->Fe[idx];
        c.Fout = 2 * c.Fin;
        const std::string pre = "decoder." + std::to_string(j);
        add_param(P, pre + ".0.real_conv.weight", pc, &c.wr, {c.Cin / 2, c.Cout / 2, 5, 2});
        add_param(P, pre + ".0.real_conv.bias", pc, &c.br, {c.Cout / 2});
        add_param(P, pre + ".0.imag_conv.weight", pc, &c.wi, {c.Cin / 2, c.Cout / 2, 5, 2});
        add_param(P, pre + ".0.imag_conv.bias", pc, &c.bi, {c.Cout / 2});
        if (j != NL - 1) {
            add_norm(c, pre, pc, bc);
        } else {
            c.gamma = c.beta = c.alpha = c.rmean = c.rvar = -1;
        }
    }
    const bool real_lstm = (flags & SEFD_PLAN_REAL_LSTM) != 0;
    if (real_lstm) {
        RealLstmExt* R = new RealLstmExt();
        P->rl = R;
        for (int l = 0; l < 2; ++l) {
            SeqLayer& Lr = R->st.l[l];
            Lr.I_real = Lr.I = l == 0 ? RL_I : RL_H;
            Lr.H = RL_H;
            Lr.kd = l == 0 ? RL_D : 1;
            const std::string sfx = "_l" + std::to_string(l);
            add_param(P, "enhance.weight_ih" + sfx, pc, &Lr.w_ih, {4ll * RL_H, Lr.I_real});
            add_param(P, "enhance.weight_hh" + sfx, pc, &Lr.w_hh, {4ll * RL_H, RL_H});
            add_param(P, "enhance.bias_ih" + sfx, pc, &Lr.b_ih, {4ll * RL_H});
            add_param(P, "enhance.bias_hh" + sfx, pc, &Lr.b_hh, {4ll * RL_H});
        }
        add_param(P, "tranform.weight", pc, &R->w_tr, {RL_I, RL_H});
        add_param(P, "tranform.bias", pc, &R->b_tr, {RL_I});
    }
    for (int l = 0; l < (real_lstm ? 0 : 2); ++l) {
        const long long I = l == 0 ? 512 : 128;
        const char* part[2] = {"real", "imag"};
        for (int p = 0; p < 2; ++p) {
            const std::string pre = "enhance." + std::to_string(l) + "." + part[p] + "_lstm.";
            add_param(P, pre + "weight_ih_l0", pc, &P->w_ih[l][p], {G4, I});
            add_param(P, pre + "weight_hh_l0", pc, &P->w_hh[l][p], {G4, RNN_H});
            add_param(P, pre + "bias_ih_l0", pc, &P->b_ih[l][p], {G4});
            add_param(P, pre + "bias_hh_l0", pc, &P->b_hh[l][p], {G4});
        }
        if (l == 1) {
            const char* tr[2] = {"r", "i"};
            for (int q = 0; q < 2; ++q) {
                const std::string pre = "enhance.1." + std::string(tr[q]) + "_trans.";
                add_param(P, pre + "weight", pc, &P->w_tr[q], {512, RNN_H});
                add_param(P, pre + "bias", pc, &P->b_tr[q], {512});
            }
        }
    }
    P->n_param_floats = pc;
    P->n_buffer_floats = bc;

    // ---- workspace ----
    const size_t Bz = B, T = P->T;
    Carver w;
    P->spec = w.floats(Bz * NBIN * T * 2);
    P->raw_wav = w.floats(Bz * L);
    P->dots = w.doubles(Bz * 8 + 8);
    size_t nstat = 0;
    for (int i = 0; i < NL; ++i) nstat += 2 * P->enc[i].Cout;
    for (int j = 0; j < NL - 1; ++j) nstat += 2 * P->dec[j].Cout;
    P->stats_all = w.doubles(nstat);
    P->stats_all_n = nstat;
    size_t sc = P->stats_all;
    size_t max_y = 0, max_w = 0;
    for (int i = 0; i < NL; ++i) {
        ConvLayer& c = P->enc[i];
        const size_t n = Bz * c.Fout * T * c.Cout;
        c.y = w.floats(n);
        c.z = w.floats(n);
        c.dz = w.floats(n);      // gradient through the skip connection (written by the decoder's data gradient)
        c.dz2 = w.floats(n);     // gradient from the next encoder layer / the LSTM; BN backward sums the two
        c.dy = w.floats(n);      // gradient w.r.t. the raw conv output (own buffer per layer: wgrad and dgrad of
                                 // different layers never alias, so weight gradients may run on a side stream)
        c.Wf = w.floats(10ull * c.Cin * c.Cout);
        c.Wt = w.floats(10ull * c.Cin * c.Cout);
        c.bias = w.floats(c.Cout);
        c.save = w.floats(5 * c.Cout);           // BatchNorm: [2][C]; ComplexBatchNorm: [9][C / 2]
        c.stats = sc;
        sc += 2 * c.Cout;
        if (n > max_y) max_y = n;
        if (10ull * c.Cin * c.Cout > max_w) max_w = 10ull * c.Cin * c.Cout;
    }
    for (int j = 0; j < NL; ++j) {
        ConvLayer& c = P->dec[j];
        const size_t ny = Bz * c.Fout * (T + 1) * c.Cout, nz = Bz * c.Fout * T * c.Cout;
        c.y = w.floats(ny);
        c.dy = w.floats(ny);
        c.z = j != NL - 1 ? w.floats(nz) : 0;
        c.dz = j != NL - 1 ? w.floats(nz) : 0;
        c.Wf = w.floats(10ull * c.Cin * c.Cout);
        c.Wt = w.floats(10ull * c.Cin * c.Cout);
        c.bias = w.floats(c.Cout);
        c.save = w.floats(5 * c.Cout + 4);
        if (j != NL - 1) {
            c.stats = sc;
            sc += 2 * c.Cout;
        }
        if (ny > max_y) max_y = ny;
        if (10ull * c.Cin * c.Cout > max_w) max_w = 10ull * c.Cin * c.Cout;
    }
    for (int l = 0; l < 2; ++l) {
        P->Gt[l] = w.floats(2 * 2 * Bz * T * G4);
        P->Hh[l] = w.floats(2 * 2 * Bz * T * RNN_H);
        P->Cc[l] = w.floats(2 * 2 * Bz * T * RNN_H);
        P->Whh[l] = w.floats(2 * G4 * RNN_H);
        P->bsum[l] = w.floats(2 * G4);
    }
    P->X1 = w.floats(2 * Bz * T * RNN_H);
    P->X2 = w.floats(2 * Bz * T * RNN_H);
    P->U = w.floats(Bz * 4 * T * 256);
    P->Wih0p = w.floats(2ull * 4 * 128 * G4);     // [p][d][c][n]
    P->Wih0T = w.floats(4ull * 2 * G4 * 128);     // [d][p*512+n][c]
    P->Wih1p = w.floats(2ull * 128 * G4);         // [p][k][n]
    P->Wih1T = w.floats(2ull * G4 * 128);         // [p*512+n][k]
    P->Wih0Q = w.floats(4ull * 128 * 2 * G4);     // [d][c][p*512+n]
    P->Wih1Q = w.floats(128ull * 2 * G4);         // [k][p*512+n]
    P->Wtrp = w.floats(2ull * 4 * 128 * 128);     // [q][d][k][c]
    P->WtrT = w.floats(2ull * 4 * 128 * 128);     // [q][d][c][k]
    P->btrp = w.floats(2ull * 4 * 128);           // [q][d][c]
    // backward scratch
    P->dU = w.floats(Bz * 4 * T * 256);
    P->dY_floats = max_y;
    P->dY = 0;
    if (max_w < 4ull * 128 * G4) max_w = 4ull * 128 * G4;
    P->dWs_floats = 16 * max_w;                   // room for the split partials of the tensor-core wgrad
    P->dWs = w.floats(16 * max_w);
    P->dbs = w.floats(1024);
    P->red = w.doubles(2 * 512 + 8);             // BatchNorm backward: 2 C + 1; ComplexBatchNorm backward: 3 C + 1 (C <= 256)
    P->red2 = w.doubles(2 * 512 + 8);
    P->cbn_coef = w.floats(9 * 128);
    for (int i = 0; i < NL; ++i) P->enc[i].cstats = w.doubles(5 * 128);
    for (int j = 0; j < NL; ++j) P->dec[j].cstats = w.doubles(5 * 128);
    P->dX = w.floats(2 * Bz * T * RNN_H);
    P->dH = w.floats(2 * 2 * Bz * T * RNN_H);
    P->dG = w.floats(2 * 2 * Bz * T * G4);
    if (P->rl) {
        RealLstmExt& R = *P->rl;
        R.x_tm = w.floats(T * Bz * RL_I);
        R.dx_tm = w.floats(T * Bz * RL_I);
        R.Wtrp = w.floats((size_t)RL_D * RL_H * RL_C);
        R.WtrT = w.floats((size_t)RL_D * RL_H * RL_C);
        R.btrp = w.floats((size_t)RL_D * RL_C);
        carve_stack(R.st, w, B, (int)T);
        carve_seq_scratch(R.sc, w, (size_t)(B + 127) / 128 * 128 * RL_H, B, RL_H, RL_I);
    }
    P->ws_bytes = align_up(w.cur, 256);
    return P;
}

// ------------------------------------------------------------------------------------------------
// weight packing (once per forward; the packed operands are reused by the backward)
// ------------------------------------------------------------------------------------------------
// part 0: encoder convs (needed first); part 1: decoder convs, LSTM and projection operands (needed after the encoder)
static int pack_weights(const sefd_plan* P, const float* prm, float* ws, int part, cudaStream_t st) {
    for (int e = part ? NL : 0; e < (part ? 2 * NL : NL); ++e) {
        const ConvLayer& c = e < NL ? P->enc[e] : P->dec[e - NL];
        CconvPackParams pp;
        pp.wr = prm + c.wr; pp.wi = prm + c.wi; pp.br = prm + c.br; pp.bi = prm + c.bi;
        pp.Ci2 = c.Cin / 2; pp.Co2 = c.Cout / 2;
        pp.transposed = e >= NL; pp.two_src = e >= NL && P->skip;
        pp.Wf = ws + c.Wf; pp.Wt = ws + c.Wt; pp.bias = ws + c.bias;
        pp.round_tf32 = sefd_get_engine_internal() == 1 && c.Cin % 32 == 0 && (c.Cout % 32 == 0);
        SEFD_TRY(sefd_pack_cconv(pp, st));
    }
    if (!part) return 0;
    const int tf = sefd_get_engine_internal() == 1;
    if (P->rl) {
        const RealLstmExt& R = *P->rl;
        SEFD_TRY(pack_stack(R.st, prm, ws, tf, st));
        Permute3Params q;
        q.src = prm + R.w_tr; q.na = RL_D; q.nb = RL_H; q.nc = RL_C; q.accumulate = 0; q.nsplit = 1; q.split_stride = 0; q.round_tf32 = tf;
        // W_tr [c * 4 + d][k] -> Wtrp[d][k][c]
        q.dst = ws + R.Wtrp; q.sa = RL_H; q.sb = 1; q.sc = (long long)RL_D * RL_H; q.da = (long long)RL_H * RL_C; q.db = RL_C; q.dc = 1;
        SEFD_TRY(sefd_permute3p(q, st));
        // -> WtrT[d][c][k]
        q.dst = ws + R.WtrT; q.nb = RL_C; q.nc = RL_H; q.sb = (long long)RL_D * RL_H; q.sc = 1; q.da = (long long)RL_C * RL_H; q.db = RL_H;
        SEFD_TRY(sefd_permute3p(q, st));
        SEFD_TRY(sefd_permute3(prm + R.b_tr, ws + R.btrp, 1, RL_D, RL_C, 0, 1, RL_D, 0, st));
        return 0;
    }
    auto perm = [&](const float* src, float* dst, int na, int nb, int nc, long long sa, long long sb, long long sc,
                    long long da, long long db, long long dc) -> int {
        Permute3Params q;
        q.src = src; q.dst = dst; q.na = na; q.nb = nb; q.nc = nc; q.sa = sa; q.sb = sb; q.sc = sc;
        q.da = da; q.db = db; q.dc = dc; q.accumulate = 0; q.nsplit = 1; q.split_stride = 0; q.round_tf32 = tf;
        return sefd_permute3p(q, st);
    };
    for (int p = 0; p < 2; ++p) {
        const float* w0 = prm + P->w_ih[0][p];      // [n][c*4+d]
        const float* w1 = prm + P->w_ih[1][p];      // [n][k]
        // layer 0: Wih0p[p][d][c][n], Wih0T[d][p*512+n][c], Wih0Q[d][c][p*512+n]
        SEFD_TRY(perm(w0, ws + P->Wih0p + (size_t)p * 4 * 128 * G4, 4, 128, G4, 1, 4, 512, 128ll * G4, G4, 1));
        SEFD_TRY(perm(w0, ws + P->Wih0T + (size_t)p * G4 * 128, 4, G4, 128, 1, 512, 4, 2ll * G4 * 128, 128, 1));
        SEFD_TRY(perm(w0, ws + P->Wih0Q + (size_t)p * G4, 4, 128, G4, 1, 4, 512, 128ll * 2 * G4, 2ll * G4, 1));
        // layer 1: Wih1p[p][k][n], Wih1T[p*512+n][k], Wih1Q[k][p*512+n]
        SEFD_TRY(perm(w1, ws + P->Wih1p + (size_t)p * 128 * G4, 1, 128, G4, 0, 1, 128, 0, G4, 1));
        SEFD_TRY(perm(w1, ws + P->Wih1T + (size_t)p * G4 * 128, 1, G4, 128, 0, 128, 1, 0, 128, 1));
        SEFD_TRY(perm(w1, ws + P->Wih1Q + (size_t)p * G4, 1, 128, G4, 0, 1, 128, 0, 2ll * G4, 1));
        for (int l = 0; l < 2; ++l) {
            SEFD_TRY(sefd_permute3(prm + P->w_hh[l][p], ws + P->Whh[l] + (size_t)p * G4 * RNN_H, 1, 1, G4 * RNN_H, 0, 0, 1, 0, st));
            SEFD_TRY(sefd_add2(prm + P->b_ih[l][p], prm + P->b_hh[l][p], ws + P->bsum[l] + (size_t)p * G4, G4, st));
        }
    }
    for (int q = 0; q < 2; ++q) {
        // W_tr [c*4+d][k] -> Wtrp[q][d][k][c] ; WtrT[q][d][c][k] ; b_tr[c*4+d] -> btrp[q][d][c]
        const float* wt = prm + P->w_tr[q];
        SEFD_TRY(perm(wt, ws + P->Wtrp + (size_t)q * 4 * 128 * 128, 4, 128, 128, 128, 1, 512, 128 * 128, 128, 1));
        SEFD_TRY(perm(wt, ws + P->WtrT + (size_t)q * 4 * 128 * 128, 4, 128, 128, 128, 512, 1, 128 * 128, 128, 1));
        SEFD_TRY(sefd_permute3(prm + P->b_tr[q], ws + P->btrp + (size_t)q * 4 * 128, 1, 4, 128, 0, 1, 4, 0, st));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
int sefd_forward_impl(const sefd_plan* P, const float* prm, float* bnbuf, const float* noisy, const float* target,
                      int train, float* out_real, float* out_imag, float* out_wav, void* wsv, size_t ws_bytes,
                      cudaStream_t st) {
    SEFD_REQUIRE(P->kind == 0, "dccrn_forward: not a DCCRN plan");
    SEFD_REQUIRE(ws_bytes >= P->ws_bytes, "forward: workspace too small (%zu < %zu)", ws_bytes, P->ws_bytes);
    SEFD_REQUIRE(((uintptr_t)wsv & 255) == 0 && ((uintptr_t)prm & 15) == 0, "forward: workspace/params misaligned");
    float* ws = (float*)wsv;
    double* wsd = (double*)wsv;
    const int B = P->B, T = P->T, L = P->L;
    cudaMemsetAsync(wsd + P->stats_all, 0, sizeof(double) * P->stats_all_n, st);
    // the ~30 small pack / permute launches for the decoder, LSTM and projection operands run on the plan's side stream
    // beside the STFT and the encoder (their consumers start after the encoder); the fork event orders them after every
    // earlier reader of the packed buffers (the previous step's backward)
    static const bool use_side = getenv("SEFD_SIDE_STREAM") == nullptr || atoi(getenv("SEFD_SIDE_STREAM")) != 0;
    cudaStream_t sx = st;
    if (use_side && !sefd_prof_on()) {
        if (!P->side) {
            SEFD_REQUIRE(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking) == cudaSuccess, "forward: side stream");
            cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&P->ev_join, cudaEventDisableTiming);
        }
        sx = P->side;
        cudaEventRecord(P->ev_fork, st);
        cudaStreamWaitEvent(sx, P->ev_fork, 0);
    }
    SEFD_TRY(pack_weights(P, prm, ws, 1, sx));
    if (sx != st) cudaEventRecord(P->ev_join, sx);
    SEFD_TRY(pack_weights(P, prm, ws, 0, st));
    SEFD_TRY(sefd_stft_launch(noisy, ws + P->spec, B, L, T, st));

    auto bn = [&](const ConvLayer& c, int Ty, int tshift) -> int {
        if (P->cbn) {
            CbnPreluFwdParams b;
            memset(&b, 0, sizeof(b));
            b.y = ws + c.y; b.z = ws + c.z;
            b.BF = B * c.Fout; b.Ty = Ty; b.T = T; b.tshift = tshift; b.C = c.Cout;
            b.stats = wsd + c.cstats; b.n_stat = (double)B * c.Fout * Ty;
            b.W[0] = prm + c.gamma; b.W[1] = prm + c.wri; b.W[2] = prm + c.wii;
            b.B2[0] = prm + c.beta; b.B2[1] = prm + c.bi2; b.alpha = prm + c.alpha;
            b.save = ws + c.save;
            if (bnbuf) {
                b.RM[0] = bnbuf + c.rmean; b.RM[1] = bnbuf + c.rmi;
                b.RV[0] = bnbuf + c.rvar; b.RV[1] = bnbuf + c.rvri; b.RV[2] = bnbuf + c.rvii;
            }
            b.momentum = BN_MOM; b.eps = BN_EPS;
            b.use_running = !train;
            b.round_tf32 = sefd_get_engine_internal() == 1;
            if (!train) SEFD_REQUIRE(bnbuf != nullptr, "forward: eval mode needs the ComplexBatchNorm running statistics");
            return sefd_cbn_prelu_fwd(b, st);
        }
        BnPreluFwdParams b;
        memset(&b, 0, sizeof(b));
        b.y = ws + c.y; b.z = ws + c.z;
        b.BF = B * c.Fout; b.Ty = Ty; b.T = T; b.tshift = tshift; b.C = c.Cout;
        b.stats = wsd + c.stats; b.n_stat = (double)B * c.Fout * Ty;
        b.gamma = prm + c.gamma; b.beta = prm + c.beta; b.alpha = prm + c.alpha;
        b.save = ws + c.save;
        b.running_mean = bnbuf ? bnbuf + c.rmean : nullptr;
        b.running_var = bnbuf ? bnbuf + c.rvar : nullptr;
        b.momentum = BN_MOM; b.eps = BN_EPS;
        b.use_running = !train;
        b.round_tf32 = sefd_get_engine_internal() == 1;
        if (!train) SEFD_REQUIRE(bnbuf != nullptr, "forward: eval mode needs the BN running statistics");
        return sefd_bn_prelu_fwd(b, st);
    };

    // ---- encoder ----
    for (int i = 0; i < NL; ++i) {
        const ConvLayer& c = P->enc[i];
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        if (i == 0) {
            g.a[0].p = ws + P->spec + (size_t)T * 2;   // DC bin dropped (models.py:184)
            g.a[0].sT = 2; g.a[0].sF = (long long)T * 2; g.a[0].sB = (long long)NBIN * T * 2; g.a[0].C = 2;
        } else {
            g.a[0] = src4(ws + P->enc[i - 1].z, c.Fin, T, c.Cin, c.Cin);
        }
        g.a[1] = no_src();
        g.o[0] = dst4(ws + c.y, c.Fout, T, c.Cout, c.Cout);
        g.o[1] = no_dst();
        g.W = ws + c.Wf; g.Wnk = ws + c.Wt; g.nslabs = 10; g.bias = ws + c.bias;
        g.stats = train ? wsd + c.stats : nullptr;
        g.B = B; g.J = c.Fout; g.Tout = T; g.Fin = c.Fin; g.Tin = T;
        conv_taps_down(g, -1);
        SEFD_TRY(sefd_tapgemm(g, st));
        SEFD_TRY(bn(c, T, 0));
    }

    if (sx != st) cudaStreamWaitEvent(st, P->ev_join, 0);      // packed LSTM / projection / decoder operands are ready
    if (P->rl) {
        // ---- cfg.lstm = 'real': one 2-layer nn.LSTM on [T, B, 1024] + tranform (models.py:213-218) ----
        const RealLstmExt& R = *P->rl;
        const int tf = sefd_get_engine_internal() == 1;
        sefd_absorb_stale_error();
        real_lstm_gather_kernel<<<148 * 4, 256, 0, st>>>(ws + P->enc[NL - 1].z, ws + R.x_tm, B, T, 0);
        SEFD_TRY(sefd_check_launch("real_lstm_gather"));
        SEFD_TRY(stack_forward(R.sc, R.st, ws, ws + R.x_tm, T, tf, nullptr, 0u, st));
        // out[t][b][c * 4 + d] -> U[b][d][t][c]: four output rows d, the source is time-major h1 [T][B][256]
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0].p = ws + R.st.l[1].h; g.a[0].sB = RL_H; g.a[0].sF = 0; g.a[0].sT = (long long)B * RL_H; g.a[0].C = RL_H;
        g.a[1] = no_src();
        g.o[0] = dst4(ws + P->U, RL_D, T, RL_C, RL_C);
        g.o[1] = no_dst();
        g.W = ws + R.Wtrp; g.wJ = (long long)RL_H * RL_C;
        g.Wnk = ws + R.WtrT; g.nslabs = RL_D; g.wJ_slabs = 1;
        g.round_out[0] = tf;                                    // U feeds decoder 0
        g.bias = ws + R.btrp; g.bJ = RL_C;
        g.B = B; g.J = RL_D; g.Tout = T; g.Fin = 1; g.Tin = T;
        g.fi_mul = 0; g.fo_mul = 1; g.fo_off = 0;
        g.ntaps = 1;
        SEFD_TRY(sefd_tapgemm(g, st));
    }
    // ---- complex LSTM x2 (tools_for_model.py:162-177) ----
    const size_t rowsz = (size_t)T * G4;
    for (int l = 0; l < (P->rl ? 0 : 2); ++l) {
        for (int p = 0; p < 2; ++p)
            for (int q = 0; q < 2; ++q) {
                TapGemmParams g;
                memset(&g, 0, sizeof(g));
                if (l == 0) {
                    g.a[0] = src4(ws + P->enc[NL - 1].z + q * 128, 4, T, 256, 128);
                    g.ntaps = 4;
                    for (int d = 0; d < 4; ++d) { g.df[d] = d; g.dt[d] = 0; g.wslab[d] = d; }
                    g.Fin = 4;
                    g.W = ws + P->Wih0p + (size_t)p * 4 * 128 * G4;
                    g.Wnk = ws + P->Wih0T + (size_t)p * G4 * 128; g.nslabs = 4; g.w_slab_stride = 2ll * G4 * 128;
                } else {
                    g.a[0] = src4(ws + P->X1 + (size_t)q * B * T * RNN_H, 1, T, RNN_H, RNN_H);
                    g.ntaps = 1;
                    g.Fin = 1;
                    g.W = ws + P->Wih1p + (size_t)p * 128 * G4;
                    g.Wnk = ws + P->Wih1T + (size_t)p * G4 * 128; g.nslabs = 1;
                }
                g.a[1] = no_src();
                g.o[0] = dst4(ws + P->Gt[l] + ((size_t)p * 2 + q) * B * rowsz, 1, T, G4, G4);
                g.o[1] = no_dst();
                g.bias = ws + P->bsum[l] + (size_t)p * G4;
                g.B = B; g.J = 1; g.Tout = T; g.Tin = T;
                g.fi_mul = 0; g.fo_mul = 1; g.fo_off = 0;
                SEFD_TRY(sefd_tapgemm(g, st));
            }
        LstmFwdParams lp;
        memset(&lp, 0, sizeof(lp));
        lp.Whh = ws + P->Whh[l]; lp.G = ws + P->Gt[l]; lp.Hh = ws + P->Hh[l]; lp.Cc = ws + P->Cc[l];
        lp.rows = 2 * B; lp.T = T;
        SEFD_TRY(sefd_lstm_fwd_launch(lp, st));
        SEFD_TRY(sefd_clstm_combine(ws + P->Hh[l], ws + (l == 0 ? P->X1 : P->X2), (long long)B * T * RNN_H,
                                    sefd_get_engine_internal() == 1, st));
    }
    // projection r_trans / i_trans (Linear 128 -> 512), output feature c*4+d -> U[b][d][t][q*128+c]
    for (int q = 0; q < (P->rl ? 0 : 2); ++q) {
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(ws + P->X2 + (size_t)q * B * T * RNN_H, 1, T, RNN_H, RNN_H);
        g.a[1] = no_src();
        g.o[0] = dst4(ws + P->U + q * 128, 4, T, 256, 128);
        g.o[1] = no_dst();
        g.W = ws + P->Wtrp + (size_t)q * 4 * 128 * 128; g.wJ = 128 * 128;
        g.Wnk = ws + P->WtrT + (size_t)q * 4 * 128 * 128; g.nslabs = 4; g.wJ_slabs = 1;
        g.round_out[0] = sefd_get_engine_internal() == 1;      // U feeds decoder 0
        g.bias = ws + P->btrp + (size_t)q * 4 * 128; g.bJ = 128;
        g.B = B; g.J = 4; g.Tout = T; g.Fin = 1; g.Tin = T;
        g.fi_mul = 0; g.fo_mul = 1; g.fo_off = 0;
        g.ntaps = 1;
        SEFD_TRY(sefd_tapgemm(g, st));
    }

    // ---- decoder (models.py:222-226): convT on complex_cat(out, skip), BN over T+1 frames, drop frame 0 ----
    for (int j = 0; j < NL; ++j) {
        const ConvLayer& c = P->dec[j];
        const int Ch = P->skip ? c.Cin / 2 : c.Cin;       // channels per source
        const float* in0 = j == 0 ? ws + P->U : ws + P->dec[j - 1].z;
        const float* in1 = ws + P->enc[NL - 1 - j].z;
        if (P->skip && sefd_skinny_up_n2_eligible(Ch, c.Cout)) {       // decoder 5: both phases in one HBM-bound pass
            SEFD_TRY(sefd_skinny_up_n2(in0, in1, ws + c.Wf, ws + c.bias, ws + c.y, B, c.Fin, T, st));
            continue;
        }
        {
            TapGemmParams g;
            memset(&g, 0, sizeof(g));
            g.a[0] = src4(in0, c.Fin, T, Ch, Ch);
            g.a[1] = P->skip ? src4(in1, c.Fin, T, Ch, Ch) : no_src();
            g.o[0] = dst4(ws + c.y, c.Fout, T + 1, c.Cout, c.Cout);
            g.o[1] = no_dst();
            g.W = ws + c.Wf; g.Wnk = ws + c.Wt; g.nslabs = 10; g.bias = ws + c.bias;
            g.stats = (train && j != NL - 1) ? wsd + c.stats : nullptr;
            g.B = B; g.J = c.Fin; g.Tout = T + 1; g.Fin = c.Fin; g.Tin = T;
            SEFD_TRY(sefd_tapgemm_up(g, 0, st));   // both output-row phases (one fused launch on tcgen05)
        }
        if (j != NL - 1) SEFD_TRY(bn(c, T + 1, 1));
    }

    // ---- mask, ISTFT, clamp ----
    const ConvLayer& last = P->dec[NL - 1];
    MaskIstftParams m;
    memset(&m, 0, sizeof(m));
    m.spec = ws + P->spec;
    m.mask = ws + last.y;
    m.mT = 2; m.mF = (long long)(T + 1) * 2; m.mB = (long long)256 * (T + 1) * 2;
    m.m_tshift = 1;
    m.mode = P->mask_mode;
    m.B = B; m.L = L; m.T = T;
    m.out_real = out_real; m.out_imag = out_imag; m.out_wav = out_wav;
    m.raw_wav = ws + P->raw_wav;
    m.target = target; m.dots = wsd + P->dots;
    return sefd_mask_istft_launch(m, st);
}

// ------------------------------------------------------------------------------------------------
int sefd_backward_impl(const sefd_plan* P, const float* prm, const float* dwav, const float* dreal, const float* dimag,
                       float* grads, void* wsv, size_t ws_bytes, cudaStream_t st, cudaEvent_t tail_ready) {
    SEFD_REQUIRE(ws_bytes >= P->ws_bytes, "backward: workspace too small");
    float* ws = (float*)wsv;
    double* wsd = (double*)wsv;
    const int B = P->B, T = P->T, L = P->L;
    float* dWs = ws + P->dWs;

    auto bn_bwd = [&](const ConvLayer& c, int Ty, int tshift, bool two) -> int {
        if (P->cbn) {
            CbnPreluBwdParams b;
            memset(&b, 0, sizeof(b));
            b.y = ws + c.y; b.dz = ws + c.dz; b.dy = ws + c.dy;
            b.dz2 = two ? ws + c.dz2 : nullptr;
            b.BF = B * c.Fout; b.Ty = Ty; b.T = T; b.tshift = tshift; b.C = c.Cout;
            b.n_stat = (double)B * c.Fout * Ty;
            b.W[0] = prm + c.gamma; b.W[1] = prm + c.wri; b.W[2] = prm + c.wii;
            b.B2[0] = prm + c.beta; b.B2[1] = prm + c.bi2; b.alpha = prm + c.alpha; b.save = ws + c.save;
            b.red = wsd + P->red; b.coef = ws + P->cbn_coef;
            b.dW[0] = grads + c.gamma; b.dW[1] = grads + c.wri; b.dW[2] = grads + c.wii;
            b.dB2[0] = grads + c.beta; b.dB2[1] = grads + c.bi2; b.dalpha = grads + c.alpha;
            b.round_tf32 = sefd_get_engine_internal() == 1;
            return sefd_cbn_prelu_bwd(b, st);
        }
        BnPreluBwdParams b;
        memset(&b, 0, sizeof(b));
        b.y = ws + c.y; b.dz = ws + c.dz; b.dy = ws + c.dy;
        b.dz2 = two ? ws + c.dz2 : nullptr;
        b.BF = B * c.Fout; b.Ty = Ty; b.T = T; b.tshift = tshift; b.C = c.Cout;
        b.n_stat = (double)B * c.Fout * Ty;
        b.gamma = prm + c.gamma; b.beta = prm + c.beta; b.alpha = prm + c.alpha; b.save = ws + c.save;
        b.red = wsd + P->red;
        b.dgamma = grads + c.gamma; b.dbeta = grads + c.beta; b.dalpha = grads + c.alpha;
        b.round_tf32 = sefd_get_engine_internal() == 1;
        return sefd_bn_prelu_bwd(b, st);
    };
    int nsplit = 1;
    long long sstride = 0;
    // ---- side stream: everything that only post-processes a finished weight gradient (fold / un-permute of the split
    // partials, bias column sums) is off the critical chain bn_bwd -> dgrad -> bn_bwd ...; these kernels are tiny
    // (no shared memory, few CTAs) and co-reside with the persistent GEMM CTAs.  Hazards: the partial buffer dWs and the
    // LSTM gate-gradient buffer dG are overwritten by later main-stream kernels -> join before those.
    static const bool use_side = getenv("SEFD_SIDE_STREAM") == nullptr || atoi(getenv("SEFD_SIDE_STREAM")) != 0;
    cudaStream_t sx = st;
    bool pending = false;
    if (use_side && !sefd_prof_on()) {
        if (!P->side) {
            SEFD_REQUIRE(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking) == cudaSuccess, "backward: side stream");
            cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&P->ev_join, cudaEventDisableTiming);
        }
        sx = P->side;
    }
    auto fork = [&]() {           // the side stream sees everything enqueued on the main stream so far
        if (sx != st) { cudaEventRecord(P->ev_fork, st); cudaStreamWaitEvent(sx, P->ev_fork, 0); }
    };
    auto side_done = [&]() {
        if (sx != st) { cudaEventRecord(P->ev_join, sx); pending = true; }
    };
    auto join = [&]() {           // the main stream waits for the side work issued so far
        if (pending) { cudaStreamWaitEvent(st, P->ev_join, 0); pending = false; }
    };
    double* red_side = wsd + (sx != st ? P->red2 : P->red);
    auto fold = [&](const ConvLayer& c, bool dec, const float* dbias) -> int {
        CconvFoldParams f;
        f.dWf = dWs; f.dbias = dbias; f.nsplit = nsplit; f.split_stride = sstride;
        f.Ci2 = c.Cin / 2; f.Co2 = c.Cout / 2; f.transposed = dec; f.two_src = dec && P->skip;
        f.dwr = grads + c.wr; f.dwi = grads + c.wi; f.dbr = grads + c.br; f.dbi = grads + c.bi;
        return sefd_fold_cconv(f, sx);
    };

    // sum the wgrad split partials while un-permuting into the reference's parameter layout
    auto unperm = [&](const float* src, float* dst, int na, int nb, int nc, long long sa, long long sb, long long sc,
                      int accumulate) -> int {
        Permute3Params q;
        q.src = src; q.dst = dst; q.na = na; q.nb = nb; q.nc = nc; q.sa = sa; q.sb = sb; q.sc = sc;
        q.da = (long long)nb * nc; q.db = nc; q.dc = 1;
        q.accumulate = accumulate; q.nsplit = nsplit; q.split_stride = sstride; q.round_tf32 = 0;
        return sefd_permute3p(q, sx);
    };

    // ---- ISTFT^T and mask Jacobian -> d(mask) laid out like dec[5].y ----
    const ConvLayer& last = P->dec[NL - 1];
    {
        MaskIstftBwdParams m;
        memset(&m, 0, sizeof(m));
        m.dwav = dwav; m.raw_wav = ws + P->raw_wav; m.spec = ws + P->spec; m.mask = ws + last.y; m.dmask = ws + last.dy;
        m.dreal = dreal; m.dimag = dimag;
        m.mT = 2; m.mF = (long long)(T + 1) * 2; m.mB = (long long)256 * (T + 1) * 2;
        m.m_tshift = 1; m.mode = P->mask_mode; m.B = B; m.L = L; m.T = T;
        SEFD_TRY(sefd_mask_istft_bwd_launch(m, st));
    }

    // ---- decoder backward ----
    for (int j = NL - 1; j >= 0; --j) {
        const ConvLayer& c = P->dec[j];
        const int Ch = P->skip ? c.Cin / 2 : c.Cin;
        const float* in0 = j == 0 ? ws + P->U : ws + P->dec[j - 1].z;
        const float* in1 = ws + P->enc[NL - 1 - j].z;
        float* dY = ws + c.dy;
        const float* dbias = nullptr;
        if (j != NL - 1) SEFD_TRY(bn_bwd(c, T + 1, 1, false));
        // weight gradient
        WgradParams wg;
        memset(&wg, 0, sizeof(wg));
        wg.a[0] = src4(in0, c.Fin, T, Ch, Ch);
        wg.a[1] = P->skip ? src4(in1, c.Fin, T, Ch, Ch) : no_src();
        wg.g = src4(dY, c.Fout, T + 1, c.Cout, c.Cout);
        wg.dW = dWs;
        wg.B = B; wg.J = c.Fin; wg.Tg = T + 1; wg.Fa = c.Fin; wg.Ta = T; wg.Fg = c.Fout;
        wg.a_mul = 1; wg.g_mul = 2; wg.ntaps = 10;
        for (int kf = 0; kf < 5; ++kf)
            for (int kt = 0; kt < 2; ++kt) {
                const int i = kf * 2 + kt;
                wg.a_off[i] = 0; wg.g_off[i] = kf - 2; wg.dt[i] = -kt; wg.wslab[i] = i;
            }
        join();
        SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 10, &nsplit, &sstride, st));
        fork();
        if (j == NL - 1) {        // decoder 5 has no BatchNorm behind it: its bias gradient is the column sum of dY
            SEFD_TRY(sefd_colsum2(dY, 1, 0, (long long)B * c.Fout * (T + 1), c.Cout, c.Cout, red_side, ws + P->dbs, sx));
            dbias = ws + P->dbs;
        }
        SEFD_TRY(fold(c, true, dbias));
        side_done();
        // data gradient: d(in0) and d(skip)
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(dY, c.Fout, T + 1, c.Cout, c.Cout);
        g.a[1] = no_src();
        g.o[0] = dst4(j == 0 ? ws + P->dU : ws + P->dec[j - 1].dz, c.Fin, T, Ch, Ch);
        g.o[1] = P->skip ? dst4(ws + P->enc[NL - 1 - j].dz, c.Fin, T, Ch, Ch) : no_dst();
        g.W = ws + c.Wt; g.Wnk = ws + c.Wf; g.nslabs = 10;
        g.round_out[0] = (j == 0) && sefd_get_engine_internal() == 1;      // dU feeds the projection GEMMs
        g.B = B; g.J = c.Fin; g.Tout = T; g.Fin = c.Fout; g.Tin = T + 1;
        conv_taps_down(g, +1);
        SEFD_TRY(sefd_tapgemm(g, st));
    }

    if (P->rl) {
        // ---- cfg.lstm = 'real': tranform backward, LSTM stack backward, gradient back into the encoder layout ----
        const RealLstmExt& R = *P->rl;
        const int tf = sefd_get_engine_internal() == 1;
        join();
        {   // d h1[t][b][k] = sum_d sum_c dU[b][d][t][c] W_d[k][c]
            TapGemmParams g;
            memset(&g, 0, sizeof(g));
            g.a[0] = src4(ws + P->dU, RL_D, T, RL_C, RL_C);
            g.a[1] = no_src();
            g.o[0].p = ws + R.st.dh[1]; g.o[0].sB = RL_H; g.o[0].sF = 0; g.o[0].sT = (long long)B * RL_H; g.o[0].N = RL_H;
            g.o[1] = no_dst();
            g.W = ws + R.WtrT; g.Wnk = ws + R.Wtrp; g.nslabs = RL_D;
            g.B = B; g.J = 1; g.Tout = T; g.Fin = RL_D; g.Tin = T;
            g.fi_mul = 0; g.fo_mul = 1;
            g.ntaps = RL_D;
            for (int d = 0; d < RL_D; ++d) { g.df[d] = d; g.dt[d] = 0; g.wslab[d] = d; }
            SEFD_TRY(sefd_tapgemm(g, st));
        }
        {   // dW_tr[c * 4 + d][k] = sum dU[b][d][t][c] h1[t][b][k]
            WgradParams wg;
            memset(&wg, 0, sizeof(wg));
            wg.a[0].p = ws + R.st.l[1].h; wg.a[0].sB = RL_H; wg.a[0].sF = 0; wg.a[0].sT = (long long)B * RL_H; wg.a[0].C = RL_H;
            wg.a[1] = no_src();
            wg.g = src4(ws + P->dU, RL_D, T, RL_C, RL_C);
            wg.dW = dWs;
            wg.B = B; wg.J = 1; wg.Tg = T; wg.Fa = 1; wg.Ta = T; wg.Fg = RL_D;
            wg.a_mul = 0; wg.g_mul = 0; wg.ntaps = RL_D;
            for (int d = 0; d < RL_D; ++d) { wg.a_off[d] = 0; wg.g_off[d] = d; wg.dt[d] = 0; wg.wslab[d] = d; }
            SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, RL_D, &nsplit, &sstride, st));
            // dWs[d][k][c] -> grads[c][d][k]
            Permute3Params q;
            q.src = dWs; q.dst = grads + R.w_tr; q.na = RL_C; q.nb = RL_D; q.nc = RL_H;
            q.sa = 1; q.sb = (long long)RL_H * RL_C; q.sc = RL_C;
            q.da = (long long)RL_D * RL_H; q.db = RL_H; q.dc = 1;
            q.accumulate = 0; q.nsplit = nsplit; q.split_stride = sstride; q.round_tf32 = 0;
            SEFD_TRY(sefd_permute3p(q, st));
            for (int d = 0; d < RL_D; ++d)
                SEFD_TRY(sefd_colsum2(ws + P->dU + (size_t)d * T * RL_C, B, (long long)RL_D * T * RL_C, T, RL_C, RL_C, wsd + P->red,
                                      ws + P->dbs + d * RL_C, st));
            SEFD_TRY(sefd_permute3(ws + P->dbs, grads + R.b_tr, 1, RL_C, RL_D, 0, 1, RL_C, 0, st));    // dbs[d][c] -> grads[c * 4 + d]
        }
        SEFD_TRY(stack_backward(R.sc, R.st, ws, ws + R.x_tm, T, tf, nullptr, 0u, ws + R.dx_tm, grads, st));
        sefd_absorb_stale_error();
        real_lstm_gather_kernel<<<148 * 4, 256, 0, st>>>(ws + (P->skip ? P->enc[NL - 1].dz2 : P->enc[NL - 1].dz), ws + R.dx_tm, B, T, 1);
        SEFD_TRY(sefd_check_launch("real_lstm_scatter"));
    }
    // ---- projection backward ----
    const long long nX = (long long)B * T * RNN_H;
    for (int q = 0; q < (P->rl ? 0 : 2); ++q) {
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(ws + P->dU + q * 128, 4, T, 256, 128);
        g.a[1] = no_src();
        g.o[0] = dst4(ws + P->dX + (size_t)q * nX, 1, T, RNN_H, RNN_H);
        g.o[1] = no_dst();
        g.W = ws + P->WtrT + (size_t)q * 4 * 128 * 128;
        g.Wnk = ws + P->Wtrp + (size_t)q * 4 * 128 * 128; g.nslabs = 4;
        g.B = B; g.J = 1; g.Tout = T; g.Fin = 4; g.Tin = T;
        g.fi_mul = 0; g.fo_mul = 1;
        g.ntaps = 4;
        for (int d = 0; d < 4; ++d) { g.df[d] = d; g.dt[d] = 0; g.wslab[d] = d; }
        SEFD_TRY(sefd_tapgemm(g, st));
        // dW_tr[c*4+d][k] = sum dU[b][d][t][q*128+c] * X2[q][b][t][k]
        WgradParams wg;
        memset(&wg, 0, sizeof(wg));
        wg.a[0] = src4(ws + P->X2 + (size_t)q * nX, 1, T, RNN_H, RNN_H);
        wg.a[1] = no_src();
        wg.g = src4(ws + P->dU + q * 128, 4, T, 256, 128);
        wg.dW = dWs;
        wg.B = B; wg.J = 1; wg.Tg = T; wg.Fa = 1; wg.Ta = T; wg.Fg = 4;
        wg.a_mul = 0; wg.g_mul = 0; wg.ntaps = 4;
        for (int d = 0; d < 4; ++d) { wg.a_off[d] = 0; wg.g_off[d] = d; wg.dt[d] = 0; wg.wslab[d] = d; }
        join();
        SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 4, &nsplit, &sstride, st));
        // dWs[d][k][c] -> grads[c][d][k]
        fork();
        SEFD_TRY(unperm(dWs, grads + P->w_tr[q], 128, 4, 128, 1, 128 * 128, 128, 0));
        side_done();
    }
    if (!P->rl) fork();
    // db_tr[q][c*4+d] = sum_{b,t} dU[b][d][t][q*128+c]
    for (int d = 0; d < (P->rl ? 0 : 4); ++d)
        SEFD_TRY(sefd_colsum2(ws + P->dU + (size_t)d * T * 256, B, (long long)4 * T * 256, T, 256, 256,
                              red_side, ws + P->dbs + d * 256, sx));
    for (int q = 0; q < (P->rl ? 0 : 2); ++q)   // dbs[d][q*128+c] -> grads[c*4+d]
        SEFD_TRY(sefd_permute3(ws + P->dbs + q * 128, grads + P->b_tr[q], 1, 128, 4, 0, 1, 256, 0, sx));
    if (!P->rl) side_done();

    // ---- LSTM backward (layer 1 then layer 0) ----
    for (int l = P->rl ? -1 : 1; l >= 0; --l) {
        join();                                            // side kernels of the previous layer still read dG
        SEFD_TRY(sefd_clstm_combine_bwd(ws + P->dX, ws + P->dH, nX, st));
        LstmBwdParams lb;
        memset(&lb, 0, sizeof(lb));
        lb.Whh = ws + P->Whh[l]; lb.G = ws + P->Gt[l]; lb.Cc = ws + P->Cc[l]; lb.dH = ws + P->dH; lb.dG = ws + P->dG;
        lb.rows = 2 * B; lb.T = T; lb.round_tf32 = sefd_get_engine_internal() == 1;
        SEFD_TRY(sefd_lstm_bwd_launch(lb, st));
        const size_t lstm_sz = (size_t)2 * B * T * G4;     // one LSTM's dG
        // data gradient into the layer input
        for (int q = 0; q < 2; ++q) {
            TapGemmParams g;
            memset(&g, 0, sizeof(g));
            g.a[0] = src4(ws + P->dG + (size_t)q * B * T * G4, 1, T, G4, G4);
            g.a[1] = src4(ws + P->dG + lstm_sz + (size_t)q * B * T * G4, 1, T, G4, G4);
            g.o[1] = no_dst();
            g.B = B; g.Tout = T; g.Fin = 1; g.Tin = T;
            g.fi_mul = 0; g.fo_mul = 1; g.ntaps = 1;
            if (l == 1) {
                g.o[0] = dst4(ws + P->dX + (size_t)q * nX, 1, T, RNN_H, RNN_H);
                g.W = ws + P->Wih1T; g.J = 1;
                g.Wnk = ws + P->Wih1Q; g.nslabs = 1;
            } else {
                g.o[0] = dst4(ws + (P->skip ? P->enc[NL - 1].dz2 : P->enc[NL - 1].dz) + q * 128, 4, T, 256, 128);   // summed with the skip gradient by BN backward
                g.W = ws + P->Wih0T; g.wJ = (long long)2 * G4 * 128; g.J = 4;
                g.Wnk = ws + P->Wih0Q; g.nslabs = 4; g.wJ_slabs = 1;
            }
            SEFD_TRY(sefd_tapgemm(g, st));
        }
        // weight gradients
        for (int p = 0; p < 2; ++p) {
            const float* dGp = ws + P->dG + (size_t)p * lstm_sz;
            // W_hh: dW[n][k] = sum_{rows, t>=1} dG[t][n] h[t-1][k]
            WgradParams wg;
            memset(&wg, 0, sizeof(wg));
            wg.a[0] = src4(ws + P->Hh[l] + (size_t)p * 2 * B * T * RNN_H, 1, T, RNN_H, RNN_H);
            wg.a[1] = no_src();
            wg.g = src4(dGp, 1, T, G4, G4);
            wg.dW = dWs;
            wg.B = 2 * B; wg.J = 1; wg.Tg = T; wg.Fa = 1; wg.Ta = T; wg.Fg = 1;
            wg.ntaps = 1; wg.dt[0] = -1;
            join();
            SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 1, &nsplit, &sstride, st));
            fork();
            SEFD_TRY(unperm(dWs, grads + P->w_hh[l][p], G4, 1, 128, 1, 0, G4, 0));   // [k][n] -> [n][k]
            side_done();
            // W_ih
            if (l == 1) {
                wg.a[0] = src4(ws + P->X1, 1, T, RNN_H, RNN_H);      // [q][B] rows are contiguous = 2B rows
                wg.dt[0] = 0;
                join();
                SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 1, &nsplit, &sstride, st));
                fork();
                SEFD_TRY(unperm(dWs, grads + P->w_ih[1][p], G4, 1, 128, 1, 0, G4, 0));
                side_done();
            } else {
                for (int q = 0; q < 2; ++q) {
                    WgradParams w0;
                    memset(&w0, 0, sizeof(w0));
                    w0.a[0] = src4(ws + P->enc[NL - 1].z + q * 128, 4, T, 256, 128);
                    w0.a[1] = no_src();
                    w0.g = src4(dGp + (size_t)q * B * T * G4, 1, T, G4, G4);
                    w0.dW = dWs;
                    w0.B = B; w0.J = 1; w0.Tg = T; w0.Fa = 4; w0.Ta = T; w0.Fg = 1;
                    w0.a_mul = 0; w0.g_mul = 0; w0.ntaps = 4;
                    for (int d = 0; d < 4; ++d) { w0.a_off[d] = d; w0.g_off[d] = 0; w0.dt[d] = 0; w0.wslab[d] = d; }
                    join();
                    SEFD_TRY(sefd_wgrad(w0, dWs, (long long)P->dWs_floats, 4, &nsplit, &sstride, st));
                    // dWs[d][c][n] -> grads[n][c][d]   (the second part accumulates onto the first)
                    fork();
                    SEFD_TRY(unperm(dWs, grads + P->w_ih[0][p], G4, 128, 4, 1, G4, (long long)128 * G4, q));
                    side_done();
                }
            }
            // biases: both get sum over rows and time of dG
            fork();
            SEFD_TRY(sefd_colsum2(dGp, 1, 0, (long long)2 * B * T, G4, G4, red_side, grads + P->b_ih[l][p], sx));
            SEFD_TRY(sefd_permute3(grads + P->b_ih[l][p], grads + P->b_hh[l][p], 1, 1, G4, 0, 0, 1, 0, sx));
            side_done();
        }
    }

    // every decoder / projection / LSTM gradient (the tail of the flat buffer, sefd_dccrn_grad_split) is final once the side
    // stream has drained up to here: their all-reduce can run beside the encoder backward
    if (tail_ready) {
        fork();
        cudaEventRecord(tail_ready, sx);
    }

    // ---- encoder backward ----
    for (int i = NL - 1; i >= 0; --i) {
        const ConvLayer& c = P->enc[i];
        float* dY = ws + c.dy;
        SEFD_TRY(bn_bwd(c, T, 0, P->skip != 0));     // without skip connections the only gradient is in dz
        WgradParams wg;
        memset(&wg, 0, sizeof(wg));
        if (i == 0) {
            wg.a[0].p = ws + P->spec + (size_t)T * 2;
            wg.a[0].sT = 2; wg.a[0].sF = (long long)T * 2; wg.a[0].sB = (long long)NBIN * T * 2; wg.a[0].C = 2;
        } else {
            wg.a[0] = src4(ws + P->enc[i - 1].z, c.Fin, T, c.Cin, c.Cin);
        }
        wg.a[1] = no_src();
        wg.g = src4(dY, c.Fout, T, c.Cout, c.Cout);
        wg.dW = dWs;
        wg.B = B; wg.J = c.Fout; wg.Tg = T; wg.Fa = c.Fin; wg.Ta = T; wg.Fg = c.Fout;
        wg.a_mul = 2; wg.g_mul = 1; wg.ntaps = 10;
        for (int kf = 0; kf < 5; ++kf)
            for (int kt = 0; kt < 2; ++kt) {
                const int k = kf * 2 + kt;
                wg.a_off[k] = kf - 2; wg.g_off[k] = 0; wg.dt[k] = kt - 1; wg.wslab[k] = k;
            }
        join();
        SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 10, &nsplit, &sstride, st));
        fork();
        SEFD_TRY(fold(c, false, nullptr));
        side_done();
        if (i > 0) {
            {
                TapGemmParams g;
                memset(&g, 0, sizeof(g));
                g.a[0] = src4(dY, c.Fout, T, c.Cout, c.Cout);
                g.a[1] = no_src();
                g.o[0] = dst4(ws + (P->skip ? P->enc[i - 1].dz2 : P->enc[i - 1].dz), c.Fin, T, c.Cin, c.Cin);   // summed with the skip gradient by BN backward
                g.o[1] = no_dst();
                g.W = ws + c.Wt; g.Wnk = ws + c.Wf; g.nslabs = 10;
                g.B = B; g.J = c.Fout; g.Tout = T; g.Fin = c.Fout; g.Tin = T;
                SEFD_TRY(sefd_tapgemm_up(g, 1, st));   // both output-row phases (one fused launch on tcgen05)
            }
        }
    }
    join();                                                // the caller's stream sees every gradient
    return 0;
}

// ------------------------------------------------------------------------------------------------
// plan accessors of the C ABI (they need the struct definition, so they live here)
// ------------------------------------------------------------------------------------------------
#include "../../include/sefd.h"

extern "C" {

void sefd_dccrn_plan_destroy(sefd_plan* plan) {
    if (plan && plan->side) {
        cudaStreamSynchronize(plan->side);
        cudaStreamDestroy(plan->side);
        cudaEventDestroy(plan->ev_fork);
        cudaEventDestroy(plan->ev_join);
    }
    if (plan && plan->fsn) sefd_fsn_plan_free_ext(plan);
    if (plan && plan->rl) delete plan->rl;
    delete plan;
}
size_t sefd_dccrn_workspace_bytes(const sefd_plan* plan) { return plan ? plan->ws_bytes : 0; }
long long sefd_dccrn_grad_split(const sefd_plan* plan) { return plan && plan->kind == 0 ? plan->dec[0].wr : 0; }
long long sefd_dccrn_param_floats(const sefd_plan* plan) { return plan ? plan->n_param_floats : 0; }
long long sefd_dccrn_buffer_floats(const sefd_plan* plan) { return plan ? plan->n_buffer_floats : 0; }
int sefd_dccrn_num_params(const sefd_plan* plan) { return plan ? (int)plan->params.size() : 0; }
int sefd_dccrn_num_buffers(const sefd_plan* plan) { return plan ? (int)plan->buffers.size() : 0; }

int sefd_dccrn_entry_info(const sefd_plan* plan, int kind, int idx, char* name, int name_cap, long long* offset,
                          long long* numel, int* ndim, long long shape[4]) {
    SEFD_REQUIRE(plan != nullptr, "entry_info: null plan");
    const std::vector<ParamInfo>& v = kind == 0 ? plan->params : plan->buffers;
    SEFD_REQUIRE(idx >= 0 && idx < (int)v.size(), "entry_info: index %d out of range", idx);
    const ParamInfo& pi = v[idx];
    SEFD_REQUIRE((int)pi.name.size() < name_cap, "entry_info: name buffer too small");
    strcpy(name, pi.name.c_str());
    *offset = pi.offset;
    *numel = pi.numel;
    *ndim = pi.ndim;
    for (int i = 0; i < 4; ++i) shape[i] = pi.shape[i];
    return 0;
}

int sefd_dccrn_tensor_info(const sefd_plan* P, const char* name, long long* off, int* ndim, long long shape[4]) {
    SEFD_REQUIRE(P != nullptr && name != nullptr, "tensor_info: null argument");
    if (P->kind == 1) return sefd_crn_tensor_info(P, name, off, ndim, shape);
    if (P->kind == 2) return sefd_fsn_tensor_info(P, name, off, ndim, shape);
    const long long B = P->B, T = P->T;
    auto set = [&](size_t o, long long a, long long b, long long c, long long d) {
        *off = (long long)o; *ndim = 4; shape[0] = a; shape[1] = b; shape[2] = c; shape[3] = d;
        return 0;
    };
    const std::string n(name);
    if (n == "spec") return set(P->spec, B, NBIN, T, 2);
    if (n == "raw_wav") return set(P->raw_wav, 1, 1, B, P->L);
    if (n == "X1") return set(P->X1, 2, B, T, RNN_H);
    if (n == "X2") return set(P->X2, 2, B, T, RNN_H);
    if (n == "dX") return set(P->dX, 2, B, T, RNN_H);
    if (n == "U") return set(P->U, B, 4, T, 256);
    if (n == "dU") return set(P->dU, B, 4, T, 256);
    if (n == "dH") return set(P->dH, 2, 2 * B, T, RNN_H);
    if (n == "dG") return set(P->dG, 2, 2 * B, T, G4);
    if (n.size() >= 6 && (n.compare(0, 3, "enc") == 0 || n.compare(0, 3, "dec") == 0)) {
        const bool dec = n[0] == 'd';
        const int i = n[3] - '0';
        SEFD_REQUIRE(i >= 0 && i < NL && n[4] == '.', "tensor_info: bad name %s", name);
        const ConvLayer& c = dec ? P->dec[i] : P->enc[i];
        const std::string f = n.substr(5);
        if (f == "y") return set(c.y, B, c.Fout, dec ? T + 1 : T, c.Cout);
        if (f == "z" && !(dec && i == NL - 1)) return set(c.z, B, c.Fout, T, c.Cout);
        if (f == "dz" && !(dec && i == NL - 1)) return set(c.dz, B, c.Fout, T, c.Cout);
        if (f == "dz2" && !dec) return set(c.dz2, B, c.Fout, T, c.Cout);
        if (f == "dy") return set(c.dy, B, c.Fout, dec ? T + 1 : T, c.Cout);
    }
    if (n.size() == 7 && n.compare(0, 4, "lstm") == 0 && n[5] == '.') {
        const int l = n[4] - '0';
        SEFD_REQUIRE(l == 0 || l == 1, "tensor_info: bad name %s", name);
        if (n[6] == 'G') return set(P->Gt[l], 2, 2 * B, T, G4);
        if (n[6] == 'H') return set(P->Hh[l], 2, 2 * B, T, RNN_H);
        if (n[6] == 'C') return set(P->Cc[l], 2, 2 * B, T, RNN_H);
    }
    sefd_set_error("tensor_info: unknown tensor '%s'", name);
    return -1;
}

int sefd_dccrn_loss(const sefd_plan* P, const float* out_wav, const float* target, int kind, int reuse_dots,
                    float* loss, float* coef, void* ws, void* stream) {
    SEFD_REQUIRE(P && out_wav && target && loss && coef && ws, "dccrn_loss: null argument");
    double* wsd = (double*)ws;
    return sefd_loss_fwd_launch(out_wav, target, P->B, P->L, kind, wsd + P->dots, reuse_dots, loss, coef,
                                (cudaStream_t)stream);
}

}  // extern "C"
