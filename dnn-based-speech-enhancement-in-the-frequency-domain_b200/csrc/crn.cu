// CRN train-step orchestration (reference: class CRN, models.py:329-565; RealConv2d / RealConvTranspose2d,
// tools_for_model.py:341-425): the real-valued twin of DCCRN.
//   wave -> STFT -> |X| (DC dropped) -> 6 x [Conv2d(5,2)/(2,1) + BN + PReLU] -> nn.LSTM(512 -> 128) -> Linear(128 -> 512)
//        -> 6 x [ConvTranspose2d on cat(out, skip) (+ BN + PReLU)] -> tanh mask x |X| with the noisy phase -> ISTFT -> clamp
// Same kernels as the DCCRN path (tap-GEMM engines, BN+PReLU passes, LSTM recurrence, FFT STFT/ISTFT); only the
// weight packing (real instead of block-complex), the single LSTM and the magnitude mask differ.
// Channels [1,16,32,64,128,128,128] (models.py:362-363: kernel_num // 2), so the 1/16-channel ends run on the fp32
// CUDA-core engine and the 32..256-channel middle on tcgen05.
#include "plan.cuh"
#include "taps.cuh"

namespace {
constexpr int C_H = 128, C_G4 = 512;      // nn.LSTM hidden size (rnn_units // 2, models.py:359) and 4 gates
}

// ------------------------------------------------------------------------------------------------
sefd_plan* sefd_crn_plan_create_impl(int B, int L) {
    if (B <= 0 || L <= 0 || L % HOP != 0) {
        sefd_set_error("crn plan: need B > 0 and L a positive multiple of %d (got B=%d L=%d)", HOP, B, L);
        return nullptr;
    }
    sefd_plan* P = new sefd_plan();
    P->kind = 1;
    P->B = B;
    P->L = L;
    P->T = L / HOP + 3;
    P->mask_mode = SEFD_MASK_MAG;
    const int kn[NL + 1] = {1, 16, 32, 64, 128, 128, 128};
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
        add_param(P, pre + ".0.conv.weight", pc, &c.wr, {c.Cout, c.Cin, 5, 2});
        add_param(P, pre + ".0.conv.bias", pc, &c.br, {c.Cout});
        add_param(P, pre + ".1.weight", pc, &c.gamma, {c.Cout});
        add_param(P, pre + ".1.bias", pc, &c.beta, {c.Cout});
        add_param(P, pre + ".2.weight", pc, &c.alpha, {1});
        add_buffer(P, pre + ".1.running_mean", bc, &c.rmean, c.Cout);
        add_buffer(P, pre + ".1.running_var", bc, &c.rvar, c.Cout);
        c.wi = c.bi = -1;
    }
    for (int j = 0; j < NL; ++j) {
        ConvLayer& c = P->dec[j];
        const int idx = NL - j;
        c.Cin = 2 * kn[idx];                      // cat(out, skip), models.py:494
        c.Cout = kn[idx - 1];
        c.Fin = P->Fe[idx];
        c.Fout = 2 * c.Fin;
        const std::string pre = "decoder." + std::to_string(j);
        add_param(P, pre + ".0.conv.weight", pc, &c.wr, {c.Cin, c.Cout, 5, 2});
        add_param(P, pre + ".0.conv.bias", pc, &c.br, {c.Cout});
        c.wi = c.bi = -1;
        if (j != NL - 1) {
            add_param(P, pre + ".1.weight", pc, &c.gamma, {c.Cout});
            add_param(P, pre + ".1.bias", pc, &c.beta, {c.Cout});
            add_param(P, pre + ".2.weight", pc, &c.alpha, {1});
            add_buffer(P, pre + ".1.running_mean", bc, &c.rmean, c.Cout);
            add_buffer(P, pre + ".1.running_var", bc, &c.rvar, c.Cout);
        } else {
            c.gamma = c.beta = c.alpha = c.rmean = c.rvar = -1;
        }
    }
    add_param(P, "enhance.weight_ih_l0", pc, &P->c_wih, {C_G4, 512});
    add_param(P, "enhance.weight_hh_l0", pc, &P->c_whh, {C_G4, C_H});
    add_param(P, "enhance.bias_ih_l0", pc, &P->c_bih, {C_G4});
    add_param(P, "enhance.bias_hh_l0", pc, &P->c_bhh, {C_G4});
    add_param(P, "tranform.weight", pc, &P->c_wtr, {512, C_H});
    add_param(P, "tranform.bias", pc, &P->c_btr, {512});
    P->n_param_floats = pc;
    P->n_buffer_floats = bc;

    // ---- workspace ----
    const size_t Bz = B, T = P->T;
    Carver w;
    P->spec = w.floats(Bz * NBIN * T * 2);
    P->tspec = w.floats(Bz * NBIN * T * 2);
    P->mag = w.floats(Bz * NBIN * T);
    P->raw_wav = w.floats(Bz * L);
    P->dots = w.doubles(Bz * 8 + 8);
    size_t nstat = 0;
    for (int i = 0; i < NL; ++i) nstat += 2 * P->enc[i].Cout;
    for (int j = 0; j < NL - 1; ++j) nstat += 2 * P->dec[j].Cout;
    P->stats_all = w.doubles(nstat);
    P->stats_all_n = nstat;
    size_t sc = P->stats_all;
    size_t max_w = 0;
    for (int i = 0; i < NL; ++i) {
        ConvLayer& c = P->enc[i];
        const size_t n = Bz * c.Fout * T * c.Cout;
        c.y = w.floats(n);
        c.z = w.floats(n);
        c.dz = w.floats(n);
        c.dz2 = w.floats(n);
        c.dy = w.floats(n);
        c.Wf = w.floats(10ull * c.Cin * c.Cout);
        c.Wt = w.floats(10ull * c.Cin * c.Cout);
        c.bias = 0;
        c.save = w.floats(2 * c.Cout);
        c.stats = sc;
        sc += 2 * c.Cout;
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
        c.bias = 0;
        c.save = w.floats(2 * c.Cout + 4);
        if (j != NL - 1) {
            c.stats = sc;
            sc += 2 * c.Cout;
        }
        if (10ull * c.Cin * c.Cout > max_w) max_w = 10ull * c.Cin * c.Cout;
    }
    P->Gt[0] = w.floats(Bz * T * C_G4);
    P->Hh[0] = w.floats(Bz * T * C_H);
    P->Cc[0] = w.floats(Bz * T * C_H);
    P->bsum[0] = w.floats(C_G4);
    P->U = w.floats(Bz * 4 * T * C_H);
    P->WihP = w.floats(4ull * C_H * C_G4);        // [d][c][n]
    P->WihT = w.floats(4ull * C_G4 * C_H);        // [d][n][c]
    P->Wtrp = w.floats(4ull * C_H * C_H);         // [d][k][c]
    P->WtrT = w.floats(4ull * C_H * C_H);         // [d][c][k]
    P->btrp = w.floats(4ull * C_H);               // [d][c]
    P->dU = w.floats(Bz * 4 * T * C_H);
    if (max_w < 4ull * C_H * C_G4) max_w = 4ull * C_H * C_G4;
    P->dWs_floats = 16 * max_w;
    P->dWs = w.floats(16 * max_w);
    P->dbs = w.floats(1024);
    P->red = w.doubles(2 * 512 + 8);
    P->dH = w.floats(Bz * T * C_H);
    P->dG = w.floats(Bz * T * C_G4);
    P->ws_bytes = align_up(w.cur, 256);
    return P;
}

// ------------------------------------------------------------------------------------------------
static bool tc_layer(int K0, int K1, int N) {
    return sefd_get_engine_internal() == 1 && K0 % 32 == 0 && K1 % 32 == 0 && N % 32 == 0;
}

static int crn_pack_weights(const sefd_plan* P, const float* prm, float* ws, cudaStream_t st) {
    for (int e = 0; e < 2 * NL; ++e) {
        const ConvLayer& c = e < NL ? P->enc[e] : P->dec[e - NL];
        RconvPackParams pp;
        pp.w = prm + c.wr; pp.Ci = c.Cin; pp.Co = c.Cout; pp.transposed = e >= NL;
        pp.Wf = ws + c.Wf; pp.Wt = ws + c.Wt;
        pp.round_tf32 = e < NL ? tc_layer(c.Cin, 0, c.Cout) : tc_layer(c.Cin / 2, c.Cin / 2, c.Cout);
        SEFD_TRY(sefd_pack_rconv(pp, st));
    }
    const int tf = sefd_get_engine_internal() == 1;
    auto perm = [&](const float* src, float* dst, int na, int nb, int nc, long long sa, long long sb, long long sc) -> int {
        Permute3Params q;
        q.src = src; q.dst = dst; q.na = na; q.nb = nb; q.nc = nc; q.sa = sa; q.sb = sb; q.sc = sc;
        q.da = (long long)nb * nc; q.db = nc; q.dc = 1; q.accumulate = 0; q.nsplit = 1; q.split_stride = 0; q.round_tf32 = tf;
        return sefd_permute3p(q, st);
    };
    const float* wih = prm + P->c_wih;            // [n][c*4+d]   (rnn_in feature = channel * 4 + bin, models.py:482)
    SEFD_TRY(perm(wih, ws + P->WihP, 4, C_H, C_G4, 1, 4, 512));      // [d][c][n]
    SEFD_TRY(perm(wih, ws + P->WihT, 4, C_G4, C_H, 1, 512, 4));      // [d][n][c]
    SEFD_TRY(sefd_add2(prm + P->c_bih, prm + P->c_bhh, ws + P->bsum[0], C_G4, st));
    const float* wt = prm + P->c_wtr;             // [c*4+d][k]
    SEFD_TRY(perm(wt, ws + P->Wtrp, 4, C_H, C_H, C_H, 1, 512));      // [d][k][c]
    SEFD_TRY(perm(wt, ws + P->WtrT, 4, C_H, C_H, C_H, 512, 1));      // [d][c][k]
    SEFD_TRY(sefd_permute3(prm + P->c_btr, ws + P->btrp, 1, 4, C_H, 0, 1, 4, 0, st));   // [d][c]
    return 0;
}

// ------------------------------------------------------------------------------------------------
// est_mags / target_mags [B][257][T] (may be null), out_wav [B][L]
int sefd_crn_forward_impl(const sefd_plan* P, const float* prm, float* bnbuf, const float* noisy, const float* target,
                          int train, float* est_mags, float* target_mags, float* out_wav, void* wsv, size_t ws_bytes,
                          cudaStream_t st) {
    SEFD_REQUIRE(P->kind == 1, "crn_forward: not a CRN plan");
    SEFD_REQUIRE(ws_bytes >= P->ws_bytes, "crn_forward: workspace too small (%zu < %zu)", ws_bytes, P->ws_bytes);
    SEFD_REQUIRE(((uintptr_t)wsv & 255) == 0 && ((uintptr_t)prm & 15) == 0, "crn_forward: workspace/params misaligned");
    float* ws = (float*)wsv;
    double* wsd = (double*)wsv;
    const int B = P->B, T = P->T, L = P->L;
    cudaMemsetAsync(wsd + P->stats_all, 0, sizeof(double) * P->stats_all_n, st);
    SEFD_TRY(crn_pack_weights(P, prm, ws, st));
    SEFD_TRY(sefd_stft_launch(noisy, ws + P->spec, B, L, T, st));
    SEFD_TRY(sefd_spec_mag_launch(ws + P->spec, ws + P->mag, (long long)B * NBIN * T, st));
    if (target && target_mags) {                  // models.py:505: target_mags, _ = self.stft(targets)
        SEFD_TRY(sefd_stft_launch(target, ws + P->tspec, B, L, T, st));
        SEFD_TRY(sefd_spec_mag_launch(ws + P->tspec, target_mags, (long long)B * NBIN * T, st));
    }

    auto bn = [&](const ConvLayer& c, int Ty, int tshift) -> int {
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
        if (!train) SEFD_REQUIRE(bnbuf != nullptr, "crn_forward: eval mode needs the BN running statistics");
        return sefd_bn_prelu_fwd(b, st);
    };

    // ---- encoder (models.py:470-473) ----
    for (int i = 0; i < NL; ++i) {
        const ConvLayer& c = P->enc[i];
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        if (i == 0) {
            g.a[0].p = ws + P->mag + (size_t)T;        // DC bin dropped (models.py:466)
            g.a[0].sT = 1; g.a[0].sF = T; g.a[0].sB = (long long)NBIN * T; g.a[0].C = 1;
        } else {
            g.a[0] = src4(ws + P->enc[i - 1].z, c.Fin, T, c.Cin, c.Cin);
        }
        g.a[1] = no_src();
        g.o[0] = dst4(ws + c.y, c.Fout, T, c.Cout, c.Cout);
        g.o[1] = no_dst();
        g.W = ws + c.Wf; g.Wnk = ws + c.Wt; g.nslabs = 10; g.bias = prm + c.br;
        g.stats = train ? wsd + c.stats : nullptr;
        g.B = B; g.J = c.Fout; g.Tout = T; g.Fin = c.Fin; g.Tin = T;
        conv_taps_down(g, -1);
        SEFD_TRY(sefd_tapgemm(g, st));
        SEFD_TRY(bn(c, T, 0));
    }

    // ---- nn.LSTM(512 -> 128) + Linear(128 -> 512) (models.py:475-485) ----
    {
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(ws + P->enc[NL - 1].z, 4, T, C_H, C_H);
        g.a[1] = no_src();
        g.ntaps = 4;
        for (int d = 0; d < 4; ++d) { g.df[d] = d; g.dt[d] = 0; g.wslab[d] = d; }
        g.Fin = 4;
        g.W = ws + P->WihP; g.Wnk = ws + P->WihT; g.nslabs = 4;
        g.o[0] = dst4(ws + P->Gt[0], 1, T, C_G4, C_G4);
        g.o[1] = no_dst();
        g.bias = ws + P->bsum[0];
        g.B = B; g.J = 1; g.Tout = T; g.Tin = T;
        g.fi_mul = 0; g.fo_mul = 1; g.fo_off = 0;
        SEFD_TRY(sefd_tapgemm(g, st));
        LstmFwdParams lp;
        memset(&lp, 0, sizeof(lp));
        lp.Whh = prm + P->c_whh; lp.G = ws + P->Gt[0]; lp.Hh = ws + P->Hh[0]; lp.Cc = ws + P->Cc[0];
        lp.rows = B; lp.T = T; lp.nl = 1;
        SEFD_TRY(sefd_lstm_fwd_launch(lp, st));
    }
    {   // tranform: output feature c*4+d -> U[b][d][t][c]
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(ws + P->Hh[0], 1, T, C_H, C_H);
        g.a[1] = no_src();
        g.o[0] = dst4(ws + P->U, 4, T, C_H, C_H);
        g.o[1] = no_dst();
        g.W = ws + P->Wtrp; g.wJ = (long long)C_H * C_H;
        g.Wnk = ws + P->WtrT; g.nslabs = 4; g.wJ_slabs = 1;
        g.round_out[0] = sefd_get_engine_internal() == 1;
        g.bias = ws + P->btrp; g.bJ = C_H;
        g.B = B; g.J = 4; g.Tout = T; g.Fin = 1; g.Tin = T;
        g.fi_mul = 0; g.fo_mul = 1; g.fo_off = 0;
        g.ntaps = 1;
        SEFD_TRY(sefd_tapgemm(g, st));
    }

    // ---- decoder (models.py:492-496): convT on cat(out, skip), BN over T+1 frames, drop frame 0 ----
    for (int j = 0; j < NL; ++j) {
        const ConvLayer& c = P->dec[j];
        const int Ch = c.Cin / 2;
        const float* in0 = j == 0 ? ws + P->U : ws + P->dec[j - 1].z;
        const float* in1 = ws + P->enc[NL - 1 - j].z;
        {
            TapGemmParams g;
            memset(&g, 0, sizeof(g));
            g.a[0] = src4(in0, c.Fin, T, Ch, Ch);
            g.a[1] = src4(in1, c.Fin, T, Ch, Ch);
            g.o[0] = dst4(ws + c.y, c.Fout, T + 1, c.Cout, c.Cout);
            g.o[1] = no_dst();
            g.W = ws + c.Wf; g.Wnk = ws + c.Wt; g.nslabs = 10; g.bias = prm + c.br;
            g.stats = (train && j != NL - 1) ? wsd + c.stats : nullptr;
            g.B = B; g.J = c.Fin; g.Tout = T + 1; g.Fin = c.Fin; g.Tin = T;
            SEFD_TRY(sefd_tapgemm_up(g, 0, st));   // both output-row phases (one fused launch on tcgen05)
        }
        if (j != NL - 1) SEFD_TRY(bn(c, T + 1, 1));
    }

    // ---- tanh mask x |X|, noisy phase, ISTFT, clamp (models.py:518-532) ----
    const ConvLayer& last = P->dec[NL - 1];
    MaskIstftParams m;
    memset(&m, 0, sizeof(m));
    m.spec = ws + P->spec;
    m.mask = ws + last.y;
    m.mT = 1; m.mF = (long long)(T + 1); m.mB = (long long)256 * (T + 1);
    m.m_tshift = 1;
    m.mode = SEFD_MASK_MAG;
    m.B = B; m.L = L; m.T = T;
    m.out_real = est_mags; m.out_imag = nullptr; m.out_wav = out_wav;
    m.raw_wav = ws + P->raw_wav;
    m.target = target; m.dots = wsd + P->dots;
    return sefd_mask_istft_launch(m, st);
}

// ------------------------------------------------------------------------------------------------
int sefd_crn_backward_impl(const sefd_plan* P, const float* prm, const float* dwav, const float* dmags, float* grads, void* wsv,
                           size_t ws_bytes, cudaStream_t st) {
    SEFD_REQUIRE(P->kind == 1, "crn_backward: not a CRN plan");
    SEFD_REQUIRE(ws_bytes >= P->ws_bytes, "crn_backward: workspace too small");
    float* ws = (float*)wsv;
    double* wsd = (double*)wsv;
    const int B = P->B, T = P->T, L = P->L;
    float* dWs = ws + P->dWs;
    int nsplit = 1;
    long long sstride = 0;

    auto bn_bwd = [&](const ConvLayer& c, int Ty, int tshift, bool two) -> int {
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
    auto fold = [&](const ConvLayer& c, bool dec, const float* dbias) -> int {
        RconvFoldParams f;
        f.dWf = dWs; f.dbias = dbias; f.nsplit = nsplit; f.split_stride = sstride;
        f.Ci = c.Cin; f.Co = c.Cout; f.transposed = dec;
        f.dw = grads + c.wr; f.db = grads + c.br;
        return sefd_fold_rconv(f, st);
    };
    auto unperm = [&](const float* src, float* dst, int na, int nb, int nc, long long sa, long long sb, long long sc) -> int {
        Permute3Params q;
        q.src = src; q.dst = dst; q.na = na; q.nb = nb; q.nc = nc; q.sa = sa; q.sb = sb; q.sc = sc;
        q.da = (long long)nb * nc; q.db = nc; q.dc = 1;
        q.accumulate = 0; q.nsplit = nsplit; q.split_stride = sstride; q.round_tf32 = 0;
        return sefd_permute3p(q, st);
    };

    // ---- ISTFT^T and the mask Jacobian -> d(mask) laid out like dec[5].y ----
    const ConvLayer& last = P->dec[NL - 1];
    {
        MaskIstftBwdParams m;
        memset(&m, 0, sizeof(m));
        m.dwav = dwav; m.raw_wav = ws + P->raw_wav; m.spec = ws + P->spec; m.mask = ws + last.y; m.dmask = ws + last.dy;
        m.dreal = dmags;                 // gradient at est_mags (perceptual LMS branch of CRN.loss, models.py:553-555)
        m.mT = 1; m.mF = (long long)(T + 1); m.mB = (long long)256 * (T + 1);
        m.m_tshift = 1; m.mode = SEFD_MASK_MAG; m.B = B; m.L = L; m.T = T;
        SEFD_TRY(sefd_mask_istft_bwd_launch(m, st));
    }

    // ---- decoder backward ----
    for (int j = NL - 1; j >= 0; --j) {
        const ConvLayer& c = P->dec[j];
        const int Ch = c.Cin / 2;
        const float* in0 = j == 0 ? ws + P->U : ws + P->dec[j - 1].z;
        const float* in1 = ws + P->enc[NL - 1 - j].z;
        float* dY = ws + c.dy;
        const float* dbias = nullptr;
        if (j != NL - 1) {
            SEFD_TRY(bn_bwd(c, T + 1, 1, false));
        } else {
            SEFD_TRY(sefd_colsum2(dY, 1, 0, (long long)B * c.Fout * (T + 1), c.Cout, c.Cout, wsd + P->red, ws + P->dbs, st));
            dbias = ws + P->dbs;
        }
        WgradParams wg;
        memset(&wg, 0, sizeof(wg));
        wg.a[0] = src4(in0, c.Fin, T, Ch, Ch);
        wg.a[1] = src4(in1, c.Fin, T, Ch, Ch);
        wg.g = src4(dY, c.Fout, T + 1, c.Cout, c.Cout);
        wg.dW = dWs;
        wg.B = B; wg.J = c.Fin; wg.Tg = T + 1; wg.Fa = c.Fin; wg.Ta = T; wg.Fg = c.Fout;
        wg.a_mul = 1; wg.g_mul = 2; wg.ntaps = 10;
        for (int kf = 0; kf < 5; ++kf)
            for (int kt = 0; kt < 2; ++kt) {
                const int i = kf * 2 + kt;
                wg.a_off[i] = 0; wg.g_off[i] = kf - 2; wg.dt[i] = -kt; wg.wslab[i] = i;
            }
        SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 10, &nsplit, &sstride, st));
        SEFD_TRY(fold(c, true, dbias));
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(dY, c.Fout, T + 1, c.Cout, c.Cout);
        g.a[1] = no_src();
        g.o[0] = dst4(j == 0 ? ws + P->dU : ws + P->dec[j - 1].dz, c.Fin, T, Ch, Ch);
        g.o[1] = dst4(ws + P->enc[NL - 1 - j].dz, c.Fin, T, Ch, Ch);
        g.W = ws + c.Wt; g.Wnk = ws + c.Wf; g.nslabs = 10;
        g.round_out[0] = (j == 0) && sefd_get_engine_internal() == 1;
        g.B = B; g.J = c.Fin; g.Tout = T; g.Fin = c.Fout; g.Tin = T + 1;
        conv_taps_down(g, +1);
        SEFD_TRY(sefd_tapgemm(g, st));
    }

    // ---- tranform backward: dH[b][t][k] = sum_d sum_c dU[b][d][t][c] W[c*4+d][k] ----
    {
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(ws + P->dU, 4, T, C_H, C_H);
        g.a[1] = no_src();
        g.o[0] = dst4(ws + P->dH, 1, T, C_H, C_H);
        g.o[1] = no_dst();
        g.W = ws + P->WtrT; g.Wnk = ws + P->Wtrp; g.nslabs = 4;
        g.B = B; g.J = 1; g.Tout = T; g.Fin = 4; g.Tin = T;
        g.fi_mul = 0; g.fo_mul = 1;
        g.ntaps = 4;
        for (int d = 0; d < 4; ++d) { g.df[d] = d; g.dt[d] = 0; g.wslab[d] = d; }
        SEFD_TRY(sefd_tapgemm(g, st));
        WgradParams wg;
        memset(&wg, 0, sizeof(wg));
        wg.a[0] = src4(ws + P->Hh[0], 1, T, C_H, C_H);
        wg.a[1] = no_src();
        wg.g = src4(ws + P->dU, 4, T, C_H, C_H);
        wg.dW = dWs;
        wg.B = B; wg.J = 1; wg.Tg = T; wg.Fa = 1; wg.Ta = T; wg.Fg = 4;
        wg.a_mul = 0; wg.g_mul = 0; wg.ntaps = 4;
        for (int d = 0; d < 4; ++d) { wg.a_off[d] = 0; wg.g_off[d] = d; wg.dt[d] = 0; wg.wslab[d] = d; }
        SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 4, &nsplit, &sstride, st));
        SEFD_TRY(unperm(dWs, grads + P->c_wtr, C_H, 4, C_H, 1, (long long)C_H * C_H, C_H));   // dWs[d][k][c] -> [c][d][k]
        for (int d = 0; d < 4; ++d)
            SEFD_TRY(sefd_colsum2(ws + P->dU + (size_t)d * T * C_H, B, (long long)4 * T * C_H, T, C_H, C_H, wsd + P->red,
                                  ws + P->dbs + d * C_H, st));
        SEFD_TRY(sefd_permute3(ws + P->dbs, grads + P->c_btr, 1, C_H, 4, 0, 1, C_H, 0, st));    // dbs[d][c] -> [c*4+d]
    }

    // ---- LSTM backward ----
    {
        LstmBwdParams lb;
        memset(&lb, 0, sizeof(lb));
        lb.Whh = prm + P->c_whh; lb.G = ws + P->Gt[0]; lb.Cc = ws + P->Cc[0]; lb.dH = ws + P->dH; lb.dG = ws + P->dG;
        lb.rows = B; lb.T = T; lb.round_tf32 = sefd_get_engine_internal() == 1; lb.nl = 1;
        SEFD_TRY(sefd_lstm_bwd_launch(lb, st));
        // data gradient into encoder 5 (summed with the skip gradient by its BN backward)
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(ws + P->dG, 1, T, C_G4, C_G4);
        g.a[1] = no_src();
        g.o[0] = dst4(ws + P->enc[NL - 1].dz2, 4, T, C_H, C_H);
        g.o[1] = no_dst();
        g.B = B; g.Tout = T; g.Fin = 1; g.Tin = T;
        g.fi_mul = 0; g.fo_mul = 1; g.ntaps = 1;
        g.W = ws + P->WihT; g.wJ = (long long)C_G4 * C_H; g.J = 4;
        g.Wnk = ws + P->WihP; g.nslabs = 4; g.wJ_slabs = 1;
        SEFD_TRY(sefd_tapgemm(g, st));
        // W_hh: dW[n][k] = sum_{rows, t>=1} dG[t][n] h[t-1][k]
        WgradParams wg;
        memset(&wg, 0, sizeof(wg));
        wg.a[0] = src4(ws + P->Hh[0], 1, T, C_H, C_H);
        wg.a[1] = no_src();
        wg.g = src4(ws + P->dG, 1, T, C_G4, C_G4);
        wg.dW = dWs;
        wg.B = B; wg.J = 1; wg.Tg = T; wg.Fa = 1; wg.Ta = T; wg.Fg = 1;
        wg.ntaps = 1; wg.dt[0] = -1;
        SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 1, &nsplit, &sstride, st));
        SEFD_TRY(unperm(dWs, grads + P->c_whh, C_G4, 1, C_H, 1, 0, C_G4));                    // [k][n] -> [n][k]
        // W_ih: dW[n][c*4+d] = sum dG[b][t][n] enc5.z[b][d][t][c]
        WgradParams w0;
        memset(&w0, 0, sizeof(w0));
        w0.a[0] = src4(ws + P->enc[NL - 1].z, 4, T, C_H, C_H);
        w0.a[1] = no_src();
        w0.g = src4(ws + P->dG, 1, T, C_G4, C_G4);
        w0.dW = dWs;
        w0.B = B; w0.J = 1; w0.Tg = T; w0.Fa = 4; w0.Ta = T; w0.Fg = 1;
        w0.a_mul = 0; w0.g_mul = 0; w0.ntaps = 4;
        for (int d = 0; d < 4; ++d) { w0.a_off[d] = d; w0.g_off[d] = 0; w0.dt[d] = 0; w0.wslab[d] = d; }
        SEFD_TRY(sefd_wgrad(w0, dWs, (long long)P->dWs_floats, 4, &nsplit, &sstride, st));
        SEFD_TRY(unperm(dWs, grads + P->c_wih, C_G4, C_H, 4, 1, C_G4, (long long)C_H * C_G4));   // dWs[d][c][n] -> [n][c][d]
        SEFD_TRY(sefd_colsum2(ws + P->dG, 1, 0, (long long)B * T, C_G4, C_G4, wsd + P->red, grads + P->c_bih, st));
        SEFD_TRY(sefd_permute3(grads + P->c_bih, grads + P->c_bhh, 1, 1, C_G4, 0, 0, 1, 0, st));
    }

    // ---- encoder backward ----
    for (int i = NL - 1; i >= 0; --i) {
        const ConvLayer& c = P->enc[i];
        float* dY = ws + c.dy;
        SEFD_TRY(bn_bwd(c, T, 0, true));
        WgradParams wg;
        memset(&wg, 0, sizeof(wg));
        if (i == 0) {
            wg.a[0].p = ws + P->mag + (size_t)T;
            wg.a[0].sT = 1; wg.a[0].sF = T; wg.a[0].sB = (long long)NBIN * T; wg.a[0].C = 1;
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
        SEFD_TRY(sefd_wgrad(wg, dWs, (long long)P->dWs_floats, 10, &nsplit, &sstride, st));
        SEFD_TRY(fold(c, false, nullptr));
        if (i > 0) {
            {
                TapGemmParams g;
                memset(&g, 0, sizeof(g));
                g.a[0] = src4(dY, c.Fout, T, c.Cout, c.Cout);
                g.a[1] = no_src();
                g.o[0] = dst4(ws + P->enc[i - 1].dz2, c.Fin, T, c.Cin, c.Cin);
                g.o[1] = no_dst();
                g.W = ws + c.Wt; g.Wnk = ws + c.Wf; g.nslabs = 10;
                g.B = B; g.J = c.Fout; g.Tout = T; g.Fin = c.Fout; g.Tin = T;
                SEFD_TRY(sefd_tapgemm_up(g, 1, st));   // both output-row phases (one fused launch on tcgen05)
            }
        }
    }
    return 0;
}

// tensor names of a CRN plan (tests / debugging); called from sefd_dccrn_tensor_info for kind == 1
int sefd_crn_tensor_info(const sefd_plan* P, const char* name, long long* off, int* ndim, long long shape[4]) {
    const long long B = P->B, T = P->T;
    auto set = [&](size_t o, long long a, long long b, long long c, long long d) {
        *off = (long long)o; *ndim = 4; shape[0] = a; shape[1] = b; shape[2] = c; shape[3] = d;
        return 0;
    };
    const std::string n(name);
    if (n == "spec") return set(P->spec, B, NBIN, T, 2);
    if (n == "mag") return set(P->mag, 1, B, NBIN, T);
    if (n == "raw_wav") return set(P->raw_wav, 1, 1, B, P->L);
    if (n == "U") return set(P->U, B, 4, T, C_H);
    if (n == "dU") return set(P->dU, B, 4, T, C_H);
    if (n == "H") return set(P->Hh[0], 1, B, T, C_H);
    if (n == "G") return set(P->Gt[0], 1, B, T, C_G4);
    if (n == "dH") return set(P->dH, 1, B, T, C_H);
    if (n == "dG") return set(P->dG, 1, B, T, C_G4);
    if (n.size() >= 6 && (n.compare(0, 3, "enc") == 0 || n.compare(0, 3, "dec") == 0)) {
        const bool dec = n[0] == 'd';
        const int i = n[3] - '0';
        if (i >= 0 && i < NL && n[4] == '.') {
            const ConvLayer& c = dec ? P->dec[i] : P->enc[i];
            const std::string f = n.substr(5);
            if (f == "y") return set(c.y, B, c.Fout, dec ? T + 1 : T, c.Cout);
            if (f == "dy") return set(c.dy, B, c.Fout, dec ? T + 1 : T, c.Cout);
            if (f == "z" && !(dec && i == NL - 1)) return set(c.z, B, c.Fout, T, c.Cout);
            if (f == "dz" && !(dec && i == NL - 1)) return set(c.dz, B, c.Fout, T, c.Cout);
            if (f == "dz2" && !dec) return set(c.dz2, B, c.Fout, T, c.Cout);
        }
    }
    sefd_set_error("tensor_info: unknown CRN tensor '%s'", name);
    return -1;
}
