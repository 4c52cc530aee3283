// extern "C" surface of libsefd.so (see include/sefd.h) and the error plumbing.
#include <stdarg.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sefd.h"
#include "dccrn.cuh"
#include "fsnet.cuh"
#include "lstm_seq.cuh"
#include "taps.cuh"
#include "prof.cuh"

static thread_local char g_err[1024] = "";

void sefd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;

int sefd_check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        sefd_set_error("%s: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

// plan internals needed here (definition lives in dccrn.cu)
struct PlanView;
extern "C" {

int sefd_abi_version(void) { return 1; }
long long sefd_launch_count(void) { return g_launches; }
const char* sefd_last_error(void) { return g_err; }

static_assert((int)SEFD_MODE_E == (int)SEFD_MASK_E && (int)SEFD_MODE_C == (int)SEFD_MASK_C && (int)SEFD_MODE_R == (int)SEFD_MASK_R, "mode enums");
static_assert((int)SEFD_MSE == (int)SEFD_LOSS_MSE && (int)SEFD_SDR == (int)SEFD_LOSS_SDR && (int)SEFD_SI_SNR == (int)SEFD_LOSS_SISNR &&
                  (int)SEFD_SI_SDR == (int)SEFD_LOSS_SISDR, "loss enums");

#define ST ((cudaStream_t)stream)
#define CHECK_L(L) SEFD_REQUIRE((L) > 0 && (L) % 100 == 0, "L=%d must be a positive multiple of the hop (100)", (L))

int sefd_stft_forward(const float* wav, float* spec, int B, int L, void* stream) {
    CHECK_L(L);
    return sefd_stft_launch(wav, spec, B, L, L / 100 + 3, ST);
}

int sefd_istft_forward(const float* spec, float* wav, int B, int L, void* stream) {
    CHECK_L(L);
    MaskIstftParams m;
    memset(&m, 0, sizeof(m));
    m.spec = spec; m.mode = SEFD_MASK_NONE; m.B = B; m.L = L; m.T = L / 100 + 3;
    m.out_wav = wav;   // note: clamped like the model output; use raw_wav for the unclamped signal
    m.raw_wav = wav;   // same buffer: the raw value is written last, so `wav` ends up unclamped
    return sefd_mask_istft_launch(m, ST);
}

int sefd_istft_backward(const float* dwav, float* dspec, int B, int L, void* stream) {
    CHECK_L(L);
    const int T = L / 100 + 3;
    MaskIstftBwdParams m;
    memset(&m, 0, sizeof(m));
    m.dwav = dwav; m.dmask = dspec; m.mode = SEFD_MASK_NONE; m.B = B; m.L = L; m.T = T;
    m.mT = 2; m.mF = (long long)T * 2; m.mB = (long long)257 * T * 2;
    return sefd_mask_istft_bwd_launch(m, ST);
}

int sefd_mask_istft_forward(const float* spec, const float* mask, int mode, int B, int L, float* out_real,
                            float* out_imag, float* out_wav, float* raw_wav, void* stream) {
    CHECK_L(L);
    SEFD_REQUIRE((mode >= SEFD_MASK_E && mode <= SEFD_MASK_R) || mode == SEFD_MASK_DIRECT, "mask mode %d unsupported", mode);
    const int T = L / 100 + 3;
    MaskIstftParams m;
    memset(&m, 0, sizeof(m));
    m.spec = spec; m.mask = mask; m.mode = mode; m.B = B; m.L = L; m.T = T;
    m.mT = 2; m.mF = (long long)T * 2; m.mB = (long long)256 * T * 2;
    m.out_real = out_real; m.out_imag = out_imag; m.out_wav = out_wav; m.raw_wav = raw_wav;
    return sefd_mask_istft_launch(m, ST);
}

// ---- either transform geometry (config.py:55-61) ------------------------------------------------------------------
static int geometry(int nfft, int L, int* T, int* nbin) {
    SEFD_REQUIRE(nfft == 512 || nfft == 1024, "fft length %d is not built (512: win 400 / hop 100, 1024: win 800 / hop 200)", nfft);
    const int hop = nfft == 512 ? 100 : 200;
    SEFD_REQUIRE(L > 0 && L % hop == 0, "L=%d must be a positive multiple of the hop (%d)", L, hop);
    *T = L / hop + 3;
    *nbin = nfft / 2 + 1;
    return 0;
}

int sefd_stft_forward_n(const float* wav, float* spec, int B, int L, int nfft, void* stream) {
    int T, F;
    SEFD_TRY(geometry(nfft, L, &T, &F));
    return sefd_stft_launch_n(wav, spec, B, L, nfft, ST);
}

int sefd_mask_istft_forward_n(const float* spec, const float* mask, int mode, int B, int L, int nfft, float* out_wav,
                              void* stream) {
    int T, F;
    SEFD_TRY(geometry(nfft, L, &T, &F));
    SEFD_REQUIRE(mode == SEFD_MASK_NONE || (mode >= SEFD_MASK_E && mode <= SEFD_MASK_R) || mode == SEFD_MASK_DIRECT,
                 "mask mode %d unsupported", mode);
    MaskIstftParams m;
    memset(&m, 0, sizeof(m));
    m.spec = spec; m.mask = mask; m.mode = mode; m.B = B; m.L = L; m.T = T;
    m.mT = 2; m.mF = (long long)T * 2; m.mB = (long long)(F - 1) * T * 2;
    m.out_wav = out_wav;
    return sefd_mask_istft_launch_n(m, nullptr, nfft, ST);
}

int sefd_stft_mask_istft_fused(const float* wav, const float* mask, int mode, int B, int L, int nfft, float* out_wav,
                               void* stream) {
    int T, F;
    SEFD_TRY(geometry(nfft, L, &T, &F));
    SEFD_REQUIRE((mode >= SEFD_MASK_E && mode <= SEFD_MASK_R), "fused STFT-mask-ISTFT: mask mode %d unsupported", mode);
    SEFD_REQUIRE(wav != out_wav, "fused STFT-mask-ISTFT: in-place operation is not supported (frames overlap across CTAs)%s", "");
    MaskIstftParams m;
    memset(&m, 0, sizeof(m));
    m.mask = mask; m.mode = mode; m.B = B; m.L = L; m.T = T;
    m.mT = 2; m.mF = (long long)T * 2; m.mB = (long long)(F - 1) * T * 2;
    m.out_wav = out_wav;
    return sefd_mask_istft_launch_n(m, wav, nfft, ST);
}

int sefd_mask_istft_backward(const float* dwav, const float* raw_wav, const float* spec, const float* mask, int mode,
                             int B, int L, float* dmask, void* stream) {
    CHECK_L(L);
    SEFD_REQUIRE((mode >= SEFD_MASK_E && mode <= SEFD_MASK_R) || mode == SEFD_MASK_DIRECT, "mask mode %d unsupported", mode);
    const int T = L / 100 + 3;
    MaskIstftBwdParams m;
    memset(&m, 0, sizeof(m));
    m.dwav = dwav; m.raw_wav = raw_wav; m.spec = spec; m.mask = mask; m.dmask = dmask; m.mode = mode;
    m.B = B; m.L = L; m.T = T;
    m.mT = 2; m.mF = (long long)T * 2; m.mB = (long long)256 * T * 2;
    return sefd_mask_istft_bwd_launch(m, ST);
}

int sefd_loss_forward(const float* est, const float* target, int B, int L, int kind, double* scratch, float* loss,
                      float* coef, void* stream) {
    return sefd_loss_fwd_launch(est, target, B, L, kind, scratch, 0, loss, coef, ST);
}

int sefd_loss_backward(const float* est, const float* target, const float* coef, const float* gout, float* d_est,
                       int B, int L, void* stream) {
    return sefd_loss_bwd_launch(est, target, coef, gout, d_est, B, L, ST);
}

// ---- complex conv ops -------------------------------------------------------------------------
size_t sefd_cconv_workspace_bytes(int Cin, int Cout) {
    // Wf, Wt, dWf (10*Cin*Cout each) + bias / dbias (Cout each, padded) + reduction scratch
    return sizeof(float) * ((2ull + 16) * 10 * Cin * Cout + 2 * 1024) + sizeof(double) * 1024 + 1024;
}

struct CconvWs {
    float *Wf, *Wt, *dW, *bias, *dbias;
    double* red;
};
static CconvWs carve_cconv(void* ws, int Cin, int Cout) {
    CconvWs c;
    const size_t n = 10ull * Cin * Cout;
    float* f = (float*)ws;
    c.Wf = f; c.Wt = f + n; c.dW = f + 2 * n; c.bias = f + 18 * n; c.dbias = c.bias + 1024;
    c.red = (double*)(((uintptr_t)(c.dbias + 1024) + 255) & ~(uintptr_t)255);
    return c;
}
static int pack_op(const float* wr, const float* br, const float* wi, const float* bi, int Cin, int Cout, int transposed,
                   const CconvWs& c, cudaStream_t st, float* zero_bias_src) {
    CconvPackParams pp;
    pp.wr = wr; pp.wi = wi; pp.br = br ? br : zero_bias_src; pp.bi = bi ? bi : zero_bias_src;
    pp.Ci2 = Cin / 2; pp.Co2 = Cout / 2; pp.transposed = transposed; pp.two_src = transposed;
    pp.Wf = c.Wf; pp.Wt = c.Wt; pp.bias = c.bias;
    pp.round_tf32 = sefd_get_engine_internal() == 1 && Cin % 32 == 0 && Cout % 32 == 0;
    return sefd_pack_cconv(pp, st);
}

int sefd_cconv2d_forward(const float* x, const float* wr, const float* br, const float* wi, const float* bi, float* y,
                         int B, int F, int T, int Cin, int Cout, void* ws, void* stream) {
    SEFD_REQUIRE(F % 2 == 0 && Cin % 2 == 0 && Cout % 2 == 0, "cconv2d: F, Cin, Cout must be even");
    CconvWs c = carve_cconv(ws, Cin, Cout);
    cudaMemsetAsync(c.dbias, 0, sizeof(float) * 1024, ST);
    SEFD_TRY(pack_op(wr, br, wi, bi, Cin, Cout, 0, c, ST, c.dbias));
    TapGemmParams g;
    memset(&g, 0, sizeof(g));
    g.a[0] = src4(x, F, T, Cin, Cin);
    g.o[0] = dst4(y, F / 2, T, Cout, Cout);
    g.W = c.Wf; g.Wnk = c.Wt; g.nslabs = 10; g.bias = c.bias;
    g.B = B; g.J = F / 2; g.Tout = T; g.Fin = F; g.Tin = T;
    conv_taps_down(g, -1);
    return sefd_tapgemm(g, ST);
}

int sefd_cconv2d_backward(const float* x, const float* wr, const float* wi, const float* dy, float* dx, float* dwr,
                          float* dbr, float* dwi, float* dbi, int B, int F, int T, int Cin, int Cout, void* ws,
                          void* stream) {
    CconvWs c = carve_cconv(ws, Cin, Cout);
    cudaMemsetAsync(c.dbias, 0, sizeof(float) * 1024, ST);
    SEFD_TRY(pack_op(wr, nullptr, wi, nullptr, Cin, Cout, 0, c, ST, c.dbias));
    WgradParams wg;
    memset(&wg, 0, sizeof(wg));
    wg.a[0] = src4(x, F, T, Cin, Cin);
    wg.g = src4(dy, F / 2, T, Cout, Cout);
    wg.dW = c.dW;
    wg.B = B; wg.J = F / 2; wg.Tg = T; wg.Fa = F; wg.Ta = T; wg.Fg = F / 2;
    wg.a_mul = 2; wg.g_mul = 1; wg.ntaps = 10;
    for (int kf = 0; kf < 5; ++kf)
        for (int kt = 0; kt < 2; ++kt) {
            const int k = kf * 2 + kt;
            wg.a_off[k] = kf - 2; wg.g_off[k] = 0; wg.dt[k] = kt - 1; wg.wslab[k] = k;
        }
    int nsplit = 1;
    long long sstride = 0;
    SEFD_TRY(sefd_wgrad(wg, c.dW, 16ll * 10 * Cin * Cout, 10, &nsplit, &sstride, ST));
    SEFD_TRY(sefd_colsum2(dy, 1, 0, (long long)B * (F / 2) * T, Cout, Cout, c.red, c.dbias, ST));
    CconvFoldParams f;
    f.dWf = c.dW; f.dbias = c.dbias; f.nsplit = nsplit; f.split_stride = sstride; f.Ci2 = Cin / 2; f.Co2 = Cout / 2; f.transposed = 0; f.two_src = 0;
    f.dwr = dwr; f.dwi = dwi; f.dbr = dbr; f.dbi = dbi;
    SEFD_TRY(sefd_fold_cconv(f, ST));
    if (dx) {
        {
            TapGemmParams g;
            memset(&g, 0, sizeof(g));
            g.a[0] = src4(dy, F / 2, T, Cout, Cout);
            g.o[0] = dst4(dx, F, T, Cin, Cin);
            g.W = c.Wt; g.Wnk = c.Wf; g.nslabs = 10;
            g.B = B; g.J = F / 2; g.Tout = T; g.Fin = F / 2; g.Tin = T;
            SEFD_TRY(sefd_tapgemm_up(g, 1, ST));   // both output-row phases (one fused launch on tcgen05)
        }
    }
    return 0;
}

int sefd_cconvT2d_forward(const float* x0, const float* x1, const float* wr, const float* br, const float* wi,
                          const float* bi, float* y, int B, int F, int T, int Cin, int Cout, void* ws, void* stream) {
    SEFD_REQUIRE(Cin % 4 == 0 && Cout % 2 == 0 && x1 != nullptr, "cconvT2d: needs the two complex_cat inputs, Cin %% 4 == 0");
    CconvWs c = carve_cconv(ws, Cin, Cout);
    cudaMemsetAsync(c.dbias, 0, sizeof(float) * 1024, ST);
    SEFD_TRY(pack_op(wr, br, wi, bi, Cin, Cout, 1, c, ST, c.dbias));
    const int Ch = Cin / 2;
    if (sefd_skinny_up_n2_eligible(Ch, Cout)) return sefd_skinny_up_n2(x0, x1, c.Wf, c.bias, y, B, F, T, ST);
    {
        TapGemmParams g;
        memset(&g, 0, sizeof(g));
        g.a[0] = src4(x0, F, T, Ch, Ch);
        g.a[1] = src4(x1, F, T, Ch, Ch);
        g.o[0] = dst4(y, 2 * F, T + 1, Cout, Cout);
        g.W = c.Wf; g.Wnk = c.Wt; g.nslabs = 10; g.bias = c.bias;
        g.B = B; g.J = F; g.Tout = T + 1; g.Fin = F; g.Tin = T;
        SEFD_TRY(sefd_tapgemm_up(g, 0, ST));   // both output-row phases (one fused launch on tcgen05)
    }
    return 0;
}

int sefd_cconvT2d_backward(const float* x0, const float* x1, const float* wr, const float* wi, const float* dy,
                           float* dx0, float* dx1, float* dwr, float* dbr, float* dwi, float* dbi, int B, int F, int T,
                           int Cin, int Cout, void* ws, void* stream) {
    CconvWs c = carve_cconv(ws, Cin, Cout);
    cudaMemsetAsync(c.dbias, 0, sizeof(float) * 1024, ST);
    SEFD_TRY(pack_op(wr, nullptr, wi, nullptr, Cin, Cout, 1, c, ST, c.dbias));
    const int Ch = Cin / 2;
    WgradParams wg;
    memset(&wg, 0, sizeof(wg));
    wg.a[0] = src4(x0, F, T, Ch, Ch);
    wg.a[1] = src4(x1, F, T, Ch, Ch);
    wg.g = src4(dy, 2 * F, T + 1, Cout, Cout);
    wg.dW = c.dW;
    wg.B = B; wg.J = F; wg.Tg = T + 1; wg.Fa = F; wg.Ta = T; wg.Fg = 2 * F;
    wg.a_mul = 1; wg.g_mul = 2; wg.ntaps = 10;
    for (int kf = 0; kf < 5; ++kf)
        for (int kt = 0; kt < 2; ++kt) {
            const int k = kf * 2 + kt;
            wg.a_off[k] = 0; wg.g_off[k] = kf - 2; wg.dt[k] = -kt; wg.wslab[k] = k;
        }
    int nsplit = 1;
    long long sstride = 0;
    SEFD_TRY(sefd_wgrad(wg, c.dW, 16ll * 10 * Cin * Cout, 10, &nsplit, &sstride, ST));
    SEFD_TRY(sefd_colsum2(dy, 1, 0, (long long)B * 2 * F * (T + 1), Cout, Cout, c.red, c.dbias, ST));
    CconvFoldParams f;
    f.dWf = c.dW; f.dbias = c.dbias; f.nsplit = nsplit; f.split_stride = sstride; f.Ci2 = Cin / 2; f.Co2 = Cout / 2; f.transposed = 1; f.two_src = 1;
    f.dwr = dwr; f.dwi = dwi; f.dbr = dbr; f.dbi = dbi;
    SEFD_TRY(sefd_fold_cconv(f, ST));
    TapGemmParams g;
    memset(&g, 0, sizeof(g));
    g.a[0] = src4(dy, 2 * F, T + 1, Cout, Cout);
    g.o[0] = dst4(dx0, F, T, Ch, Ch);
    g.o[1] = dst4(dx1, F, T, Ch, Ch);
    g.W = c.Wt; g.Wnk = c.Wf; g.nslabs = 10;
    g.B = B; g.J = F; g.Tout = T; g.Fin = 2 * F; g.Tin = T + 1;
    conv_taps_down(g, +1);
    return sefd_tapgemm(g, ST);
}

// ---- BN + PReLU -------------------------------------------------------------------------------
__global__ void channel_stats_kernel(const float* __restrict__ y, long long rows, int C, double* stats) {
    const int c = threadIdx.x % C, lane = threadIdx.x / C, lanes = blockDim.x / C;
    double s = 0, s2 = 0;
    if (lane < lanes) {
        for (long long r = (long long)blockIdx.x * lanes + lane; r < rows; r += (long long)gridDim.x * lanes) {
            const double v = y[r * C + c];
            s += v; s2 += v * v;
        }
        atomicAdd(stats + c, s);
        atomicAdd(stats + C + c, s2);
    }
}

int sefd_bn_prelu_forward(const float* y, float* z, long long rows, int C, const float* gamma, const float* beta,
                          const float* alpha, float* save, float* running_mean, float* running_var, double* scratch,
                          void* stream) {
    SEFD_REQUIRE(C >= 4 && C <= 512 && C % 4 == 0, "bn_prelu: C=%d unsupported", C);
    cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, ST);
    sefd_absorb_stale_error();
    channel_stats_kernel<<<148 * 2, C > 256 ? 512 : 256, 0, ST>>>(y, rows, C, scratch);
    SEFD_TRY(sefd_check_launch("channel_stats"));
    BnPreluFwdParams b;
    memset(&b, 0, sizeof(b));
    b.y = y; b.z = z; b.BF = 1; b.Ty = (int)rows; b.T = (int)rows; b.C = C;
    b.stats = scratch; b.n_stat = (double)rows;
    b.gamma = gamma; b.beta = beta; b.alpha = alpha; b.save = save;
    b.running_mean = running_mean; b.running_var = running_var;
    b.momentum = 0.1f; b.eps = 1e-5f;
    return sefd_bn_prelu_fwd(b, ST);
}

int sefd_bn_prelu_backward(const float* y, const float* dz, float* dy, long long rows, int C, const float* gamma,
                           const float* beta, const float* alpha, const float* save, float* dgamma, float* dbeta,
                           float* dalpha, double* scratch, void* stream) {
    BnPreluBwdParams b;
    memset(&b, 0, sizeof(b));
    b.y = y; b.dz = dz; b.dy = dy; b.BF = 1; b.Ty = (int)rows; b.T = (int)rows; b.C = C;
    b.n_stat = (double)rows;
    b.gamma = gamma; b.beta = beta; b.alpha = alpha; b.save = save; b.red = scratch;
    b.dgamma = dgamma; b.dbeta = dbeta; b.dalpha = dalpha;
    return sefd_bn_prelu_bwd(b, ST);
}

// ComplexBatchNorm (+ PReLU with the given slope; a slope of 1 is the bare module) on channels-last rows [rows][C]
int sefd_cbn_prelu_forward(const float* y, float* z, long long rows, int C, const float* w3h, const float* b2h, const float* alpha,
                           float* save, float* running5h, int use_running, double* scratch, void* stream) {
    SEFD_REQUIRE(y && z && w3h && b2h && alpha && save && (use_running ? running5h != nullptr : scratch != nullptr) && rows > 0,
                 "cbn_prelu_forward: bad argument%s", "");
    const int h = C / 2;
    CbnPreluFwdParams b;
    memset(&b, 0, sizeof(b));
    b.y = y; b.z = z; b.BF = 1; b.Ty = (int)rows; b.T = (int)rows; b.C = C;
    b.stats = scratch; b.n_stat = (double)rows;
    for (int i = 0; i < 3; ++i) b.W[i] = w3h + i * h;
    for (int i = 0; i < 2; ++i) b.B2[i] = b2h + i * h;
    b.alpha = alpha; b.save = save;
    if (running5h) {
        b.RM[0] = running5h; b.RM[1] = running5h + h;
        for (int i = 0; i < 3; ++i) b.RV[i] = running5h + (2 + i) * h;
    }
    b.momentum = 0.1f; b.eps = 1e-5f; b.use_running = use_running;
    return sefd_cbn_prelu_fwd(b, ST);
}

int sefd_cbn_prelu_backward(const float* y, const float* dz, float* dy, long long rows, int C, const float* w3h, const float* b2h,
                            const float* alpha, const float* save, float* dw3h, float* db2h, float* dalpha, double* scratch,
                            float* coef, void* stream) {
    SEFD_REQUIRE(y && dz && dy && w3h && b2h && alpha && save && dw3h && db2h && dalpha && scratch && coef && rows > 0,
                 "cbn_prelu_backward: bad argument%s", "");
    const int h = C / 2;
    CbnPreluBwdParams b;
    memset(&b, 0, sizeof(b));
    b.y = y; b.dz = dz; b.dy = dy; b.BF = 1; b.Ty = (int)rows; b.T = (int)rows; b.C = C;
    b.n_stat = (double)rows;
    for (int i = 0; i < 3; ++i) { b.W[i] = w3h + i * h; b.dW[i] = dw3h + i * h; }
    for (int i = 0; i < 2; ++i) { b.B2[i] = b2h + i * h; b.dB2[i] = db2h + i * h; }
    b.alpha = alpha; b.save = save; b.red = scratch; b.coef = coef; b.dalpha = dalpha;
    return sefd_cbn_prelu_bwd(b, ST);
}

int sefd_lstm_forward(const float* w_hh, float* gates, float* h, float* c, int rows, int T, void* stream) {
    LstmFwdParams p;
    memset(&p, 0, sizeof(p));
    p.Whh = w_hh; p.G = gates; p.Hh = h; p.Cc = c; p.rows = rows; p.T = T;
    return sefd_lstm_fwd_launch(p, ST);
}

int sefd_lstm_backward(const float* w_hh, const float* gates, const float* c, const float* dh, float* dgates, int rows,
                       int T, void* stream) {
    LstmBwdParams p;
    memset(&p, 0, sizeof(p));
    p.Whh = w_hh; p.G = gates; p.Cc = c; p.dH = dh; p.dG = dgates; p.rows = rows; p.T = T; p.round_tf32 = 0;
    return sefd_lstm_bwd_launch(p, ST);
}

int sefd_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, int step, float gscale, void* stream) {
    SEFD_REQUIRE(step >= 1, "adam: step counts from 1");
    return sefd_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, gscale, ST);
}

int sefd_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                       float beta1, float beta2, float eps, int* step_dev, float* bc_scratch2, float gscale, void* stream) {
    SEFD_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_dev && bc_scratch2, "adam_step_dev: null argument");
    return sefd_adam_dev(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step_dev, bc_scratch2, gscale, ST);
}

int sefd_axpby(float* y, const float* x, float a, float b, long long n, void* stream) {
    SEFD_REQUIRE(y && x && n >= 0, "axpby: null argument or negative length%s", "");
    return sefd_axpby_launch(y, x, a, b, n, ST);
}

int sefd_counters_inc(long long* const* counters_dev, int n, long long inc, void* stream) {
    SEFD_REQUIRE(counters_dev && n >= 0, "counters_inc: null argument or negative count%s", "");
    return sefd_counters_inc_launch(counters_dev, n, inc, ST);
}

// ---- model level ------------------------------------------------------------------------------
sefd_plan* sefd_dccrn_plan_create(int B, int L, int masking_mode) { return sefd_plan_create_impl(B, L, masking_mode, 0); }
sefd_plan* sefd_dccrn_plan_create_ex(int B, int L, int masking_mode, int flags) {
    if (flags & ~(SEFD_PLAN_NO_SKIP | SEFD_PLAN_REAL_LSTM | SEFD_PLAN_CBN)) {
        sefd_set_error("plan: unknown flag bits 0x%x", flags & ~(SEFD_PLAN_NO_SKIP | SEFD_PLAN_REAL_LSTM | SEFD_PLAN_CBN));
        return nullptr;
    }
    return sefd_plan_create_impl(B, L, masking_mode, flags);
}

int sefd_dccrn_forward(const sefd_plan* plan, const float* params, float* bn_buffers, const float* noisy,
                       const float* target, int train, float* out_real, float* out_imag, float* out_wav, void* ws,
                       size_t ws_bytes, void* stream) {
    SEFD_REQUIRE(plan && params && noisy && out_wav && ws, "dccrn_forward: null argument");
    return sefd_forward_impl(plan, params, bn_buffers, noisy, target, train, out_real, out_imag, out_wav, ws, ws_bytes, ST);
}

int sefd_dccrn_backward(const sefd_plan* plan, const float* params, const float* d_wav, float* grads, void* ws,
                        size_t ws_bytes, void* stream) {
    SEFD_REQUIRE(plan && params && d_wav && grads && ws, "dccrn_backward: null argument");
    return sefd_backward_impl(plan, params, d_wav, nullptr, nullptr, grads, ws, ws_bytes, ST);
}

int sefd_dccrn_backward_overlap(const sefd_plan* plan, const float* params, const float* d_wav, float* grads, void* ws,
                                size_t ws_bytes, void* stream, void* tail_ready_event) {
    SEFD_REQUIRE(plan && params && d_wav && grads && ws, "dccrn_backward_overlap: null argument");
    return sefd_backward_impl(plan, params, d_wav, nullptr, nullptr, grads, ws, ws_bytes, ST, (cudaEvent_t)tail_ready_event);
}

int sefd_dccrn_backward_spec(const sefd_plan* plan, const float* params, const float* d_wav, const float* d_out_real,
                             const float* d_out_imag, float* grads, void* ws, size_t ws_bytes, void* stream) {
    SEFD_REQUIRE(plan && params && grads && ws, "dccrn_backward_spec: null argument");
    SEFD_REQUIRE(d_wav || d_out_real, "dccrn_backward_spec: no gradient given");
    SEFD_REQUIRE((d_out_real == nullptr) == (d_out_imag == nullptr), "dccrn_backward_spec: d_out_real and d_out_imag go together");
    return sefd_backward_impl(plan, params, d_wav, d_out_real, d_out_imag, grads, ws, ws_bytes, ST);
}

// ---- CRN (models.py:329-565) ----------------------------------------------------------------------
sefd_plan* sefd_crn_plan_create(int B, int L) { return sefd_crn_plan_create_impl(B, L); }

int sefd_crn_forward(const sefd_plan* plan, const float* params, float* bn_buffers, const float* noisy, const float* target,
                     int train, float* est_mags, float* target_mags, float* out_wav, void* ws, size_t ws_bytes, void* stream) {
    SEFD_REQUIRE(plan && params && noisy && out_wav && ws, "crn_forward: null argument");
    return sefd_crn_forward_impl(plan, params, bn_buffers, noisy, target, train, est_mags, target_mags, out_wav, ws, ws_bytes, ST);
}

int sefd_crn_backward(const sefd_plan* plan, const float* params, const float* d_wav, float* grads, void* ws, size_t ws_bytes,
                      void* stream) {
    SEFD_REQUIRE(plan && params && d_wav && grads && ws, "crn_backward: null argument");
    return sefd_crn_backward_impl(plan, params, d_wav, nullptr, grads, ws, ws_bytes, ST);
}

int sefd_crn_backward_spec(const sefd_plan* plan, const float* params, const float* d_wav, const float* d_est_mags, float* grads,
                           void* ws, size_t ws_bytes, void* stream) {
    SEFD_REQUIRE(plan && params && grads && ws && (d_wav || d_est_mags), "crn_backward_spec: null argument");
    return sefd_crn_backward_impl(plan, params, d_wav, d_est_mags, grads, ws, ws_bytes, ST);
}

// ---- FullSubNet (models.py:568-682) ------------------------------------------------------------------
sefd_plan* sefd_fsn_plan_create(int B, int frames) { return sefd_fsn_plan_create_impl(B, frames); }

int sefd_fsn_forward(const sefd_plan* plan, const float* params, const float* noisy_mag, int train, float dropout_p,
                     const float* mask_fb, const float* mask_sb, unsigned long long seed, float* crm, void* ws, size_t ws_bytes,
                     void* stream) {
    SEFD_REQUIRE(plan && params && noisy_mag && crm && ws, "fsn_forward: null argument");
    return sefd_fsn_forward_impl(plan, params, noisy_mag, train, dropout_p, mask_fb, mask_sb, seed, crm, ws, ws_bytes, ST);
}

int sefd_dropout_forward(const float* x, float* y, long long n, float p, const float* mask, unsigned long long seed,
                         unsigned int stream_id, void* stream) {
    SEFD_REQUIRE(x && y && n > 0, "dropout_forward: bad argument");
    return sefd_dropout_apply(x, y, n, p, mask, seed, stream_id, 0, ST);
}

int sefd_fsn_backward(const sefd_plan* plan, const float* params, const float* d_crm, float* grads, void* ws, size_t ws_bytes,
                      void* stream) {
    SEFD_REQUIRE(plan && params && d_crm && grads && ws, "fsn_backward: null argument");
    return sefd_fsn_backward_impl(plan, params, d_crm, grads, ws, ws_bytes, ST);
}

}  // extern "C"
