// Plan data structures shared by the model orchestrations (dccrn.cu: complex DCCRN, crn.cu: real CRN):
// parameter layout in the flat buffer, workspace carving, per-layer offsets.
#pragma once
#include <string.h>

#include <initializer_list>
#include <string>
#include <vector>

#include "../../include/sefd.h"
#include "dccrn.cuh"

namespace {

constexpr int NL = 6;            // encoder / decoder depth (config.py:50 dccrn_kernel_num)
constexpr int NBIN = 257, HOP = 100;
constexpr int RNN_H = 128, G4 = 512;
constexpr float BN_EPS = 1e-5f, BN_MOM = 0.1f;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ParamInfo {
    std::string name;
    long long offset, numel;
    int ndim;
    long long shape[4];
};

struct ConvLayer {
    int Cin, Cout;          // real channel counts (Cin includes the skip half for decoders)
    int Fin, Fout;
    long long wr, br, wi, bi, gamma, beta, alpha;   // param offsets (gamma < 0: no BN/PReLU)
    long long rmean, rvar;                           // bn buffer offsets
    // use_cbn (SEFD_PLAN_CBN): gamma / beta hold the offsets of Wrr / Br, rmean / rvar those of RMr / RVrr; the others:
    long long wri, wii, bi2, rmi, rvri, rvii;
    size_t cstats /*doubles: 5 * Cout / 2 moments of the ComplexBatchNorm statistics pass*/;
    // workspace offsets (floats)
    size_t y, z, Wf, Wt, bias, stats /*doubles*/, save, dz, dz2, dy;
};

}  // namespace

struct FsnExt;               // fsnet.cu
struct RealLstmExt;          // dccrn.cu: cfg.lstm = 'real'

struct sefd_plan {
    int kind;                 // 0: DCCRN (complex), 1: CRN (real), 2: FullSubNet (fsnet.cu; everything below except the
                              // parameter list / workspace size lives in `fsn`)
    FsnExt* fsn = nullptr;
    RealLstmExt* rl = nullptr; // non-null: the recurrent part is one 2-layer nn.LSTM(1024 -> 256) + Linear (models.py:96-105)
    int B, L, T, mask_mode;
    int skip = 1;             // 1: decoder convs read complex_cat(out, encoder skip) (cfg.skip_type, models.py:107-169)
    int cbn = 0;              // 1: ComplexBatchNorm instead of BatchNorm2d (use_cbn, models.py:76, 120, 151; cbn.cu)
    size_t cbn_coef = 0;      // floats: [9][128] coefficients of the ComplexBatchNorm backward apply pass
    int ch[NL + 1], Fe[NL + 1];
    ConvLayer enc[NL], dec[NL];
    // LSTM parameter offsets [layer][lstm]
    long long w_ih[2][2], w_hh[2][2], b_ih[2][2], b_hh[2][2], w_tr[2], b_tr[2];
    std::vector<ParamInfo> params, buffers;
    long long n_param_floats, n_buffer_floats;
    // workspace (float offsets unless noted)
    size_t ws_bytes;
    size_t spec, raw_wav, dots /*double*/, stats_all /*double*/, stats_all_n;
    size_t Gt[2], Hh[2], Cc[2], X1, X2, U;
    size_t Wih0p, Wih0T, Wih0Q, Wih1p, Wih1T, Wih1Q, Whh[2], bsum[2], Wtrp, WtrT, btrp;
    size_t dU, dY, dWs, dbs, red /*double*/, dX, dH, dG, dzd[NL];
    size_t dY_floats, dWs_floats;
    // side stream of the backward: gradient folds / un-permutes / bias column sums run beside the GEMM chain
    mutable cudaStream_t side = nullptr;
    mutable cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    size_t red2 = 0;          /*double*/
    // CRN only (crn.cu)
    long long c_wih, c_whh, c_bih, c_bhh, c_wtr, c_btr;       // parameter offsets of enhance / tranform
    size_t mag, tmag, tspec, WihP, WihT;                      // workspace offsets (floats)
};

namespace {

void add_param(sefd_plan* P, const std::string& name, long long& cursor, long long* off, std::initializer_list<long long> shape) {
    ParamInfo pi;
    pi.name = name;
    pi.ndim = (int)shape.size();
    pi.numel = 1;
    int i = 0;
    for (long long s : shape) {
        pi.shape[i++] = s;
        pi.numel *= s;
    }
    for (; i < 4; ++i) pi.shape[i] = 1;
    pi.offset = cursor;
    *off = cursor;
    cursor += (pi.numel + 3) / 4 * 4;
    P->params.push_back(pi);
}

void add_buffer(sefd_plan* P, const std::string& name, long long& cursor, long long* off, long long n) {
    ParamInfo pi;
    pi.name = name;
    pi.ndim = 1;
    pi.numel = n;
    pi.shape[0] = n;
    pi.shape[1] = pi.shape[2] = pi.shape[3] = 1;
    pi.offset = cursor;
    *off = cursor;
    cursor += (n + 3) / 4 * 4;
    P->buffers.push_back(pi);
}

struct Carver {
    size_t cur = 0;   // bytes
    size_t floats(size_t n) {
        cur = align_up(cur, 256);
        size_t o = cur / 4;
        cur += n * 4;
        return o;
    }
    size_t doubles(size_t n) {
        cur = align_up(cur, 256);
        size_t o = cur / 8;
        cur += n * 8;
        return o;
    }
};

}  // namespace

