#pragma once
#include "common.cuh"

struct BnPreluFwdParams {
    const float* y;        // [BF][Ty][C] raw conv output
    float* z;              // [BF][T][C]  z[bf,t] = prelu(bn(y[bf, t+tshift]))
    int BF, Ty, T, tshift, C;
    const double* stats;   // [2][C] sum / sum of squares over the n_stat elements of y (train mode)
    double n_stat;
    const float *gamma, *beta, *alpha;
    float* save;           // [2][C] batch mean, inv-std (written by block 0 in train mode)
    float *running_mean, *running_var;   // updated in train mode when non-null; read when use_running
    float momentum, eps;
    int use_running;       // eval mode
    int round_tf32;        // z feeds a tensor-core GEMM: round to tf32 while writing
};

struct BnPreluBwdParams {
    const float* y;        // [BF][Ty][C]
    const float* dz;       // [BF][T][C]
    const float* dz2;      // optional second gradient source summed with dz (skip-connection gradient), or nullptr
    float* dy;             // [BF][Ty][C]
    int BF, Ty, T, tshift, C;
    double n_stat;
    const float *gamma, *beta, *alpha, *save;
    double* red;           // [2C+1] scratch
    float *dgamma, *dbeta, *dalpha;
    int round_tf32;        // dy feeds tensor-core GEMMs
};

// ComplexBatchNorm + PReLU (cbn.cu; use_cbn = True, tools_for_model.py:430-603): C = 2 h channels, real parts in [0, h),
// imaginary parts in [h, 2 h) of every channels-last row
struct CbnPreluFwdParams {
    const float* y;        // [BF][Ty][C] raw conv output
    float* z;              // [BF][T][C]  z[bf, t] = prelu(cbn(y[bf, t + tshift]))
    int BF, Ty, T, tshift, C;
    double* stats;         // [5][h] scratch of the moment pass (train mode)
    double n_stat;
    const float* W[3];     // Wrr, Wri, Wii [h]
    const float* B2[2];    // Br, Bi [h]
    const float* alpha;
    float* save;           // [9][h]: Mr, Mi, Zrr, Zri, Zir, Zii, Vrr + eps, Vri, Vii + eps (kept for the backward)
    float* RM[2];          // RMr, RMi: updated in train mode when non-null; read when use_running
    float* RV[3];          // RVrr, RVri, RVii
    float momentum, eps;
    int use_running;       // eval mode
    int round_tf32;
};

struct CbnPreluBwdParams {
    const float* y;        // [BF][Ty][C]
    const float* dz;       // [BF][T][C]
    const float* dz2;      // optional second gradient source (skip connection) or nullptr
    float* dy;             // [BF][Ty][C]
    int BF, Ty, T, tshift, C;
    double n_stat;
    const float* W[3];
    const float* B2[2];
    const float *alpha, *save;
    double* red;           // [6 h + 1] scratch
    float* coef;           // [9][h] scratch
    float* dW[3];          // d Wrr, d Wri, d Wii
    float* dB2[2];         // d Br, d Bi
    float* dalpha;
    int round_tf32;
};
int sefd_cbn_prelu_fwd(const CbnPreluFwdParams& p, cudaStream_t st);
int sefd_cbn_prelu_bwd(const CbnPreluBwdParams& p, cudaStream_t st);

struct CconvPackParams {
    const float *wr, *wi, *br, *bi;
    int Ci2, Co2;          // complex channel counts (half of the real channel counts)
    int transposed;        // 0: Conv2d weight [Co2][Ci2][5][2]; 1: ConvTranspose2d weight [Ci2][Co2][5][2]
    int two_src;           // K ordering for the skip-concat input
    float* Wf;             // [10][K][N]
    float* Wt;             // [10][N][K]
    float* bias;           // [N]
    int round_tf32;
};

struct CconvFoldParams {
    const float* dWf;      // nsplit x [10][K][N] block-real weight gradient partials, split_stride floats apart
    int nsplit;
    long long split_stride;
    const float* dbias;    // [N] block-real bias gradient or nullptr (-> zeros)
    int Ci2, Co2, transposed, two_src;
    float *dwr, *dwi, *dbr, *dbi;
};

// real Conv2d / ConvTranspose2d weights (RealConv2d / RealConvTranspose2d, tools_for_model.py:341-425) <-> GEMM operands
struct RconvPackParams {
    const float* w;        // Conv2d [Co][Ci][5][2] or ConvTranspose2d [Ci][Co][5][2]
    int Ci, Co, transposed;
    float* Wf;             // [10][K = Ci][N = Co]
    float* Wt;             // [10][N][K]
    int round_tf32;
};
struct RconvFoldParams {
    const float* dWf;      // nsplit x [10][K][N] partials
    int nsplit;
    long long split_stride;
    const float* dbias;    // [Co] or nullptr (-> zeros: the bias sits in front of a BatchNorm)
    int Ci, Co, transposed;
    float *dw, *db;
};
int sefd_pack_rconv(const RconvPackParams& p, cudaStream_t st);
int sefd_fold_rconv(const RconvFoldParams& p, cudaStream_t st);

int sefd_bn_prelu_fwd(const BnPreluFwdParams& p, cudaStream_t st);
int sefd_bn_prelu_bwd(const BnPreluBwdParams& p, cudaStream_t st);
int sefd_pack_cconv(const CconvPackParams& p, cudaStream_t st);
int sefd_fold_cconv(const CconvFoldParams& p, cudaStream_t st);
// dst[a*da + b*db + c*dc] (+)= sum_{s < nsplit} src[s*split_stride + a*sa + b*sb + c*sc]   (optionally tf32-rounded)
struct Permute3Params {
    const float* src;
    float* dst;
    int na, nb, nc;
    long long sa, sb, sc, da, db, dc;
    int accumulate, nsplit;
    long long split_stride;
    int round_tf32;
};
int sefd_permute3p(const Permute3Params& q, cudaStream_t st);
int sefd_permute3(const float* src, float* dst, int na, int nb, int nc, long long sa, long long sb, long long sc,
                  int accumulate, cudaStream_t st);
int sefd_add2(const float* a, const float* b, float* o, long long n, cudaStream_t st);
int sefd_colsum2(const float* x, int nO, long long sO, long long nI, long long sI, int C, double* scratch, float* out,
                 cudaStream_t st);
int sefd_clstm_combine(const float* H, float* X, long long n, int round_tf32, cudaStream_t st);
int sefd_clstm_combine_bwd(const float* dX, float* dH, long long n, cudaStream_t st);
int sefd_adam_dev(float* w, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                  int* step_dev, float* bc_dev, float gscale, cudaStream_t st);
int sefd_adam(float* w, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
              int step, float gscale, cudaStream_t st);

// y = a x + b y (the perceptual step's gradient / loss mixing, trainer.py:166-169)
int sefd_axpby_launch(float* y, const float* x, float a, float b, long long n, cudaStream_t st);
// *ptrs[i] += inc for n device counters (BatchNorm num_batches_tracked of every layer in one launch)
int sefd_counters_inc_launch(long long* const* ptrs, int n, long long inc, cudaStream_t st);
