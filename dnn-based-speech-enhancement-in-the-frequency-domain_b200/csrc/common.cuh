// Shared helpers for the sefd CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SEFD_MAX_TAPS 12

// ---- error plumbing (never throws across the C ABI) ---------------------------------
void sefd_set_error(const char* fmt, ...);
int sefd_check_launch(const char* what);   // returns 0 or negative code, records message

#define SEFD_REQUIRE(cond, ...)                       \
    do {                                              \
        if (!(cond)) {                                \
            sefd_set_error(__VA_ARGS__);              \
            return -1;                                \
        }                                             \
    } while (0)

#define SEFD_TRY(expr)                 \
    do {                               \
        int _rc = (expr);              \
        if (_rc != 0) return _rc;      \
    } while (0)

// ---- tensor views (strides in floats) -----------------------------------------------
// Activations on the path are channels-last: element (b, f, t, c) at p[b*sB + f*sF + t*sT + c].
struct TapSrc {
    const float* p;
    long long sB, sF, sT;
    int C;
};
struct TapDst {
    float* p;
    long long sB, sF, sT;
    int N;
};

// out[b, j*fo_mul+fo_off, t, n] (+)= bias[n] + sum_tap sum_k A[b, j*fi_mul+df[tap], t+dt[tap], k] * W[wslab[tap]][k][n]
// with A = concat_k(a[0], a[1]); n split over o[0], o[1]; source coords outside [0,Fin)x[0,Tin) read as 0.
struct TapGemmParams {
    TapSrc a[2];
    TapDst o[2];
    const float* W;      // [slab][K][N], n contiguous (operand layout of the fp32 CUDA-core engine)
    long long wJ;        // extra weight offset per j (weights that vary with the output row)
    const float* Wnk;    // same weights as [slab][N][K], k contiguous (operand layout of the tensor-core engine) or null
    int nslabs;          // slabs addressable from Wnk
    int wJ_slabs;        // slab offset per j (the tensor-core engine's equivalent of wJ)
    long long w_ldk;     // row pitch of Wnk in floats (0: K)
    long long w_slab_stride;   // slab pitch of Wnk in floats (0: K*N)
    const float* bias;   // [N] or nullptr
    long long bJ;        // extra bias offset per j
    double* stats;       // [2][N] (sum, sum of squares over all written outputs) or nullptr
    int B, J, Tout;
    int Fin, Tin;
    int fi_mul, fo_mul, fo_off;
    int ntaps;
    int df[SEFD_MAX_TAPS], dt[SEFD_MAX_TAPS], wslab[SEFD_MAX_TAPS];
    int accum[2];
    int round_out[2];    // round the stored values to tf32 (the destination feeds a tensor-core GEMM)
    // tensor-core engine only: nacc = 2 computes TWO output rows per tile (row j*fo_mul + fo_off + tap_acc[tap]) from
    // one set of activation tiles - the two output-row phases of a stride-2 "up" conv share their input rows
    int nacc;            // 0 / 1: one output row per tile
    int tap_acc[SEFD_MAX_TAPS];
};

// dW[wslab[tap]][k][n] += sum_{b,j,t} A[b, j*a_mul+a_off[tap], t+dt[tap], k] * G[b, j*g_mul+g_off[tap], t, n]
struct WgradParams {
    TapSrc a[2];
    TapSrc g;            // C field = N
    float* dW;           // [slab][K][N]; accumulated with atomics (caller zeroes)
    int B, J, Tg;        // t runs over [0,Tg)
    int Fa, Ta, Fg;
    int a_mul, g_mul;
    int ntaps;
    int a_off[SEFD_MAX_TAPS], g_off[SEFD_MAX_TAPS], dt[SEFD_MAX_TAPS], wslab[SEFD_MAX_TAPS];
    int rows_per_cta;    // (b,j) rows handled by one CTA
    // g_tiled = 1: G is TILE-MAJOR [Fg][Tg / 128 tiles][C / 32][128][32] (lstm_step_tc.cu: 128-position tiles whose
    // 32-channel blocks are contiguous 16 KB TMA boxes) instead of channels-last; tensor-core engine only, B = 1
    int g_tiled;
};

// > 0: the tensor-core weight gradient uses at most n CTAs (and sizes its row splits for them), leaving SMs free for a kernel
// that runs beside it on another stream; 0 restores the whole device
void sefd_wgrad_tc_set_cta_limit(int n);
int sefd_tapgemm_simt(const TapGemmParams& p, cudaStream_t st);
bool sefd_tapgemm_tc_eligible(const TapGemmParams& p);
int sefd_tapgemm_tc(const TapGemmParams& p, cudaStream_t st);
// dispatch: tensor-core engine when selected (default) and the problem is eligible, else the fp32 engine
int sefd_tapgemm(const TapGemmParams& p, cudaStream_t st);
// stride-2 "up" conv (convT forward: mode 0, conv data gradient: mode 1): both output-row phases.  `g` carries the
// sources / destinations / weights / B, J, Tout, Fin, Tin; taps and row mapping are filled in here.  One fused launch
// on the tensor-core engine when eligible (N <= 128), else one launch per phase.
int sefd_tapgemm_up(const TapGemmParams& g, int mode, cudaStream_t st);
int sefd_get_engine_internal();
// dedicated kernels for the 2-channel ends of the network (K = 2 or N = 2 per tap)
bool sefd_skinny_conv_eligible(const TapGemmParams& p);
int sefd_skinny_conv(const TapGemmParams& p, cudaStream_t st);
bool sefd_skinny_up_n2_eligible(int Ch, int Cout);
int sefd_skinny_up_n2(const float* x0, const float* x1, const float* W, const float* bias, float* y, int B, int F, int T,
                      cudaStream_t st);
bool sefd_skinny_wgrad_eligible(const WgradParams& p);
int sefd_skinny_wgrad(const WgradParams& p, cudaStream_t st);
int sefd_wgrad_simt(const WgradParams& p, cudaStream_t st);
// dispatch (tensor-core engine when eligible): writes *nsplit partial gradients [nslabs][K][N], *split_stride
// floats apart, into `partial`; the caller's fold step sums them.
int sefd_wgrad(const WgradParams& w, float* partial, long long cap_floats, int nslabs, int* nsplit,
               long long* split_stride, cudaStream_t st);

// ---- small device helpers ------------------------------------------------------------
// round-to-nearest fp32 -> tf32 (10-bit mantissa).  The tcgen05 tf32 datapath ignores the low 13 mantissa bits
// of its operands (truncation, biased); producers of tensor-core operands round here instead, for free.
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
