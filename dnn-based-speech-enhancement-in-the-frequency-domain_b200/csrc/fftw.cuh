// Complex FFT of N = 512 or 1024 points by ONE WARP: N / 32 points per lane in registers, ONE shared-memory exchange
// private to the warp (no CTA barrier), natural-order result left in the warp's buffer.
//
// Why a second FFT next to fft512.cuh (64 threads, three radix-8 passes, two exchanges, seven __syncthreads): the
// STFT / ISTFT kernels of BASELINE configs[4] spent their time waiting at those barriers and on the exchanges (ncu:
// short_scoreboard + barrier the top stalls, 0.15 of HBM).  Here the only synchronisation is __syncwarp, the exchange
// moves each point once, and the same code serves both transform sizes the reference's config allows
// (config.py:55-61: win 400 / hop 100 / fft 512 and win 800 / hop 200 / fft 1024).
//
// Index split  n = 32 n1 + n0  (n0 = lane),  k = k1 + R1 k2,  R1 = N / 32 (16 or 32):
//     W_N^{nk} = W_R1^{n1 k1} * W_N^{n0 k1} * W_32^{n0 k2}
//   step 1: lane n0 : DFT_R1 over n1, times W_N^{n0 k1}                       -> A[k1][n0]  (buffer, row pitch 33)
//   step 2: N = 1024: lane k1 : DFT_32 over n0                                 -> X[k1 + 32 k2]
//           N = 512 : lane (k1, p), n0 = 2 m + p: DFT_16 over m, times W_32^{p q}, then the radix-2 butterfly with the
//                     partner lane p ^ 1 through a shuffle                     -> X[k1 + 16 q + 256 p]
// Row pitch 33 makes both the row-wise writes and the column-wise reads of A conflict-free for 8-byte accesses.
#pragma once
#include "fft512.cuh"

namespace fftw {

constexpr int PITCH = 33;
template <int NFFT> struct Buf { static constexpr int LEN = (NFFT / 32) * PITCH; };     // float2 per warp (>= NFFT)

// e^{-2 pi i q / 32}, q < 16; every use has a compile-time q after unrolling
__device__ __forceinline__ float2 w32(int q) {
    switch (q) {
        case 0: return make_float2(1.f, 0.f);
        case 1: return make_float2(0.98078528040323043f, -0.19509032201612825f);
        case 2: return make_float2(0.92387953251128674f, -0.38268343236508977f);
        case 3: return make_float2(0.83146961230254524f, -0.55557023301960218f);
        case 4: return make_float2(0.70710678118654752f, -0.70710678118654752f);
        case 5: return make_float2(0.55557023301960218f, -0.83146961230254524f);
        case 6: return make_float2(0.38268343236508977f, -0.92387953251128674f);
        case 7: return make_float2(0.19509032201612825f, -0.98078528040323043f);
        case 8: return make_float2(0.f, -1.f);
        case 9: return make_float2(-0.19509032201612825f, -0.98078528040323043f);
        case 10: return make_float2(-0.38268343236508977f, -0.92387953251128674f);
        case 11: return make_float2(-0.55557023301960218f, -0.83146961230254524f);
        case 12: return make_float2(-0.70710678118654752f, -0.70710678118654752f);
        case 13: return make_float2(-0.83146961230254524f, -0.55557023301960218f);
        case 14: return make_float2(-0.92387953251128674f, -0.38268343236508977f);
        default: return make_float2(-0.98078528040323043f, -0.19509032201612825f);
    }
}

// in-place DFT of R register values (R = 8, 16, 32), natural-order output; radix-2 decimation in time on top of dft8
template <int R, bool INV>
struct Dft {
    static __device__ __forceinline__ void run(float2* v) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int m = 0; m < R / 2; ++m) { e[m] = v[2 * m]; o[m] = v[2 * m + 1]; }
        Dft<R / 2, INV>::run(e);
        Dft<R / 2, INV>::run(o);
#pragma unroll
        for (int q = 0; q < R / 2; ++q) {
            const int j = q * (32 / R);
            float2 t;
            if (j == 0) t = o[q];
            else if (j == 8) t = c_rot<INV>(o[q]);
            else t = c_mul(o[q], c_tw<INV>(w32(j)));
            v[q] = c_add(e[q], t);
            v[q + R / 2] = c_sub(e[q], t);
        }
    }
};
template <bool INV>
struct Dft<8, INV> {
    static __device__ __forceinline__ void run(float2* v) { dft8<INV>(v); }
};

// step-1 twiddles of one lane, W_N^{lane k1} = wa[k1 / 4] * wb[k1 % 4]: R1 / 4 + 4 register pairs instead of R1
template <int NFFT>
struct Twiddles {
    static constexpr int R1 = NFFT / 32;
    float2 wa[R1 / 4], wb[4];
    __device__ __forceinline__ void init(int lane) {
#pragma unroll
        for (int a = 0; a < R1 / 4; ++a) {
            float s, c;
            sincospif(2.0f * (float)((4 * a * lane) % NFFT) / NFFT, &s, &c);
            wa[a] = make_float2(c, -s);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float s, c;
            sincospif(2.0f * (float)(b * lane) / NFFT, &s, &c);
            wb[b] = make_float2(c, -s);
        }
    }
    __device__ __forceinline__ float2 get(int k1) const {
        if (k1 < 4) return wb[k1];
        if ((k1 & 3) == 0) return wa[k1 >> 2];
        return c_mul(wa[k1 >> 2], wb[k1 & 3]);
    }
};

// v[n1] = x[32 n1 + lane] on entry; on return the warp's buffer s holds X[k] at s[k], k < NFFT (a __syncwarp has been
// passed, every lane may read any element).  s is private to the warp; lanes must have finished reading an earlier result
// before the call only in the sense of program order (the function starts with __syncwarp).
template <int NFFT, bool INV>
__device__ __forceinline__ void fft_warp(float2 (&v)[NFFT / 32], float2* s, const Twiddles<NFFT>& tw, int lane) {
    constexpr int R1 = NFFT / 32;
    Dft<R1, INV>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < R1; ++k1) v[k1] = c_mul(v[k1], c_tw<INV>(tw.get(k1)));
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) s[k1 * PITCH + lane] = v[k1];
    __syncwarp();
    if (NFFT == 1024) {
#pragma unroll
        for (int n0 = 0; n0 < R1; ++n0) v[n0] = s[lane * PITCH + n0];
        Dft<R1, INV>::run(v);
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < R1; ++k2) s[lane + 32 * k2] = v[k2];
    } else {
        const int k1 = lane & 15, p = lane >> 4;
#pragma unroll
        for (int m = 0; m < R1; ++m) v[m] = s[k1 * PITCH + 2 * m + p];
        Dft<R1, INV>::run(v);
#pragma unroll
        for (int q = 0; q < R1; ++q) {
            if (q > 0) {                                              // times W_32^{p q}
                const float2 w = w32(q);
                v[q] = c_mul(v[q], c_tw<INV>(make_float2(p ? w.x : 1.f, p ? w.y : 0.f)));
            }
            const float ox = __shfl_xor_sync(0xffffffffu, v[q].x, 16), oy = __shfl_xor_sync(0xffffffffu, v[q].y, 16);
            v[q] = p ? make_float2(ox - v[q].x, oy - v[q].y) : make_float2(v[q].x + ox, v[q].y + oy);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < R1; ++q) s[k1 + 16 * q + 256 * p] = v[q];
    }
    __syncwarp();
}

}  // namespace fftw
