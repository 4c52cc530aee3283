// 512-point complex FFT, radix-8 x 3 passes, 64 threads per transform, 8 points per thread in
// registers, two shared-memory exchanges.  Host/device dual so the index algebra is unit-tested on
// the CPU (tests/test_fft_host.py builds csrc/fft512_host_test.cpp with g++).
//
// Index split n = 64 n2 + 8 n1 + n0, k = k0 + 8 k1 + 64 k2:
//   W512^{nk} = W8^{n2 k0} * W64^{n1 k0} * W8^{n1 k1} * W512^{n0 (k0+8k1)} * W8^{n0 k2}
//   pass 1: thread (n1,n0) : DFT8 over n2, times W64^{n1 k0}          -> A[k0][n1][n0]
//   pass 2: thread (k0,n0) : DFT8 over n1, times W512^{n0 (k0+8 k1)}   -> B[k0][k1][n0]
//   pass 3: thread (k0,k1) : DFT8 over n0                              -> X[k0 + 8 k1 + 64 k2]
#pragma once
#ifdef __CUDACC__
#define FFT_HD __host__ __device__ __forceinline__
#else
#define FFT_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

FFT_HD float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FFT_HD float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FFT_HD float2 c_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
FFT_HD float2 c_rot(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
template <bool INV>
FFT_HD float2 c_tw(float2 w) { return INV ? make_float2(w.x, -w.y) : w; }   // table holds e^{-i..}

// in-place 8-point DFT (decimation in frequency), natural-order output
template <bool INV>
FFT_HD void dft8(float2* v) {
    const float r = 0.70710678118654752440f;
    float2 a0 = c_add(v[0], v[4]), a4 = c_sub(v[0], v[4]);
    float2 a1 = c_add(v[1], v[5]), a5 = c_sub(v[1], v[5]);
    float2 a2 = c_add(v[2], v[6]), a6 = c_sub(v[2], v[6]);
    float2 a3 = c_add(v[3], v[7]), a7 = c_sub(v[3], v[7]);
    // twiddles W8^1, W8^2, W8^3 on the odd branch
    a5 = INV ? make_float2(r * (a5.x - a5.y), r * (a5.x + a5.y)) : make_float2(r * (a5.x + a5.y), r * (a5.y - a5.x));
    a6 = c_rot<INV>(a6);
    a7 = INV ? make_float2(-r * (a7.x + a7.y), r * (a7.x - a7.y)) : make_float2(r * (a7.y - a7.x), -r * (a7.x + a7.y));
    // two 4-point DFTs
    {
        float2 c0 = c_add(a0, a2), c2 = c_sub(a0, a2), c1 = c_add(a1, a3), c3 = c_rot<INV>(c_sub(a1, a3));
        v[0] = c_add(c0, c1);
        v[4] = c_sub(c0, c1);
        v[2] = c_add(c2, c3);
        v[6] = c_sub(c2, c3);
    }
    {
        float2 c0 = c_add(a4, a6), c2 = c_sub(a4, a6), c1 = c_add(a5, a7), c3 = c_rot<INV>(c_sub(a5, a7));
        v[1] = c_add(c0, c1);
        v[5] = c_sub(c0, c1);
        v[3] = c_add(c2, c3);
        v[7] = c_sub(c2, c3);
    }
}

// Shared-memory placement of logical element i.  The three exchanges read / write the buffer at strides 1, 8 and 64;
// stored linearly the stride-8 and stride-64 patterns put a half-warp's sixteen 8-byte accesses on 2 - 8 words of the same
// banks (measured on stft_fwd_kernel: 5x the ideal shared wavefronts, short_scoreboard the top stall).  XOR-ing index bits
// 3..6 into bits 0..3 makes all six patterns conflict-free; every access to a transform buffer, inside this file or in a
// caller that packs / unpacks frames, goes through fft_at().
FFT_HD int fft_at(int i) { return i ^ ((i >> 3) & 15); }

// The three passes, each split into "gather + compute" and "scatter" so that a barrier can sit between
// them.  s: 512 complex values (shared memory on the device), element i stored at fft_at(i); tw[j] = exp(-2*pi*i*j/512), j < 512.
template <bool INV>
FFT_HD void fft512_pass1(const float2* s, const float2* tw, int tid, float2* v) {
    for (int n2 = 0; n2 < 8; ++n2) v[n2] = s[fft_at(64 * n2 + tid)];
    dft8<INV>(v);
    const int n1 = tid >> 3;
    for (int k0 = 1; k0 < 8; ++k0) v[k0] = c_mul(v[k0], c_tw<INV>(tw[8 * n1 * k0]));
}
FFT_HD void fft512_scatter1(float2* s, int tid, const float2* v) {
    for (int k0 = 0; k0 < 8; ++k0) s[fft_at(64 * k0 + tid)] = v[k0];      // A[k0][n1][n0], tid = 8 n1 + n0
}
template <bool INV>
FFT_HD void fft512_pass2(const float2* s, const float2* tw, int tid, float2* v) {
    const int k0 = tid >> 3, n0 = tid & 7;
    for (int n1 = 0; n1 < 8; ++n1) v[n1] = s[fft_at(64 * k0 + 8 * n1 + n0)];
    dft8<INV>(v);
    for (int k1 = 0; k1 < 8; ++k1) v[k1] = c_mul(v[k1], c_tw<INV>(tw[n0 * (k0 + 8 * k1)]));
}
FFT_HD void fft512_scatter2(float2* s, int tid, const float2* v) {
    const int k0 = tid >> 3, n0 = tid & 7;
    for (int k1 = 0; k1 < 8; ++k1) s[fft_at(64 * k0 + 8 * k1 + n0)] = v[k1];   // B[k0][k1][n0]
}
template <bool INV>
FFT_HD void fft512_pass3(const float2* s, int tid, float2* v) {
    for (int n0 = 0; n0 < 8; ++n0) v[n0] = s[fft_at(8 * tid + n0)];             // tid = 8 k0 + k1
    dft8<INV>(v);
}
FFT_HD void fft512_scatter3(float2* s, int tid, const float2* v) {
    const int k0 = tid >> 3, k1 = tid & 7;
    for (int k2 = 0; k2 < 8; ++k2) s[fft_at(k0 + 8 * k1 + 64 * k2)] = v[k2];
}

#ifdef __CUDACC__
// Device driver: all threads of the CTA must call this together (it uses __syncthreads);
// `tid` is the thread's index inside its 64-thread transform group, `s` that group's buffer.
template <bool INV>
__device__ __forceinline__ void fft512_cta(float2* s, const float2* tw, int tid) {
    float2 v[8];
    fft512_pass1<INV>(s, tw, tid, v);
    __syncthreads();
    fft512_scatter1(s, tid, v);
    __syncthreads();
    fft512_pass2<INV>(s, tw, tid, v);
    __syncthreads();
    fft512_scatter2(s, tid, v);
    __syncthreads();
    fft512_pass3<INV>(s, tid, v);
    __syncthreads();
    fft512_scatter3(s, tid, v);
    __syncthreads();
}
#endif
