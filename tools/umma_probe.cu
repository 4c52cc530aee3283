// Standalone probe of tcgen05.mma kind::tf32 operand descriptors (K-major vs MN-major, LBO/SBO roles).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
// For each variant: TMA-load A (M=128 x K=32) and B (N=64 x K=32) tiles the way the kernels do, issue 4 MMAs
// (K=8 each), read the 128x64 accumulator back and compare with the exact product on the host.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <stdarg.h>
#include "../dnn-based-speech-enhancement-in-the-frequency-domain_b200/csrc/tc_common.cuh"
void sefd_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); }
int sefd_check_launch(const char*) { return 0; }

struct Variant {
    int mn_major;           // 0: K-major tiles, 1: MN-major tiles
    uint32_t lbo, sbo;      // descriptor fields in bytes
    uint32_t kadv;          // start-address advance per K=8 step in bytes
    uint32_t a_major_bit, b_major_bit;
    uint32_t layout;        // UMMA layout type (2 = SW128, 1 = SW128 base 32B)
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    Variant v, float* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar_full), 1);
        mbar_init(smem_u32(&bar_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t fb = smem_u32(&bar_full);
        if (v.mn_major == 2) {
            // K-major A tile of 136 rows (one extra swizzle atom); the MMA starts v.lbo rows into it
            mbar_expect_tx(fb, 136 * 128 + 64 * 128);
            tma_load_4d(&tmA, fb, sa, 0, 0, 0, 0);
            tma_load_4d(&tmB, fb, sb, 0, 0, 0, 0);
        } else if (!v.mn_major) {
            // K-major: A global [M=128 rows][K=32] -> one box {32 k, 128 rows}; B [N=64][K=32] -> {32, 64}
            mbar_expect_tx(fb, 128 * 128 + 64 * 128);
            tma_load_4d(&tmA, fb, sa, 0, 0, 0, 0);
            tma_load_4d(&tmB, fb, sb, 0, 0, 0, 0);
        } else {
            // MN-major: A global [K=32 positions][M=128 ch] -> 4 boxes {32 ch, 32 pos}; B [K=32][N=64] -> 2 boxes
            mbar_expect_tx(fb, 6 * 4096);
            for (int i = 0; i < 4; ++i) tma_load_4d(&tmA, fb, sa + i * 4096, 32 * i, 0, 0, 0);
            for (int i = 0; i < 2; ++i) tma_load_4d(&tmB, fb, sb + i * 4096, 32 * i, 0, 0, 0);
        }
        mbar_wait(fb, 0);
        tc_fence_after();
        const uint32_t amaj = v.mn_major == 2 ? 0u : v.a_major_bit, bmaj = v.mn_major == 2 ? 0u : v.b_major_bit;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (amaj << 15) | (bmaj << 16) |
                               ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k8 = 0; k8 < 4; ++k8) {
            uint64_t ad = make_desc_full(sa + k8 * v.kadv, v.lbo, v.sbo, v.layout);
            uint64_t bd = make_desc_full(sb + k8 * v.kadv, v.lbo, v.sbo, v.layout);
            if (v.mn_major == 2) {
                ad = make_desc_full(sa + v.lbo * 128 + k8 * 32, 16, 1024, 2) | ((uint64_t)(v.a_major_bit & 7) << 49);   // base_offset
                bd = make_desc_full(sb + k8 * 32, 16, 1024, 2);
            }
            tc_mma_tf32(tmem, ad, bd, idesc, k8 ? 1u : 0u);
        }
        tc_commit(smem_u32(&bar_done));
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    float vals[32];
    for (int ch = 0; ch < 2; ++ch) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, vals);
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + ch * 32 + i] = vals[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

static int make4(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1,
                 CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t dims[4] = {d0, d1, 1, 1};
    cuuint64_t str[3] = {d0 * 4, d0 * d1 * 4, d0 * d1 * 4};
    cuuint32_t box[4] = {b0, b1, 1, 1};
    return make_map(m, base, 4, dims, str, box, swz);
}

int main() {
    const int M = 128, N = 64, K = 32;
    std::vector<float> A(M * K), B(N * K), D(M * N), Akm(M * K), Bkm(N * K), Amn(K * M), Bmn(K * N);
    srand(3);
    for (auto& x : A) x = (float)((rand() % 17) - 8) / 8.0f;   // exactly representable in tf32
    for (auto& x : B) x = (float)((rand() % 13) - 6) / 4.0f;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
            D[m * N + n] = (float)s;
        }
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { Akm[m * K + k] = A[m * K + k]; Amn[k * M + m] = A[m * K + k]; }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { Bkm[n * K + k] = B[n * K + k]; Bmn[k * N + n] = B[n * K + k]; }
    float *dAk, *dBk, *dAm, *dBm, *dout;
    cudaMalloc(&dAk, M * K * 4); cudaMalloc(&dBk, N * K * 4); cudaMalloc(&dAm, M * K * 4); cudaMalloc(&dBm, N * K * 4);
    cudaMalloc(&dout, M * N * 4);
    cudaMemcpy(dAk, Akm.data(), M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBk, Bkm.data(), N * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dAm, Amn.data(), M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBm, Bmn.data(), N * K * 4, cudaMemcpyHostToDevice);
    std::vector<float> A2(136 * K);
    for (auto& x : A2) x = (float)((rand() % 17) - 8) / 8.0f;
    float* dA2; cudaMalloc(&dA2, 136 * K * 4); cudaMemcpy(dA2, A2.data(), 136 * K * 4, cudaMemcpyHostToDevice);
    CUtensorMap mA2;
    if (make4(&mA2, dA2, K, 136, 32, 136)) return 1;
    CUtensorMap mAk, mBk, mAm, mBm, mAm32, mBm32;
    if (make4(&mAk, dAk, K, M, 32, 128) || make4(&mBk, dBk, K, N, 32, 64) || make4(&mAm, dAm, M, K, 32, 32) || make4(&mBm, dBm, N, K, 32, 32)) return 1;
    if (make4(&mAm32, dAm, M, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) || make4(&mBm32, dBm, N, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    Variant vs[] = {
        {0, 16, 1024, 32, 0, 0, 2},         // K-major reference (what tapgemm_tc does)
        {1, 4096, 1024, 1024, 1, 1, 2},     // MN-major with plain SW128: expected unsupported for tf32
        {1, 4096, 512, 1024, 1, 1, 1},      // SW128 base-32B: LBO = chunk stride, SBO = 4-row atom
        {1, 512, 4096, 1024, 1, 1, 1},      // roles swapped
        {1, 4096, 1024, 1024, 1, 1, 1},
        {1, 1024, 4096, 1024, 1, 1, 1},
        {1, 4096, 4096, 1024, 1, 1, 1},
        {1, 512, 512, 1024, 1, 1, 1},
        {2, 0, 1024, 32, 0, 0, 2},          // K-major, 136-row tile, start at row 0 (sanity)
        {2, 1, 1024, 32, 0, 0, 2},          // start at row 1, base_offset 0
        {2, 1, 1024, 32, 1, 0, 2},          // start at row 1, base_offset 1
        {2, 3, 1024, 32, 3, 0, 2},          // start at row 3, base_offset 3
        {2, 3, 1024, 32, 0, 0, 2},          // start at row 3, base_offset 0
    };
    std::vector<float> got(M * N);
    for (size_t i = 0; i < sizeof(vs) / sizeof(vs[0]); ++i) {
        cudaMemset(dout, 0xff, M * N * 4);
        const Variant& v = vs[i];
        if (v.mn_major == 2) probe_kernel<<<1, 128, 64 * 1024>>>(mA2, mBk, v, dout);
        else probe_kernel<<<1, 128, 64 * 1024>>>(v.mn_major ? (v.layout == 1 ? mAm32 : mAm) : mAk, v.mn_major ? (v.layout == 1 ? mBm32 : mBm) : mBk, v, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %zu: CUDA error %s\n", i, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(got.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        double worst = 0, nz = 0; int bad = 0;
        std::vector<float> Dr(D);
        if (v.mn_major == 2)
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double sacc = 0;
                    for (int k = 0; k < K; ++k) sacc += (double)A2[(m + v.lbo) * K + k] * B[n * K + k];
                    Dr[m * N + n] = (float)sacc;
                }
        for (int j = 0; j < M * N; ++j) { double d = fabs((double)got[j] - Dr[j]); if (!(d < 1e-3)) ++bad; if (d > worst) worst = d; nz += got[j] != 0; }
        printf("variant %zu mn=%d layout=%u lbo=%u sbo=%u kadv=%u amaj=%u bmaj=%u : bad %d / %d, max err %.3g, nonzero %.0f, D[0..3]= %g %g %g %g (ref %g %g %g %g)\n",
               i, v.mn_major, v.layout, v.lbo, v.sbo, v.kadv, v.a_major_bit, v.b_major_bit, bad, M * N, worst, nz, got[0], got[1], got[2], got[3],
               D[0], D[1], D[2], D[3]);
    }
    return 0;
}
