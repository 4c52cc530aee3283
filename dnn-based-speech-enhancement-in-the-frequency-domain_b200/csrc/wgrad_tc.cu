// Tensor-core weight gradient for the tap-GEMM family (sm_100a, tcgen05 kind::tf32):
//   dW[slab(tap)][k][n] = sum_{b,j,t} A[b, j*a_mul+a_off[tap], t+dt[tap], k] * G[b, j*g_mul+g_off[tap], t, n]
// The contraction index is the POSITION (b, j, t); both operands are channels-last, i.e. MN-major for the MMA
// (the channel index is contiguous, positions are strided).  TMA fetches [32 ch x 32 positions] boxes
// (128B span / 32B-atom swizzle, the only MN-major layout tcgen05 accepts for tf32) - 4 per stage for the 128 k-channels of A, BN/32 for G - and one tcgen05.mma consumes 8
// positions.  Work unit = (tap, 128-channel slice of K, BN slice of N, split of the (b,j) rows); every unit
// writes its own partial tile (no atomics), the fold kernel sums the splits.
#include <string.h>

#include "prof.cuh"
#include "tc_common.cuh"

namespace {

constexpr int WM = 128, PB = 32, CHB = PB * 32 * 4;   // 4 KB per [32 ch x 32 pos] box
constexpr int STG_LD = 33, NTHREADS = 192;

struct WgParams {
    float* out;
    long long split_stride;
    int B, J, Tg, Fa, Fg, a_mul, g_mul;
    int ntaps;
    int a_off[SEFD_MAX_TAPS], g_off[SEFD_MAX_TAPS], dt[SEFD_MAX_TAPS], wslab[SEFD_MAX_TAPS];
    int C0, C1, K, N;
    int m_tiles, n_tiles, splits, rows_per_split, t_blocks;
    long long units;
};

template <int BN>
struct WCfg {
    static constexpr int A_BYTES = 4 * CHB;
    static constexpr int G_BYTES = (BN / 32) * CHB;
    static constexpr int STAGE_BYTES = A_BYTES + G_BYTES;
    static constexpr int NSTAGE = BN == 256 ? 3 : (BN == 128 ? 5 : 6);
    static constexpr int SMEM = NSTAGE * STAGE_BYTES + 2 * WM * STG_LD * 4 + (2 * NSTAGE + 4) * 8 + 16 + 1024;
};

struct Unit {
    int tap, m0, n0, r0, r1;
};
__device__ __forceinline__ Unit decode_unit(const WgParams& p, long long u, int BN) {
    Unit x;
    const int split = (int)(u % p.splits);
    u /= p.splits;
    x.n0 = (int)(u % p.n_tiles) * BN;
    u /= p.n_tiles;
    x.m0 = (int)(u % p.m_tiles) * WM;
    x.tap = (int)(u / p.m_tiles);
    x.r0 = split * p.rows_per_split;
    const int rows = p.B * p.J;
    x.r1 = x.r0 + p.rows_per_split < rows ? x.r0 + p.rows_per_split : rows;
    return x;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmG, const WgParams p) {
    using C = WCfg<BN>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* stages = smem;
    float* stg = reinterpret_cast<float*>(smem + C::NSTAGE * C::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 2 * WM * STG_LD);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::NSTAGE;
    uint64_t* tfull = bars + 2 * C::NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NSTAGE; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tfull[a]), 1);
            mbar_init(smem_u32(&tempty[a]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
                const Unit x = decode_unit(p, u, BN);
                const int nA = (p.K - x.m0) / 32 < 4 ? (p.K - x.m0) / 32 : 4;
                const int dt = p.dt[x.tap];
                for (int r = x.r0; r < x.r1; ++r) {
                    const int b = r / p.J, j = r % p.J;
                    const int fa = j * p.a_mul + p.a_off[x.tap], fg = j * p.g_mul + p.g_off[x.tap];
                    if (fa < 0 || fa >= p.Fa || fg < 0 || fg >= p.Fg) continue;
                    for (int tb = 0; tb < p.t_blocks; ++tb) {
                        mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                        const uint32_t fb = smem_u32(&full[stage]);
                        const uint32_t sa = smem_u32(stages + stage * C::STAGE_BYTES);
                        mbar_expect_tx(fb, (uint32_t)((nA + BN / 32) * CHB));
                        const int t0 = tb * PB;
                        for (int i = 0; i < nA; ++i) {
                            const int kc = x.m0 + 32 * i;
                            if (kc < p.C0) tma_load_4d(&tmA0, fb, sa + i * CHB, kc, t0 + dt, fa, b);
                            else tma_load_4d(&tmA1, fb, sa + i * CHB, kc - p.C0, t0 + dt, fa, b);
                        }
#pragma unroll
                        for (int i = 0; i < BN / 32; ++i)
                            tma_load_4d(&tmG, fb, sa + C::A_BYTES + i * CHB, x.n0 + 32 * i, t0, fg, b);
                        if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D fp32, A/B tf32, both MN-major (bits 15, 16), N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);
            int stage = 0, abuf = 0;
            uint32_t phase = 0, aphase = 0;
            for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
                const Unit x = decode_unit(p, u, BN);
                mbar_wait(smem_u32(&tempty[abuf]), aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * 256);
                uint32_t acc = 0;
                for (int r = x.r0; r < x.r1; ++r) {
                    const int j = r % p.J;
                    const int fa = j * p.a_mul + p.a_off[x.tap], fg = j * p.g_mul + p.g_off[x.tap];
                    if (fa < 0 || fa >= p.Fa || fg < 0 || fg >= p.Fg) continue;
                    for (int tb = 0; tb < p.t_blocks; ++tb) {
                        mbar_wait(smem_u32(&full[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stages + stage * C::STAGE_BYTES);
#pragma unroll
                        for (int k8 = 0; k8 < PB / 8; ++k8) {
                            // MN-major tf32: SWIZZLE_128B_BASE32B (layout type 1), LBO = stride between 32-channel
                            // chunks, SBO = one 4-position swizzle atom (probed on hardware: tools/umma_probe.cu)
                            const uint64_t ad = make_desc_full(sa + k8 * 1024, CHB, 512, 1);
                            const uint64_t bd = make_desc_full(sa + C::A_BYTES + k8 * 1024, CHB, 512, 1);
                            tc_mma_tf32(d_tmem, ad, bd, idesc, acc);
                            acc = 1;
                        }
                        tc_commit(smem_u32(&empty[stage]));
                        if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
                tc_commit(smem_u32(&tfull[abuf]));
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;
        int abuf = 0, sb = 0;
        uint32_t aphase = 0;
        for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
            const Unit x = decode_unit(p, u, BN);
            int nk = 0;
            for (int r = x.r0; r < x.r1; ++r) {
                const int j = r % p.J;
                const int fa = j * p.a_mul + p.a_off[x.tap], fg = j * p.g_mul + p.g_off[x.tap];
                nk += (fa >= 0 && fa < p.Fa && fg >= 0 && fg < p.Fg);
            }
            const int split = (int)(u % p.splits);
            float* obase = p.out + split * p.split_stride + (long long)p.wslab[x.tap] * p.K * p.N + x.n0;
            mbar_wait(smem_u32(&tfull[abuf]), aphase);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * 256 + ch * 32), v);
                if (ch == BN / 32 - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tempty[abuf]));
                }
                float* srow = stg + (sb * WM + row) * STG_LD;
#pragma unroll
                for (int i = 0; i < 32; ++i) srow[i] = nk ? v[i] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const float* sbuf = stg + sb * WM * STG_LD;
#pragma unroll
                for (int pass = 0; pass < 8; ++pass) {
                    const int idx = pass * 128 + et;
                    const int r = idx >> 3, c4 = (idx & 7) * 4;
                    const int k = x.m0 + r;
                    if (k < p.K) {
                        const float* sp = sbuf + r * STG_LD + c4;
                        *reinterpret_cast<float4*>(obase + (long long)k * p.N + ch * 32 + c4) =
                            make_float4(sp[0], sp[1], sp[2], sp[3]);
                    }
                }
                sb ^= 1;
            }
            if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int make_pos_map(CUtensorMap* m, const TapSrc& s, int F, int T, int B) {
    cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)T, (cuuint64_t)F, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)s.sT * 4, (cuuint64_t)(s.sF ? s.sF : s.sT * T) * 4,
                         (cuuint64_t)(s.sB ? s.sB : s.sT * T * F) * 4};
    cuuint32_t box[4] = {32, PB, 1, 1};
    return make_map(m, s.p, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

template <int BN>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& g, const WgParams& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<BN>::SMEM);
        attr = true;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(p.units < sms ? p.units : sms);
    wgrad_tc_kernel<BN><<<grid, NTHREADS, WCfg<BN>::SMEM, st>>>(a0, a1, g, p);
    return sefd_check_launch("wgrad_tc");
}

}  // namespace

bool sefd_wgrad_tc_eligible(const WgradParams& p) {
    if (p.a[0].C % 32 || p.a[1].C % 32 || p.a[0].C == 0 || p.g.C % 32 || p.g.C == 0) return false;
    const TapSrc* v[3] = {&p.a[0], &p.a[1], &p.g};
    for (int i = 0; i < 3; ++i) {
        if (!v[i]->C) continue;
        if (((uintptr_t)v[i]->p & 15) || v[i]->sT % 4 || v[i]->sF % 4 || v[i]->sB % 4) return false;
    }
    return true;
}

// Writes `*nsplit` partial gradients, `split_stride` floats apart, starting at `partial`.
int sefd_wgrad_tc(const WgradParams& w, float* partial, long long cap_floats, int nslabs, int* nsplit,
                  long long* split_stride, cudaStream_t st) {
    SEFD_REQUIRE(sefd_wgrad_tc_eligible(w), "wgrad_tc: problem not eligible for the tensor-core engine");
    const int K = w.a[0].C + w.a[1].C, N = w.g.C;
    const int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
    WgParams p;
    memset(&p, 0, sizeof(p));
    p.B = w.B; p.J = w.J; p.Tg = w.Tg; p.Fa = w.Fa; p.Fg = w.Fg; p.a_mul = w.a_mul; p.g_mul = w.g_mul;
    p.ntaps = w.ntaps;
    for (int i = 0; i < w.ntaps; ++i) {
        p.a_off[i] = w.a_off[i]; p.g_off[i] = w.g_off[i]; p.dt[i] = w.dt[i]; p.wslab[i] = w.wslab[i];
    }
    p.C0 = w.a[0].C; p.C1 = w.a[1].C; p.K = K; p.N = N;
    p.m_tiles = (K + WM - 1) / WM;
    p.n_tiles = N / BN;
    p.t_blocks = (w.Tg + PB - 1) / PB;
    const int rows = w.B * w.J;
    const long long tiles = (long long)w.ntaps * p.m_tiles * p.n_tiles;
    const long long one = (long long)nslabs * K * N;
    long long splits = (2 * 148 + tiles - 1) / tiles;
    if (splits > rows) splits = rows;
    if (splits > cap_floats / one) splits = cap_floats / one;
    if (splits > 64) splits = 64;
    SEFD_REQUIRE(splits >= 1, "wgrad_tc: partial buffer too small");
    p.rows_per_split = (int)((rows + splits - 1) / splits);
    p.splits = (rows + p.rows_per_split - 1) / p.rows_per_split;
    p.units = tiles * p.splits;
    p.out = partial;
    p.split_stride = one;
    *nsplit = p.splits;
    *split_stride = one;

    CUtensorMap a0, a1, g;
    SEFD_TRY(make_pos_map(&a0, w.a[0], w.Fa, w.Ta, w.B));
    if (w.a[1].C) SEFD_TRY(make_pos_map(&a1, w.a[1], w.Fa, w.Ta, w.B));
    else a1 = a0;
    SEFD_TRY(make_pos_map(&g, w.g, w.Fg, w.Tg, w.B));
    const double pos = (double)w.B * w.J * w.Tg;
    sefd_prof_label("wgrad_tc BN%d K%d N%d taps%d J%d splits%d units%lld", BN, K, N, w.ntaps, w.J, p.splits, p.units);
    SefdProfScope prof(SEFD_PROF_WGRAD, 2.0 * pos * K * N * w.ntaps,
                       4.0 * ((double)w.B * w.J * (w.a_mul > 1 ? w.a_mul : 1) * w.Ta * K +
                              (double)w.B * w.J * (w.g_mul > 1 ? w.g_mul : 1) * w.Tg * N), st);
    switch (BN) {
        case 256: return launch<256>(a0, a1, g, p, st);
        case 128: return launch<128>(a0, a1, g, p, st);
        case 64: return launch<64>(a0, a1, g, p, st);
        default: return launch<32>(a0, a1, g, p, st);
    }
}

// dispatch: returns the number of partial buffers the fold step has to sum
int sefd_wgrad(const WgradParams& w, float* partial, long long cap_floats, int nslabs, int* nsplit,
               long long* split_stride, cudaStream_t st) {
    if (sefd_get_engine_internal() == 1 && sefd_wgrad_tc_eligible(w))
        return sefd_wgrad_tc(w, partial, cap_floats, nslabs, nsplit, split_stride, st);
    const long long one = (long long)nslabs * (w.a[0].C + w.a[1].C) * w.g.C;
    SEFD_REQUIRE(one <= cap_floats, "wgrad: gradient scratch too small");
    cudaMemsetAsync(partial, 0, sizeof(float) * one, st);
    WgradParams p = w;
    p.dW = partial;
    *nsplit = 1;
    *split_stride = one;
    if (sefd_skinny_wgrad_eligible(p)) return sefd_skinny_wgrad(p, st);
    return sefd_wgrad_simt(p, st);
}
