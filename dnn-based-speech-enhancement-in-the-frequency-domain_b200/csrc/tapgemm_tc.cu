// Tensor-core "tap-GEMM" for sm_100a: the same contract as tapgemm_simt.cu (conv / convT forward, their data
// gradients, LSTM input projections, output Linear), computed with tcgen05.mma kind::tf32 (fp32 operands read
// as TF32, fp32 accumulation in TMEM).
//
// Persistent, warp-specialised CTA (192 threads, one per SM):
//   warp 0      : TMA producer  - per (tap, 32-channel block): one 4-D box [32 ch x 128 t] of the channels-last
//                 activation (shifted by the tap, zero-filled outside the tensor = conv padding) and one 3-D box
//                 [32 k x BN n] of the packed weights, 128B-swizzled, completing on a full[] mbarrier
//   warp 1      : MMA issuer    - 4 x tcgen05.mma (K = 8) per stage into a 128 x BN fp32 accumulator in TMEM,
//                 tcgen05.commit releases the stage (empty[]) and finally publishes the accumulator (tmem_full[])
//   warps 2..5  : epilogue      - tcgen05.ld 32 columns at a time, + bias, stage through padded shared memory,
//                 coalesced 128-byte row stores (optionally += for gradient accumulation) and per-channel
//                 sum / sum-of-squares for the following BatchNorm; two accumulators (TMEM double buffer) let
//                 the epilogue of tile i overlap the main loop of tile i+1.
#include <cuda.h>

#include "common.cuh"
#include <string.h>

#include "prof.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;                 // tile rows = consecutive time positions of one (b, f) row
constexpr int KB = 32;                  // fp32 channels per k-block = 128 B = swizzle span
constexpr int STG_LD = 33;
constexpr int NTHREADS = 192;

struct TcParams {
    TapDst o[2];
    int accum[2], round_out[2];
    const float* bias;
    long long bJ;
    double* stats;
    int B, J, Tout, Fin;
    int fi_mul, fo_mul, fo_off;
    // work items: one activation tile per item, shared by up to two taps whose time shifts differ by one frame
    int nitems;
    int it_df[SEFD_MAX_TAPS], it_dt0[SEFD_MAX_TAPS], it_n[SEFD_MAX_TAPS];
    int it_slab[SEFD_MAX_TAPS][2], it_roff[SEFD_MAX_TAPS][2];
    int C0, C1, N, wJ_slabs;
    int t_tiles, n_tiles;
    long long total_tiles;
};

// PAIR: the activation tile holds 136 rows (one extra swizzle atom) so that two taps that differ by one frame read
// it at row offsets 0 and 1 (the MMA descriptor start address moves by 128 B; swizzling is on absolute address bits -
// verified on hardware with tools/umma_probe.cu), and the stage carries both taps' weight tiles.
template <int BN>
struct Cfg {
    static constexpr bool PAIR = BN <= 128;
    static constexpr int A_ROWS = PAIR ? 136 : TM;
    static constexpr int A_TILE = A_ROWS * KB * 4;
    static constexpr int B_BYTES = BN * KB * 4;
    static constexpr int NW = PAIR ? 2 : 1;
    static constexpr int STAGE_BYTES = A_TILE + NW * B_BYTES;
    static constexpr int NSTAGE = BN == 256 ? 3 : (BN == 128 ? 3 : (BN == 64 ? 5 : 6));
    static constexpr int STG_BYTES = 2 * TM * STG_LD * 4;
    static constexpr int STAT_BYTES = 2 * BN * 4;
    static constexpr int BAR_BYTES = (2 * NSTAGE + 4) * 8 + 16;
    static constexpr int SMEM = NSTAGE * STAGE_BYTES + STG_BYTES + STAT_BYTES + BAR_BYTES + 1024;   // + align slack
};

struct TileCoord {
    int b, j, t0, n0;
};
__device__ __forceinline__ TileCoord decode(const TcParams& p, long long tile, int BN) {
    TileCoord c;
    c.n0 = (int)(tile % p.n_tiles) * BN;
    tile /= p.n_tiles;
    c.t0 = (int)(tile % p.t_tiles) * TM;
    tile /= p.t_tiles;
    c.j = (int)(tile % p.J);
    c.b = (int)(tile / p.J);
    return c;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
tapgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmW, const TcParams p) {
    using C = Cfg<BN>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* stages = smem;
    float* stg = reinterpret_cast<float*>(smem + C::NSTAGE * C::STAGE_BYTES);
    float* s_stat = stg + 2 * TM * STG_LD;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * BN);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::NSTAGE;
    uint64_t* tfull = bars + 2 * C::NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = p.C0 + p.C1;
    const int kblocks = K / KB;

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
    for (int i = threadIdx.x; i < 2 * BN; i += NTHREADS) s_stat[i] = 0.f;
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
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode(p, tile, BN);
                for (int it = 0; it < p.nitems; ++it) {
                    const int fi = tc.j * p.fi_mul + p.it_df[it];
                    if (fi < 0 || fi >= p.Fin) continue;
                    const int n = p.it_n[it];
                    const int tin = tc.t0 + p.it_dt0[it];
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                        const uint32_t fb = smem_u32(&full[stage]);
                        const uint32_t sa = smem_u32(stages + stage * C::STAGE_BYTES);
                        mbar_expect_tx(fb, (uint32_t)(C::A_TILE + n * C::B_BYTES));
                        const int k = kb * KB;
                        if (k < p.C0) tma_load_4d(&tmA0, fb, sa, k, tin, fi, tc.b);
                        else tma_load_4d(&tmA1, fb, sa, k - p.C0, tin, fi, tc.b);
                        for (int w = 0; w < n; ++w)
                            tma_load_3d(&tmW, fb, sa + C::A_TILE + w * C::B_BYTES, k, tc.n0, tc.j * p.wJ_slabs + p.it_slab[it][w]);
                        if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            int stage = 0, abuf = 0;
            uint32_t phase = 0, aphase = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode(p, tile, BN);
                mbar_wait(smem_u32(&tempty[abuf]), aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * 256);
                uint32_t acc = 0;
                for (int it = 0; it < p.nitems; ++it) {
                    const int fi = tc.j * p.fi_mul + p.it_df[it];
                    if (fi < 0 || fi >= p.Fin) continue;
                    const int n = p.it_n[it];
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(smem_u32(&full[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stages + stage * C::STAGE_BYTES);
                        for (int w = 0; w < n; ++w) {
                            const uint64_t ad = make_desc(sa + (uint32_t)(p.it_roff[it][w] * 128));
                            const uint64_t bd = make_desc(sa + C::A_TILE + w * C::B_BYTES);
#pragma unroll
                            for (int k8 = 0; k8 < KB / 8; ++k8) {
                                tc_mma_tf32(d_tmem, ad + 2 * k8, bd + 2 * k8, idesc, acc);   // +32 B per K=8 step
                                acc = 1;
                            }
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
        // ================= epilogue (128 threads; thread = accumulator row) =================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;        // 0..127
        int abuf = 0, sb = 0;
        uint32_t aphase = 0;
        const int N0 = p.o[0].N;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord tc = decode(p, tile, BN);
            int nk = 0;
            for (int it = 0; it < p.nitems; ++it) {
                const int fi = tc.j * p.fi_mul + p.it_df[it];
                nk += (fi >= 0 && fi < p.Fin);
            }
            mbar_wait(smem_u32(&tfull[abuf]), aphase);
            tc_fence_after();
            const int fo = tc.j * p.fo_mul + p.fo_off;
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * 256 + ch * 32), v);
                if (ch == BN / 32 - 1) {        // accumulator fully read: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tempty[abuf]));
                }
                const int n = tc.n0 + ch * 32;
                float* srow = stg + (sb * TM + row) * STG_LD;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float x = nk ? v[i] : 0.f;
                    if (p.bias) x += __ldg(p.bias + p.bJ * tc.j + n + i);
                    srow[i] = x;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int d = n < N0 ? 0 : 1;
                const int nn = d ? n - N0 : n;
                float* obase = p.o[d].p + tc.b * p.o[d].sB + fo * p.o[d].sF + nn;
                const float* sbuf = stg + sb * TM * STG_LD;
#pragma unroll
                for (int pass = 0; pass < 8; ++pass) {
                    const int idx = pass * 128 + et;
                    const int r = idx >> 3, c4 = (idx & 7) * 4;
                    const int t = tc.t0 + r;
                    if (t < p.Tout) {
                        const float* sp = sbuf + r * STG_LD + c4;
                        float4 o4 = make_float4(sp[0], sp[1], sp[2], sp[3]);
                        float* dst = obase + (long long)t * p.o[d].sT + c4;
                        if (p.accum[d]) {
                            const float4 old = *reinterpret_cast<const float4*>(dst);
                            o4.x += old.x; o4.y += old.y; o4.z += old.z; o4.w += old.w;
                        }
                        if (p.round_out[d]) { o4.x = tf32_rn(o4.x); o4.y = tf32_rn(o4.y); o4.z = tf32_rn(o4.z); o4.w = tf32_rn(o4.w); }
                        *reinterpret_cast<float4*>(dst) = o4;
                    }
                }
                if (p.stats) {
                    const int c = et & 31, part = et >> 5;
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                    for (int i = 0; i < 32; ++i) {
                        const int r = part * 32 + i;
                        if (tc.t0 + r < p.Tout) {
                            const float x = sbuf[r * STG_LD + c];
                            s1 += x;
                            s2 += x * x;
                        }
                    }
                    atomicAdd(&s_stat[ch * 32 + c], s1);
                    atomicAdd(&s_stat[BN + ch * 32 + c], s2);
                }
                sb ^= 1;
            }
            if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = et; i < BN; i += 128) {
                atomicAdd(p.stats + i, (double)s_stat[i]);
                atomicAdd(p.stats + p.N + i, (double)s_stat[BN + i]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------
int make_act_map(CUtensorMap* m, const TapSrc& s, int F, int T, int B, int rows) {
    cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)T, (cuuint64_t)F, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)s.sT * 4, (cuuint64_t)(s.sF ? s.sF : s.sT * T) * 4,
                         (cuuint64_t)(s.sB ? s.sB : s.sT * T * F) * 4};
    cuuint32_t box[4] = {KB, (cuuint32_t)rows, 1, 1};
    return make_map(m, s.p, 4, dims, str, box);
}

template <int BN>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const TcParams& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(tapgemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM);
        attr = true;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
    tapgemm_tc_kernel<BN><<<grid, NTHREADS, Cfg<BN>::SMEM, st>>>(a0, a1, w, p);
    return sefd_check_launch("tapgemm_tc");
}

}  // namespace

bool sefd_tapgemm_tc_eligible(const TapGemmParams& p) {
    const int N = p.o[0].N + p.o[1].N;
    if (!p.Wnk || p.nslabs <= 0) return false;
    if (p.a[0].C % KB || p.a[1].C % KB || p.a[0].C == 0) return false;
    if (N % 32 || (p.o[1].N && p.o[0].N % 32)) return false;
    if (p.stats && N > 256) return false;
    for (int s = 0; s < 2; ++s) {
        if (!p.a[s].C) continue;
        if (((uintptr_t)p.a[s].p & 15) || p.a[s].sT % 4 || p.a[s].sF % 4 || p.a[s].sB % 4) return false;
    }
    for (int d = 0; d < 2; ++d) {
        if (!p.o[d].N) continue;
        if (((uintptr_t)p.o[d].p & 15) || p.o[d].sT % 4 || p.o[d].sF % 4 || p.o[d].sB % 4) return false;
    }
    if ((uintptr_t)p.Wnk & 15) return false;
    return true;
}

int sefd_tapgemm_tc(const TapGemmParams& g, cudaStream_t st) {
    SEFD_REQUIRE(sefd_tapgemm_tc_eligible(g), "tapgemm_tc: problem not eligible for the tensor-core engine");
    const int N = g.o[0].N + g.o[1].N, K = g.a[0].C + g.a[1].C;
    const int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.o[0] = g.o[0]; p.o[1] = g.o[1]; p.accum[0] = g.accum[0]; p.accum[1] = g.accum[1];
    p.round_out[0] = g.round_out[0]; p.round_out[1] = g.round_out[1];
    p.bias = g.bias; p.bJ = g.bJ; p.stats = g.stats;
    p.B = g.B; p.J = g.J; p.Tout = g.Tout; p.Fin = g.Fin;
    p.fi_mul = g.fi_mul; p.fo_mul = g.fo_mul; p.fo_off = g.fo_off;
    // pair up taps on the same source row whose time shifts differ by one frame (BN <= 128 only)
    const bool pair = BN <= 128;
    bool used[SEFD_MAX_TAPS] = {false};
    p.nitems = 0;
    for (int i = 0; i < g.ntaps; ++i) {
        if (used[i]) continue;
        int it = p.nitems++;
        p.it_df[it] = g.df[i]; p.it_dt0[it] = g.dt[i]; p.it_n[it] = 1;
        p.it_slab[it][0] = g.wslab[i]; p.it_roff[it][0] = 0;
        used[i] = true;
        if (!pair) continue;
        for (int j2 = i + 1; j2 < g.ntaps; ++j2) {
            if (used[j2] || g.df[j2] != g.df[i]) continue;
            const int d = g.dt[j2] - g.dt[i];
            if (d == 1 || d == -1) {
                used[j2] = true;
                p.it_n[it] = 2;
                if (d == 1) {
                    p.it_slab[it][1] = g.wslab[j2]; p.it_roff[it][1] = 1;
                } else {            // the partner starts one frame earlier: it becomes row offset 0
                    p.it_dt0[it] = g.dt[j2];
                    p.it_slab[it][1] = g.wslab[i]; p.it_roff[it][1] = 1;
                    p.it_slab[it][0] = g.wslab[j2]; p.it_roff[it][0] = 0;
                }
                break;
            }
        }
    }
    p.C0 = g.a[0].C; p.C1 = g.a[1].C; p.N = N; p.wJ_slabs = g.wJ_slabs;
    p.t_tiles = (g.Tout + TM - 1) / TM;
    p.n_tiles = N / BN;
    p.total_tiles = (long long)g.B * g.J * p.t_tiles * p.n_tiles;

    CUtensorMap a0, a1, w;
    const int arows = pair ? 136 : TM;
    SEFD_TRY(make_act_map(&a0, g.a[0], g.Fin, g.Tin, g.B, arows));
    if (g.a[1].C) SEFD_TRY(make_act_map(&a1, g.a[1], g.Fin, g.Tin, g.B, arows));
    else a1 = a0;
    {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)g.nslabs};
        cuuint64_t str[2] = {(cuuint64_t)(g.w_ldk ? g.w_ldk : K) * 4, (cuuint64_t)(g.w_slab_stride ? g.w_slab_stride : (long long)K * N) * 4};
        cuuint32_t box[3] = {KB, (cuuint32_t)BN, 1};
        SEFD_TRY(make_map(&w, g.Wnk, 3, dims, str, box));
    }
    const double pos = (double)g.B * g.J * g.Tout;
    sefd_prof_label("tapgemm_tc BN%d K%d N%d taps%d items%d J%d Tout%d tiles%lld", BN, K, N, g.ntaps, p.nitems, g.J, g.Tout, p.total_tiles);
    SefdProfScope prof(SEFD_PROF_TAPGEMM, 2.0 * pos * N * K * g.ntaps,
                       4.0 * ((double)g.B * g.J * (g.fi_mul > 1 ? g.fi_mul : 1) * g.Tin * K + pos * N), st);
    switch (BN) {
        case 256: return launch<256>(a0, a1, w, p, st);
        case 128: return launch<128>(a0, a1, w, p, st);
        case 64: return launch<64>(a0, a1, w, p, st);
        default: return launch<32>(a0, a1, w, p, st);
    }
}

static int g_engine = 1;   // 0: fp32 CUDA-core engine everywhere (exact; tests), 1: tcgen05 TF32 where eligible

int sefd_tapgemm(const TapGemmParams& p, cudaStream_t st) {
    if (sefd_skinny_conv_eligible(p)) return sefd_skinny_conv(p, st);
    if (g_engine == 1 && sefd_tapgemm_tc_eligible(p)) return sefd_tapgemm_tc(p, st);
    return sefd_tapgemm_simt(p, st);
}

extern "C" int sefd_set_engine(int engine) {
    SEFD_REQUIRE(engine == 0 || engine == 1, "set_engine: 0 (fp32 CUDA cores) or 1 (tcgen05 TF32)");
    g_engine = engine;
    return 0;
}
extern "C" int sefd_get_engine(void) { return g_engine; }
int sefd_get_engine_internal() { return g_engine; }
