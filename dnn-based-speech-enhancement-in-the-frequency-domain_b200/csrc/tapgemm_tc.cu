// Tensor-core "tap-GEMM" for sm_100a: the same contract as tapgemm_simt.cu (conv / convT forward, their data
// gradients, LSTM input projections, output Linear), computed with tcgen05.mma kind::tf32 (fp32 operands read
// as TF32, fp32 accumulation in TMEM).
//
// Persistent, warp-specialised CTA (192 threads, one per SM):
//   warp 0      : TMA producer  - per (work item, group of k-blocks) ONE 5-D box of the channels-last activation
//                 ([32 ch x 128|136 t x kbs channel blocks], shifted by the tap, zero-filled outside the tensor =
//                 conv padding) and ONE 4-D box of the packed weights ([32 k x BN n x kbs x 1|2 taps]), all
//                 128B-swizzled, completing on a full[] mbarrier.  A work item is one source row with one or two
//                 taps: two taps that differ by one frame share the activation tile (136 rows; the second tap's
//                 MMA descriptor starts 128 B = one row further).
//   warp 1      : MMA issuer    - 4 x tcgen05.mma (K = 8) per (tap, k-block) into a 128 x BN fp32 accumulator in
//                 TMEM; tcgen05.commit releases the stage (empty[]) and publishes the accumulator (tfull[])
//   warps 2..5  : epilogue      - tcgen05.ld 32 columns at a time (thread = row: 128 contiguous output bytes), + bias,
//                 optional += / tf32 rounding, four 256-bit stores per thread straight from registers (no smem
//                 staging, no barriers); per-channel sum / sum-of-squares for the following BatchNorm through a
//                 31-shuffle transposing butterfly; two accumulators (TMEM double buffer) let the epilogue of
//                 tile i overlap the main loop of tile i+1.
// (ncu, round 1: with one TMA instruction per 16-24 KB the single producer thread was ~64 % busy issuing and the
//  small-tile layers sat at a 0.3 ms floor; hence the multi-block boxes.)
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "prof.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;                 // tile rows = consecutive time positions of one (b, f) row
constexpr int KB = 32;                  // fp32 channels per k-block = 128 B = swizzle span
constexpr int NTHREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 225 * 1024;
#ifndef SEFD_FUSE_UP_MAX_N
#define SEFD_FUSE_UP_MAX_N 128      // widest N for which the two output-row phases of an up-conv share one launch
#endif

struct TcParams {
    TapDst o[2];
    int accum[2], round_out[2];
    const float* bias;
    long long bJ;
    double* stats;
    int B, J, Tout, Fin;
    int fi_mul, fo_mul, fo_off;
    // work items: one activation tile per item, shared by up to two taps whose time shifts differ by one frame
    int nitems, nacc;                                    // nacc: output rows (accumulators) per tile, 1 or 2
    int it_df[SEFD_MAX_TAPS], it_dt0[SEFD_MAX_TAPS], it_n[SEFD_MAX_TAPS];
    int it_slab_lo[SEFD_MAX_TAPS], it_wbox[SEFD_MAX_TAPS];   // first weight slab / slabs (1, 2 or 4) of the item's box
    // per tap of an item: weight sub-tile inside the box, activation row offset (0/1), accumulator (output-row phase)
    int it_wt[SEFD_MAX_TAPS][4], it_roff[SEFD_MAX_TAPS][4], it_acc[SEFD_MAX_TAPS][4];
    // merged item (two output rows per tile): its four taps are (phase 0 | phase 1) x (frame shift a | b) on ONE activation
    // tile; the weight box lands as [shift][k-block][phase][n][k], so that the two phases' tiles of one shift are 2 BN
    // consecutive operand rows and ONE MMA of width 2 BN feeds both accumulators (adjacent in TMEM): the A operand is read
    // from shared memory once per shift instead of once per tap.  it_mroff[it][s]: activation row offset of shift s.
    int it_merge[SEFD_MAX_TAPS], it_mroff[SEFD_MAX_TAPS][2];
    int C0, C1, N, wJ_slabs;
    int t_tiles, n_tiles;
    long long total_tiles;
    int kbs, nstage, stage_bytes, a_bytes;               // k-blocks per stage, pipeline depth, bytes
    int vec8;                                            // destinations are 32-byte aligned: 256-bit stores
};

template <int BN>
struct Cfg {
    static constexpr bool PAIR = BN <= 128;
    static constexpr int A_ROWS = PAIR ? 136 : TM;
    static constexpr int A_TILE = A_ROWS * KB * 4;
    static constexpr int B_BYTES = BN * KB * 4;
    static constexpr int FIXED = 8 * BN * 4 + (2 * MAX_STAGES + 4) * 8 + 64;
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
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmW4, const __grid_constant__ CUtensorMap tmWm, const TcParams p) {
    using C = Cfg<BN>;
    // declared alignment keeps every derived pointer in the shared address space (LDS/STS/ATOMS, not generic)
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_stat = reinterpret_cast<float*>(smem + p.nstage * p.stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 8 * BN);   // s_stat[part 0..3][sum | sum sq][BN]
    uint64_t* full = bars;
    uint64_t* empty = bars + MAX_STAGES;
    uint64_t* tfull = bars + 2 * MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    // canonical warp index: the shuffle makes it warp-uniform FOR THE COMPILER, so the role branches below are uniform
    // branches and the MMA / TMA warps keep their descriptors in uniform registers (no R2UR per instruction)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int kgroups = (p.C0 + p.C1) / KB / p.kbs;
    const int kg0 = p.C0 / KB / p.kbs;          // k-groups that come from source 0

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();   // the swizzled tiles need 1024-byte aligned bases
        for (int s = 0; s < p.nstage; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tfull[a]), 1);
            mbar_init(smem_u32(&tempty[a]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 8 * BN; i += NTHREADS) s_stat[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0) {
        // ================= TMA producer (warp-uniform loops, elected lane issues) =================
        {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode(p, tile, BN);
                for (int it = 0; it < p.nitems; ++it) {
                    const int fi = tc.j * p.fi_mul + p.it_df[it];
                    if (fi < 0 || fi >= p.Fin) continue;
                    const int wbox = p.it_wbox[it];
                    const int tin = tc.t0 + p.it_dt0[it];
                    const int slab = tc.j * p.wJ_slabs + p.it_slab_lo[it];
                    const uint32_t bytes = (uint32_t)(p.a_bytes + wbox * p.kbs * C::B_BYTES);
                    for (int kg = 0; kg < kgroups; ++kg) {
                        mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                        const uint32_t fb = smem_u32(&full[stage]);
                        const uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
                        if (elect_one_sync()) {
                            mbar_expect_tx(fb, bytes);
                            if (kg < kg0) tma_load_5d(&tmA0, fb, sa, 0, tin, kg * p.kbs, fi, tc.b);
                            else tma_load_5d(&tmA1, fb, sa, 0, tin, (kg - kg0) * p.kbs, fi, tc.b);
                            if (p.it_merge[it]) tma_load_5d(&tmWm, fb, sa + p.a_bytes, 0, tc.n0, slab >> 1, kg * p.kbs, 0);
                            else if (wbox == 4) tma_load_4d(&tmW4, fb, sa + p.a_bytes, 0, tc.n0, kg * p.kbs, slab);
                            else if (wbox == 2) tma_load_4d(&tmW2, fb, sa + p.a_bytes, 0, tc.n0, kg * p.kbs, slab);
                            else tma_load_4d(&tmW1, fb, sa + p.a_bytes, 0, tc.n0, kg * p.kbs, slab);
                        }
                        __syncwarp();
                        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform loops, elected lane issues) =================
        {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN > 256 ? 256 : 2 * BN) >> 3) << 17) |
                                    ((uint32_t)(TM >> 4) << 24);                                    // N = 2 BN (merged items, BN <= 128)
            int stage = 0, abuf = 0;
            uint32_t phase = 0, aphase = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode(p, tile, BN);
                mbar_wait(smem_u32(&tempty[abuf]), aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * 256);
                uint32_t acc0 = 0, acc1 = 0;            // per accumulator: 0 until its first MMA of this tile
                for (int it = 0; it < p.nitems; ++it) {
                    const int fi = tc.j * p.fi_mul + p.it_df[it];
                    if (fi < 0 || fi >= p.Fin) continue;
                    const int n = p.it_n[it];
                    for (int kg = 0; kg < kgroups; ++kg) {
                        mbar_wait(smem_u32(&full[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
                        if (p.it_merge[it]) {
                            // both accumulators in one MMA of width 2 BN per frame shift (they start together: merged items
                            // come first in the item list, so acc0 == acc1 here)
                            for (int sh = 0; sh < 2; ++sh) {
                                for (int kb = 0; kb < p.kbs; ++kb) {
                                    const uint64_t ad = make_desc(sa + (uint32_t)(kb * C::A_TILE + p.it_mroff[it][sh] * 128));
                                    const uint64_t bd = make_desc(sa + (uint32_t)(p.a_bytes + (sh * p.kbs + kb) * 2 * C::B_BYTES));
#pragma unroll
                                    for (int k8 = 0; k8 < KB / 8; ++k8) {
                                        if (elect_one_sync()) tc_mma_tf32(d_tmem, ad + 2 * k8, bd + 2 * k8, idesc2, acc0);
                                        acc0 = acc1 = 1;
                                    }
                                }
                            }
                        } else
                        for (int w = 0; w < n; ++w) {
                            const int a = p.it_acc[it][w];
                            const uint32_t dt = d_tmem + (uint32_t)(a * BN);
                            for (int kb = 0; kb < p.kbs; ++kb) {
                                const uint64_t ad = make_desc(sa + (uint32_t)(kb * C::A_TILE + p.it_roff[it][w] * 128));
                                const uint64_t bd = make_desc(sa + (uint32_t)(p.a_bytes + (p.it_wt[it][w] * p.kbs + kb) * C::B_BYTES));
#pragma unroll
                                for (int k8 = 0; k8 < KB / 8; ++k8) {
                                    const uint32_t accf = a ? acc1 : acc0;
                                    if (elect_one_sync()) tc_mma_tf32(dt, ad + 2 * k8, bd + 2 * k8, idesc, accf);   // +32 B per K=8 step
                                    if (a) acc1 = 1; else acc0 = 1;
                                }
                            }
                        }
                        if (elect_one_sync()) tc_commit(smem_u32(&empty[stage]));
                        __syncwarp();
                        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                    }
                }
                if (elect_one_sync()) tc_commit(smem_u32(&tfull[abuf]));
                __syncwarp();
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (128 threads; thread = accumulator row) =================
        // Straight from registers: tcgen05.ld hands every thread 32 consecutive channels of its own row (128 B), which
        // leave as four 256-bit stores (full 32-byte sectors) - no shared-memory staging, no barriers.  The BatchNorm
        // statistics are reduced with a 31-shuffle transposing butterfly (lane c ends up with column c's sum over the
        // warp's 32 rows) into per-warp slots of s_stat (fixed order: deterministic).
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;        // 0..127
        int abuf = 0;
        uint32_t aphase = 0;
        const int N0 = p.o[0].N;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord tc = decode(p, tile, BN);
            int nk0 = 0, nk1 = 0;                // taps that contributed to each accumulator (0: the row is bias only)
            for (int it = 0; it < p.nitems; ++it) {
                const int fi = tc.j * p.fi_mul + p.it_df[it];
                if (fi >= 0 && fi < p.Fin)
                    for (int w = 0; w < p.it_n[it]; ++w) {
                        if (p.it_acc[it][w]) ++nk1; else ++nk0;
                    }
            }
            mbar_wait(smem_u32(&tfull[abuf]), aphase);
            tc_fence_after();
            const int t = tc.t0 + row;
            const bool rv = t < p.Tout;
            const int nchunks = p.nacc * (BN / 32);
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                const int a = cc / (BN / 32), ch = cc - a * (BN / 32);
                const int nk = a ? nk1 : nk0;
                const int fo = tc.j * p.fo_mul + p.fo_off + a;
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * 256 + a * BN + ch * 32), v);
                if (cc == nchunks - 1) {        // accumulators fully read: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tempty[abuf]));
                }
                const int n = tc.n0 + ch * 32;
                if (!nk) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                if (p.bias) {                   // same 128 bytes for every lane: L1 broadcast
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + p.bJ * tc.j + n);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = __ldg(bp + i);
                        v[4 * i + 0] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
                    }
                }
                const int d = n < N0 ? 0 : 1;
                const int nn = d ? n - N0 : n;
                if (rv) {
                    float* dst = p.o[d].p + tc.b * p.o[d].sB + fo * p.o[d].sF + (long long)t * p.o[d].sT + nn;
                    if (p.accum[d]) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 o4 = __ldcg(reinterpret_cast<const float4*>(dst) + i);
                            v[4 * i + 0] += o4.x; v[4 * i + 1] += o4.y; v[4 * i + 2] += o4.z; v[4 * i + 3] += o4.w;
                        }
                    }
                    if (p.round_out[d]) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
                    }
                    if (p.vec8) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * i),
                                         "f"(v[8 * i]), "f"(v[8 * i + 1]), "f"(v[8 * i + 2]), "f"(v[8 * i + 3]), "f"(v[8 * i + 4]),
                                         "f"(v[8 * i + 5]), "f"(v[8 * i + 6]), "f"(v[8 * i + 7])
                                         : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                }
                if (p.stats) {
                    float w[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] = rv ? v[i] : 0.f;
                        w[i] = v[i] * v[i];
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float sv = up ? v[i] : v[i + off], kv = up ? v[i + off] : v[i];
                            const float sw = up ? w[i] : w[i + off], kw = up ? w[i + off] : w[i];
                            v[i] = kv + __shfl_xor_sync(0xffffffffu, sv, off);
                            w[i] = kw + __shfl_xor_sync(0xffffffffu, sw, off);
                        }
                    }
                    // slot (warp quarter, column) is owned by this lane: plain adds, fixed order
                    s_stat[q * 2 * BN + ch * 32 + lane] += v[0];
                    s_stat[q * 2 * BN + BN + ch * 32 + lane] += w[0];
                }
            }
            if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = et; i < BN; i += 128) {
                const float a = (s_stat[i] + s_stat[2 * BN + i]) + (s_stat[4 * BN + i] + s_stat[6 * BN + i]);
                const float b = (s_stat[BN + i] + s_stat[3 * BN + i]) + (s_stat[5 * BN + i] + s_stat[7 * BN + i]);
                atomicAdd(p.stats + i, (double)a);
                atomicAdd(p.stats + p.N + i, (double)b);
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

template <int BN>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w1, const CUtensorMap& w2, const CUtensorMap& w4,
           const CUtensorMap& wm, const TcParams& p, int smem, cudaStream_t st) {
    static int cur = 0;
    if (smem > cur) {
        cudaFuncSetAttribute(tapgemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cur = smem;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
    tapgemm_tc_kernel<BN><<<grid, NTHREADS, smem, st>>>(a0, a1, w1, w2, w4, wm, p);
    return sefd_check_launch("tapgemm_tc");
}

int make_w_map(CUtensorMap* m, const TapGemmParams& g, int K, int N, int BN, int kbs, int nslab_box) {
    // packed weights [slab][N][K] seen as (k_inner 32, n, k_block, slab): one box = kbs k-blocks x nslab_box taps
    cuuint64_t dims[4] = {32, (cuuint64_t)N, (cuuint64_t)(K / 32), (cuuint64_t)g.nslabs};
    cuuint64_t str[3] = {(cuuint64_t)(g.w_ldk ? g.w_ldk : K) * 4, 128,
                         (cuuint64_t)(g.w_slab_stride ? g.w_slab_stride : (long long)K * N) * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)BN, (cuuint32_t)kbs, (cuuint32_t)nslab_box};
    return make_map(m, g.Wnk, 4, dims, str, box);
}

int make_w_map_merged(CUtensorMap* m, const TapGemmParams& g, int K, int N, int BN, int kbs) {
    // slab = 4 m + 2 phase + shift: (k_inner 32, n, q = 2 m + phase [stride 2 slabs], k_block, shift [stride 1 slab]); the box
    // {32, BN, 2, kbs, 2} lands in shared memory as [shift][k-block][phase][n][k]
    const cuuint64_t slab_bytes = (cuuint64_t)(g.w_slab_stride ? g.w_slab_stride : (long long)K * N) * 4;
    cuuint64_t dims[5] = {32, (cuuint64_t)N, (cuuint64_t)(g.nslabs / 2), (cuuint64_t)(K / 32), 2};
    cuuint64_t str[4] = {(cuuint64_t)(g.w_ldk ? g.w_ldk : K) * 4, 2 * slab_bytes, 128, slab_bytes};
    cuuint32_t box[5] = {32, (cuuint32_t)BN, 2, (cuuint32_t)kbs, 2};
    return make_map(m, g.Wnk, 5, dims, str, box);
}

template <int BN>
int run(const TapGemmParams& g, TcParams& p, cudaStream_t st) {
    using C = Cfg<BN>;
    const int N = p.N, K = p.C0 + p.C1;
    // k-blocks per stage: as many as keep >= 3 stages in flight
    int wmax = 1;                                  // widest weight box of any item (1, 2 or 4 taps)
    for (int it = 0; it < p.nitems; ++it)
        if (p.it_wbox[it] > wmax) wmax = p.it_wbox[it];
    int kbs = 4;
    for (;; kbs >>= 1) {
        const bool div = (p.C0 / KB) % kbs == 0 && (p.C1 == 0 || (p.C1 / KB) % kbs == 0);
        const int stage = kbs * (C::A_TILE + wmax * C::B_BYTES);
        if (kbs == 1 || (div && (SMEM_BUDGET - C::FIXED) / stage >= 3)) break;
    }
    p.kbs = kbs;
    p.a_bytes = kbs * C::A_TILE;
    p.stage_bytes = kbs * (C::A_TILE + wmax * C::B_BYTES);
    p.nstage = (SMEM_BUDGET - C::FIXED) / p.stage_bytes;
    if (p.nstage > MAX_STAGES) p.nstage = MAX_STAGES;
    SEFD_REQUIRE(p.nstage >= 2, "tapgemm_tc: no room for a pipeline (stage %d bytes)", p.stage_bytes);
    const int smem = p.nstage * p.stage_bytes + C::FIXED;

    CUtensorMap a0, a1, w1, w2, w4;
    SEFD_TRY(make_act_map5(&a0, g.a[0], g.Fin, g.Tin, g.B, C::A_ROWS, kbs));
    if (g.a[1].C) SEFD_TRY(make_act_map5(&a1, g.a[1], g.Fin, g.Tin, g.B, C::A_ROWS, kbs));
    else a1 = a0;
    SEFD_TRY(make_w_map(&w1, g, K, N, BN, kbs, 1));
    if (wmax >= 2) SEFD_TRY(make_w_map(&w2, g, K, N, BN, kbs, 2));
    else w2 = w1;
    if (wmax >= 4) SEFD_TRY(make_w_map(&w4, g, K, N, BN, kbs, 4));
    else w4 = w1;
    CUtensorMap wm = w1;
    bool any_merge = false;
    for (int it = 0; it < p.nitems; ++it) any_merge = any_merge || p.it_merge[it];
    if (any_merge) SEFD_TRY(make_w_map_merged(&wm, g, K, N, BN, kbs));
    const double pos = (double)g.B * g.J * g.Tout;
    sefd_prof_label("tapgemm_tc BN%d K%d N%d taps%d items%d acc%d kbs%d x%d J%d Tout%d tiles%lld", BN, K, N, g.ntaps, p.nitems,
                    p.nacc, kbs, p.nstage, g.J, g.Tout, p.total_tiles);
    SefdProfScope prof(SEFD_PROF_TAPGEMM, 2.0 * pos * N * K * g.ntaps,
                       4.0 * ((double)g.B * g.J * (g.fi_mul > 1 ? g.fi_mul : 1) * g.Tin * K + pos * N * p.nacc), st);
    return launch<BN>(a0, a1, w1, w2, w4, wm, p, smem, st);
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
    if (p.bias && (((uintptr_t)p.bias & 15) || p.bJ % 4)) return false;
    return true;
}

int sefd_tapgemm_tc(const TapGemmParams& g, cudaStream_t st) {
    SEFD_REQUIRE(sefd_tapgemm_tc_eligible(g), "tapgemm_tc: problem not eligible for the tensor-core engine");
    const int N = g.o[0].N + g.o[1].N;
    const int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.o[0] = g.o[0]; p.o[1] = g.o[1]; p.accum[0] = g.accum[0]; p.accum[1] = g.accum[1];
    p.round_out[0] = g.round_out[0]; p.round_out[1] = g.round_out[1];
    p.bias = g.bias; p.bJ = g.bJ; p.stats = g.stats;
    p.B = g.B; p.J = g.J; p.Tout = g.Tout; p.Fin = g.Fin;
    p.fi_mul = g.fi_mul; p.fo_mul = g.fo_mul; p.fo_off = g.fo_off;
    // Work items: taps that read the same source row in a one-frame window share ONE activation tile (136 rows; the
    // later tap's MMA descriptor starts one row further) when their weight slabs fit one box of 1, 2 or 4 consecutive
    // slabs (BN <= 128 only: there the activation traffic dominates).  With nacc = 2 the taps of both output-row
    // phases land in the same item and feed different accumulators.
    p.nacc = g.nacc == 2 ? 2 : 1;
    SEFD_REQUIRE(p.nacc == 1 || BN <= 128, "tapgemm_tc: two output rows per tile need N <= 128");
    const bool pair = BN <= 128;
    static const bool merge_ok = getenv("SEFD_TAPGEMM_MERGE") == nullptr || atoi(getenv("SEFD_TAPGEMM_MERGE")) != 0;
    bool used[SEFD_MAX_TAPS] = {false};
    p.nitems = 0;
    for (int i = 0; i < g.ntaps; ++i) {
        if (used[i]) continue;
        const int it = p.nitems++;
        int mem[4], nm = 0;
        mem[nm++] = i;
        used[i] = true;
        int dt_lo = g.dt[i], dt_hi = g.dt[i], s_lo = g.wslab[i], s_hi = g.wslab[i];
        if (pair) {
            for (int j2 = i + 1; j2 < g.ntaps && nm < 4; ++j2) {
                if (used[j2] || g.df[j2] != g.df[i]) continue;
                const int nlo = g.dt[j2] < dt_lo ? g.dt[j2] : dt_lo, nhi = g.dt[j2] > dt_hi ? g.dt[j2] : dt_hi;
                const int slo = g.wslab[j2] < s_lo ? g.wslab[j2] : s_lo, shi = g.wslab[j2] > s_hi ? g.wslab[j2] : s_hi;
                int span = shi - slo + 1;
                span = span <= 1 ? 1 : (span <= 2 ? 2 : 4);
                if (nhi - nlo > 1 || shi - slo >= 4 || slo + span > g.nslabs) continue;
                mem[nm++] = j2;
                used[j2] = true;
                dt_lo = nlo; dt_hi = nhi; s_lo = slo; s_hi = shi;
            }
        }
        const int span = s_hi - s_lo + 1;
        p.it_df[it] = g.df[i];
        p.it_dt0[it] = dt_lo;
        p.it_n[it] = nm;
        p.it_slab_lo[it] = s_lo;
        p.it_wbox[it] = span <= 1 ? 1 : (span <= 2 ? 2 : 4);
        for (int w = 0; w < nm; ++w) {
            const int tp = mem[w];
            p.it_roff[it][w] = g.dt[tp] - dt_lo;
            p.it_wt[it][w] = g.wslab[tp] - s_lo;
            p.it_acc[it][w] = p.nacc == 2 ? g.tap_acc[tp] : 0;
        }
        // merged form: slabs s_lo + 2 phase + shift with s_lo a multiple of 4, both shifts present for both phases, the two
        // phases of a shift reading the activation tile at the same row offset
        p.it_merge[it] = 0;
        if (merge_ok && p.nacc == 2 && nm == 4 && s_lo % 4 == 0 && g.wJ_slabs == 0 && g.nslabs % 2 == 0) {
            int roff[2][2] = {{-1, -1}, {-1, -1}};
            bool ok = true;
            for (int w = 0; w < 4; ++w) {
                const int rel = p.it_wt[it][w], ph = rel >> 1, sh = rel & 1;
                ok = ok && p.it_acc[it][w] == ph && roff[ph][sh] < 0;
                roff[ph][sh] = p.it_roff[it][w];
            }
            ok = ok && roff[0][0] == roff[1][0] && roff[0][1] == roff[1][1] && roff[0][0] >= 0 && roff[0][1] >= 0;
            for (int e = 0; e < it; ++e) ok = ok && p.it_merge[e];      // merged items must come first (common accumulate flag)
            if (ok) {
                p.it_merge[it] = 1;
                p.it_mroff[it][0] = roff[0][0];
                p.it_mroff[it][1] = roff[0][1];
            }
        }
    }
    p.vec8 = 1;
    for (int d = 0; d < 2; ++d)
        if (g.o[d].N && (((uintptr_t)g.o[d].p & 31) || g.o[d].sT % 8 || g.o[d].sF % 8 || g.o[d].sB % 8)) p.vec8 = 0;
    p.C0 = g.a[0].C; p.C1 = g.a[1].C; p.N = N; p.wJ_slabs = g.wJ_slabs;
    p.t_tiles = (g.Tout + TM - 1) / TM;
    p.n_tiles = N / BN;
    p.total_tiles = (long long)g.B * g.J * p.t_tiles * p.n_tiles;
    switch (BN) {
        case 256: return run<256>(g, p, st);
        case 128: return run<128>(g, p, st);
        case 64: return run<64>(g, p, st);
        default: return run<32>(g, p, st);
    }
}

static int g_engine = 1;   // 0: fp32 CUDA-core engine everywhere (exact; tests), 1: tcgen05 TF32 where eligible

int sefd_tapgemm(const TapGemmParams& p, cudaStream_t st) {
    SEFD_REQUIRE(p.nacc != 2, "tapgemm: two output rows per tile are a tensor-core-engine feature (use sefd_tapgemm_up)");
    if (sefd_skinny_conv_eligible(p)) return sefd_skinny_conv(p, st);
    if (g_engine == 1 && sefd_tapgemm_tc_eligible(p)) return sefd_tapgemm_tc(p, st);
    return sefd_tapgemm_simt(p, st);
}

int sefd_tapgemm_up(const TapGemmParams& g0, int mode, cudaStream_t st) {
    const int N = g0.o[0].N + g0.o[1].N;
    TapGemmParams f = g0;
    // fused: all 10 taps, tap_acc = output-row phase (kf & 1)
    f.ntaps = 0;
    for (int kf = 0; kf < 5; ++kf)
        for (int kt = 0; kt < 2; ++kt) {
            const int i = f.ntaps++;
            const int ph = kf & 1;
            f.df[i] = (2 + ph - kf) / 2;         // even: kf 0,2,4 -> +1,0,-1 ; odd: kf 1,3 -> +1,0
            f.dt[i] = mode == 0 ? -kt : 1 - kt;
            f.wslab[i] = kf * 2 + kt;
            f.tap_acc[i] = ph;
        }
    f.fi_mul = 1; f.fo_mul = 2; f.fo_off = 0; f.nacc = 2;
    if (g_engine == 1 && N % 32 == 0 && N <= SEFD_FUSE_UP_MAX_N && !sefd_skinny_conv_eligible(f) && sefd_tapgemm_tc_eligible(f))
        return sefd_tapgemm_tc(f, st);
    for (int ph = 0; ph < 2; ++ph) {
        TapGemmParams g = g0;
        g.nacc = 0;
        g.ntaps = 0;
        for (int kf = ph; kf < 5; kf += 2)
            for (int kt = 0; kt < 2; ++kt) {
                const int i = g.ntaps++;
                g.df[i] = (2 + ph - kf) / 2;
                g.dt[i] = mode == 0 ? -kt : 1 - kt;
                g.wslab[i] = kf * 2 + kt;
            }
        g.fi_mul = 1; g.fo_mul = 2; g.fo_off = ph;
        SEFD_TRY(sefd_tapgemm(g, st));
    }
    return 0;
}

extern "C" int sefd_set_engine(int engine) {
    SEFD_REQUIRE(engine == 0 || engine == 1, "set_engine: 0 (fp32 CUDA cores) or 1 (tcgen05 TF32)");
    g_engine = engine;
    return 0;
}
extern "C" int sefd_get_engine(void) { return g_engine; }
int sefd_get_engine_internal() { return g_engine; }
