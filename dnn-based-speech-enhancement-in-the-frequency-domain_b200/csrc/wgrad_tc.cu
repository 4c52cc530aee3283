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
constexpr int MAX_OPS = 10, MAX_GROUPS = 5;
constexpr int SMEM_BUDGET = 225 * 1024;

// A tap group: per k-step the producer loads nA activation tiles and nG gradient tiles; op i multiplies
// A slot op_a[i] with G slot op_g[i] into TMEM accumulator i (BN columns each) and ends up in slab op_slab[i].
struct Group {
    int nA, nG, nOps;
    int a_roff[MAX_OPS], a_toff[MAX_OPS];    // A tile: row = j*a_mul + a_roff, time = t0 + a_toff
    int g_roff[MAX_OPS];                     // G tile: row = j*g_mul + g_roff, time = t0
    int op_a[MAX_OPS], op_g[MAX_OPS], op_slab[MAX_OPS];
};

struct WgParams {
    float* out;
    long long split_stride;
    int B, J, a_mul, g_mul;
    int C0, C1, K, N;
    int ngroups;
    Group grp[MAX_GROUPS];
    int m_tiles, n_tiles, splits, rows_per_split, t_blocks;
    int nstage, stage_bytes;
    int a_sp;          // spacing of the A slots inside a stage, in 4 KB chunks (= min(4, K/32))
    int a_box;         // 32-channel blocks fetched by one A TMA instruction (divides a_sp and the source widths)
    long long units;
};

struct Unit {
    int g, m0, n0, r0, r1, split;
};
__device__ __forceinline__ Unit decode_unit(const WgParams& p, long long u, int BN) {
    Unit x;
    x.split = (int)(u % p.splits);
    u /= p.splits;
    x.n0 = (int)(u % p.n_tiles) * BN;
    u /= p.n_tiles;
    x.m0 = (int)(u % p.m_tiles) * WM;
    x.g = (int)(u / p.m_tiles);
    x.r0 = x.split * p.rows_per_split;
    const int rows = p.B * p.J;
    x.r1 = x.r0 + p.rows_per_split < rows ? x.r0 + p.rows_per_split : rows;
    return x;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmG, const WgParams p) {
    constexpr int NG = BN / 32;                 // 32-channel chunks of one G tile
    // declared alignment keeps derived pointers in the shared address space (no generic LD/ST in the epilogue)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stages = smem;
    float* stg = reinterpret_cast<float*>(smem + p.nstage * p.stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + WM * STG_LD);
    uint64_t* full = bars;
    uint64_t* empty = bars + 8;
    uint64_t* tfull = bars + 16;
    uint64_t* tempty = bars + 17;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < p.nstage; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        mbar_init(smem_u32(tfull), 1);
        mbar_init(smem_u32(tempty), 4);
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
        // ================= TMA producer: the whole warp issues, one box per lane =================
        int stage = 0;
        uint32_t phase = 0;
        for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
            const Unit x = decode_unit(p, u, BN);
            const Group& G = p.grp[x.g];
            const int opa = p.a_sp / p.a_box;                                // TMA instructions per A tile
            const int nAb = G.nA * opa, nGb = G.nG;                          // TMA instructions per k-step
            const uint32_t bytes = (uint32_t)((G.nA * p.a_sp + G.nG * NG) * CHB);
            for (int r = x.r0; r < x.r1; ++r) {
                const int b = r / p.J, j = r % p.J;
                for (int tb = 0; tb < p.t_blocks; ++tb) {
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full[stage]);
                    const uint32_t sa = smem_u32(stages + stage * p.stage_bytes);
                    if (lane == 0) mbar_expect_tx(fb, bytes);
                    __syncwarp();
                    const int t0 = tb * PB;
                    for (int op = lane; op < nAb + nGb; op += 32) {
                        if (op < nAb) {
                            const int a = op / opa, i = (op % opa) * p.a_box;          // first 32-channel block of this box
                            const int fa = j * p.a_mul + G.a_roff[a], ta = t0 + G.a_toff[a];
                            const int kc = x.m0 + 32 * i;
                            const uint32_t dst = sa + (uint32_t)((a * p.a_sp + i) * CHB);
                            // blocks beyond the tensor (last m-tile of a narrow K) are zero-filled by TMA
                            if (kc < p.C0) tma_load_5d(&tmA0, fb, dst, 0, ta, kc / 32, fa, b);
                            else tma_load_5d(&tmA1, fb, dst, 0, ta, (kc - p.C0) / 32, fa, b);
                        } else {
                            const int g = op - nAb;
                            const int fg = j * p.g_mul + G.g_roff[g];
                            const uint32_t sg = sa + (uint32_t)(G.nA * p.a_sp * CHB);
                            tma_load_5d(&tmG, fb, sg + (uint32_t)(g * NG * CHB), 0, t0, x.n0 / 32, fg, b);
                        }
                    }
                    if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // D fp32, A/B tf32, both MN-major (bits 15, 16), N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
                const Unit x = decode_unit(p, u, BN);
                const Group& G = p.grp[x.g];
                mbar_wait(smem_u32(tempty), aphase ^ 1);
                tc_fence_after();
                uint32_t acc = 0;
                for (int r = x.r0; r < x.r1; ++r) {
                    for (int tb = 0; tb < p.t_blocks; ++tb) {
                        mbar_wait(smem_u32(&full[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stages + stage * p.stage_bytes);
                        const uint32_t sg = sa + (uint32_t)(G.nA * p.a_sp * CHB);
                        for (int o = 0; o < G.nOps; ++o) {
                            const uint32_t abase = sa + (uint32_t)(G.op_a[o] * p.a_sp * CHB);
                            const uint32_t gbase = sg + (uint32_t)(G.op_g[o] * NG * CHB);
#pragma unroll
                            for (int k8 = 0; k8 < PB / 8; ++k8) {
                                // MN-major tf32: SWIZZLE_128B_BASE32B (layout type 1), LBO = stride between 32-channel
                                // chunks, SBO = one 4-position swizzle atom (probed on hardware: tools/umma_probe.cu)
                                tc_mma_tf32(tmem_base + (uint32_t)(o * BN), make_desc_full(abase + k8 * 1024, CHB, 512, 1),
                                            make_desc_full(gbase + k8 * 1024, CHB, 512, 1), idesc, acc | (uint32_t)(k8 > 0));
                            }
                        }
                        acc = 1;
                        tc_commit(smem_u32(&empty[stage]));
                        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                    }
                }
                tc_commit(smem_u32(tfull));
                aphase ^= 1;
            }
        }
    } else {
        // ================= epilogue: TMEM -> padded smem -> coalesced rows of the partial buffer =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;
        uint32_t aphase = 0;
        for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
            const Unit x = decode_unit(p, u, BN);
            const Group& G = p.grp[x.g];
            mbar_wait(smem_u32(tfull), aphase);
            tc_fence_after();
            for (int o = 0; o < G.nOps; ++o) {
                float* obase = p.out + x.split * p.split_stride + (long long)G.op_slab[o] * p.K * p.N + x.n0;
#pragma unroll 1
                for (int ch = 0; ch < NG; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(o * BN + ch * 32), v);
                    asm volatile("bar.sync 1, 128;" ::: "memory");      // previous chunk's readers are done
                    float* srow = stg + row * STG_LD;
#pragma unroll
                    for (int i = 0; i < 32; ++i) srow[i] = v[i];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                    for (int pass = 0; pass < 8; ++pass) {
                        const int idx = pass * 128 + et;
                        const int r = idx >> 3, c4 = (idx & 7) * 4;
                        const int k = x.m0 + r;
                        if (k < p.K) {
                            const float* sp = stg + r * STG_LD + c4;
                            *reinterpret_cast<float4*>(obase + (long long)k * p.N + ch * 32 + c4) =
                                make_float4(sp[0], sp[1], sp[2], sp[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty));
            aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int make_pos_map(CUtensorMap* m, const TapSrc& s, int F, int T, int B, int nblk) {
    return make_act_map5(m, s, F, T, B, PB, nblk, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

template <int BN>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& g, const WgParams& p, int smem, cudaStream_t st) {
    static int cur = 0;
    if (smem > cur) {
        cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cur = smem;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(p.units < sms ? p.units : sms);
    wgrad_tc_kernel<BN><<<grid, NTHREADS, smem, st>>>(a0, a1, g, p);
    return sefd_check_launch("wgrad_tc");
}

// Partition the taps into groups that share operand tiles.  Taps with the same A tile (a_off, dt) share an A slot,
// taps with the same G row share a G slot.  A group is grown tap-by-tap (in kf order) while its accumulators fit
// TMEM (nOps * BN <= 512 columns) and its stage fits the shared-memory budget.
int build_groups(const WgradParams& w, int BN, int nAc, WgParams& p) {
    const int NG = BN / 32;
    const int max_ops = 512 / BN < MAX_OPS ? 512 / BN : MAX_OPS;
    const int max_stage = 64 * 1024;
    p.ngroups = 0;
    Group cur;
    memset(&cur, 0, sizeof(cur));
    auto flush = [&]() {
        if (cur.nOps) p.grp[p.ngroups++] = cur;
        memset(&cur, 0, sizeof(cur));
    };
    for (int t = 0; t < w.ntaps; ++t) {
        // find or add slots
        Group trial = cur;
        int ia = -1, ig = -1;
        for (int a = 0; a < trial.nA; ++a)
            if (trial.a_roff[a] == w.a_off[t] && trial.a_toff[a] == w.dt[t]) ia = a;
        if (ia < 0) { ia = trial.nA; trial.a_roff[ia] = w.a_off[t]; trial.a_toff[ia] = w.dt[t]; ++trial.nA; }
        for (int g = 0; g < trial.nG; ++g)
            if (trial.g_roff[g] == w.g_off[t]) ig = g;
        if (ig < 0) { ig = trial.nG; trial.g_roff[ig] = w.g_off[t]; ++trial.nG; }
        trial.op_a[trial.nOps] = ia; trial.op_g[trial.nOps] = ig; trial.op_slab[trial.nOps] = w.wslab[t];
        ++trial.nOps;
        const int bytes = (trial.nA * nAc + trial.nG * NG) * CHB;   // A slots are spaced nAc chunks apart
        if (cur.nOps && (trial.nOps > max_ops || bytes > max_stage)) {
            flush();
            --t;            // retry this tap in a fresh group
            continue;
        }
        cur = trial;
        if (p.ngroups >= MAX_GROUPS) return -1;
    }
    flush();
    return p.ngroups <= MAX_GROUPS ? 0 : -1;
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
    p.B = w.B; p.J = w.J; p.a_mul = w.a_mul; p.g_mul = w.g_mul;
    p.C0 = w.a[0].C; p.C1 = w.a[1].C; p.K = K; p.N = N;
    p.m_tiles = (K + WM - 1) / WM;
    p.n_tiles = N / BN;
    p.t_blocks = (w.Tg + PB - 1) / PB;
    p.a_sp = K / 32 < 4 ? K / 32 : 4;
    p.a_box = p.a_sp;
    if (w.a[1].C) {
        while ((w.a[0].C / 32) % p.a_box || (w.a[1].C / 32) % p.a_box) p.a_box >>= 1;
    }
    SEFD_REQUIRE(build_groups(w, BN, p.a_sp, p) == 0, "wgrad_tc: tap grouping failed");
    int stage = 0;
    for (int g = 0; g < p.ngroups; ++g) {
        const int b = (p.grp[g].nA * p.a_sp + p.grp[g].nG * (BN / 32)) * CHB + 3 * CHB;   // + slack: the MMA always reads 4 chunks
        if (b > stage) stage = b;
    }
    p.stage_bytes = stage;
    const int fixed = WM * STG_LD * 4 + 20 * 8 + 16 + 1024;
    p.nstage = (SMEM_BUDGET - fixed) / stage;
    if (p.nstage > 8) p.nstage = 8;
    SEFD_REQUIRE(p.nstage >= 2, "wgrad_tc: stage of %d bytes leaves no room for a pipeline", stage);
    const int smem = p.nstage * stage + fixed;

    const int rows = w.B * w.J;
    const long long tiles = (long long)p.ngroups * p.m_tiles * p.n_tiles;
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
    SEFD_TRY(make_pos_map(&a0, w.a[0], w.Fa, w.Ta, w.B, p.a_box));
    if (w.a[1].C) SEFD_TRY(make_pos_map(&a1, w.a[1], w.Fa, w.Ta, w.B, p.a_box));
    else a1 = a0;
    SEFD_TRY(make_pos_map(&g, w.g, w.Fg, w.Tg, w.B, BN / 32));
    const double pos = (double)w.B * w.J * w.Tg;
    sefd_prof_label("wgrad_tc BN%d K%d N%d taps%d J%d groups%d stage%dK x%d splits%d units%lld", BN, K, N, w.ntaps, w.J,
                    p.ngroups, stage / 1024, p.nstage, p.splits, p.units);
    SefdProfScope prof(SEFD_PROF_WGRAD, 2.0 * pos * K * N * w.ntaps,
                       4.0 * ((double)w.B * w.J * (w.a_mul > 1 ? w.a_mul : 1) * w.Ta * K +
                              (double)w.B * w.J * (w.g_mul > 1 ? w.g_mul : 1) * w.Tg * N), st);
    switch (BN) {
        case 256: return launch<256>(a0, a1, g, p, smem, st);
        case 128: return launch<128>(a0, a1, g, p, smem, st);
        case 64: return launch<64>(a0, a1, g, p, smem, st);
        default: return launch<32>(a0, a1, g, p, smem, st);
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
