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
constexpr int MAX_OPS = 10, MAX_TILES = 16, MAX_GROUPS = 5;
constexpr int SMEM_BUDGET = 225 * 1024;

// A tap group: per k-step the producer loads nA activation tiles and nG gradient tiles into consecutive slots of the
// stage (A slots a_sp chunks apart, then G slots BN/32 chunks apart).  MMA op i multiplies the A chunks starting at
// chunk op_a[i] with the G chunks starting at chunk op_g[i] into the TMEM columns [op_col[i], op_col[i] + op_n[i]).
// Taps are MERGED into one MMA where the operand layout allows it:
//   * N-merge (narrow N): taps that share the A tile and whose G tiles sit in consecutive slots become one MMA with
//     N = run x BN (the G tiles are successive 32-channel chunks of one MN-major operand);
//   * M-merge (K < 128): taps that share the G tile and whose A tiles sit in consecutive slots fill the 128 rows of
//     one MMA (row = tap_in_run * K + k) instead of leaving 128 - K rows of the tensor-core tile idle.
// The accumulator is described to the epilogue as BN-wide tiles: tile tt holds tile_nsub[tt] row blocks of K
// (or min(128, K - m0)) rows, block s belongs to weight slab tile_slab[tt][s].
struct Group {
    int nA, nG, nOps, nTiles;
    int a_roff[MAX_OPS], a_toff[MAX_OPS];    // A tile: row = j*a_mul + a_roff, time = t0 + a_toff
    int g_roff[MAX_OPS];                     // G tile: row = j*g_mul + g_roff, time = t0
    int op_a[MAX_OPS], op_g[MAX_OPS], op_n[MAX_OPS], op_col[MAX_OPS];
    int tile_nsub[MAX_TILES], tile_slab[MAX_TILES][4];
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
    int g_tiled;       // G operand is tile-major (see WgradParams)
    long long units;
};

struct Unit {
    int g, m0, n0, r0, r1, split;
};
__device__ __forceinline__ Unit decode_unit(const WgParams& p, long long u, int BN) {
    Unit x;
    // tile index fastest: the ~148 units in flight at any time are ALL (tap group, m, n) tiles of the same few row splits,
    // so the activation / gradient rows they stream are shared through L2 instead of being re-read from HBM once per tile
    // (split-fastest order measured 1.86x the algorithmic DRAM bytes)
    const long long tiles = (long long)p.ngroups * p.m_tiles * p.n_tiles;
    x.split = (int)(u / tiles);
    u %= tiles;
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
    // canonical (compiler-visible warp-uniform) warp index, see tapgemm_tc.cu
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

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
                            if (p.g_tiled) tma_load_5d(&tmG, fb, sg + (uint32_t)(g * NG * CHB), 0, t0 & 127, x.n0 / 32, t0 >> 7, fg);
                            else tma_load_5d(&tmG, fb, sg + (uint32_t)(g * NG * CHB), 0, t0, x.n0 / 32, fg, b);
                        }
                    }
                    if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform loops, elected lane issues) =================
        {
            // D fp32, A/B tf32, both MN-major (bits 15, 16), M = 128; N is set per op (merged taps)
            const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(WM >> 4) << 24);
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
                            const uint32_t abase = sa + (uint32_t)(G.op_a[o] * CHB);
                            const uint32_t gbase = sg + (uint32_t)(G.op_g[o] * CHB);
                            const uint32_t id = idesc0 | ((uint32_t)(G.op_n[o] >> 3) << 17);
#pragma unroll
                            for (int k8 = 0; k8 < PB / 8; ++k8) {
                                // MN-major tf32: SWIZZLE_128B_BASE32B (layout type 1), LBO = stride between 32-channel
                                // chunks, SBO = one 4-position swizzle atom (probed on hardware: tools/umma_probe.cu)
                                if (elect_one_sync())
                                    tc_mma_tf32(tmem_base + (uint32_t)G.op_col[o], make_desc_full(abase + k8 * 1024, CHB, 512, 1),
                                                make_desc_full(gbase + k8 * 1024, CHB, 512, 1), id, acc | (uint32_t)(k8 > 0));
                            }
                        }
                        acc = 1;
                        if (elect_one_sync()) tc_commit(smem_u32(&empty[stage]));
                        __syncwarp();
                        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                    }
                }
                if (elect_one_sync()) tc_commit(smem_u32(tfull));
                __syncwarp();
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
            // rows of an M-merged tile: row r = sub * K + k; otherwise r = k - m0
            const int rows_per_sub = p.K < WM ? p.K : WM;
            for (int tt = 0; tt < G.nTiles; ++tt) {
                float* tbase = p.out + x.split * p.split_stride + x.n0;
#pragma unroll 1
                for (int ch = 0; ch < NG; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tt * BN + ch * 32), v);
                    asm volatile("bar.sync 1, 128;" ::: "memory");      // previous chunk's readers are done
                    float* srow = stg + row * STG_LD;
#pragma unroll
                    for (int i = 0; i < 32; ++i) srow[i] = v[i];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                    for (int pass = 0; pass < 8; ++pass) {
                        const int idx = pass * 128 + et;
                        const int r = idx >> 3, c4 = (idx & 7) * 4;
                        const int sub = r / rows_per_sub;
                        const int k = x.m0 + r - sub * rows_per_sub;
                        if (sub < G.tile_nsub[tt] && k < p.K) {
                            const float* sp = stg + r * STG_LD + c4;
                            *reinterpret_cast<float4*>(tbase + ((long long)G.tile_slab[tt][sub] * p.K + k) * p.N + ch * 32 + c4) =
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

int g_cta_limit = 0;      // > 0: use at most this many CTAs (the caller keeps SMs free for a kernel on another stream)

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
    const int ctas = g_cta_limit > 0 && g_cta_limit < sms ? g_cta_limit : sms;
    const int grid = (int)(p.units < ctas ? p.units : ctas);
    wgrad_tc_kernel<BN><<<grid, NTHREADS, smem, st>>>(a0, a1, g, p);
    return sefd_check_launch("wgrad_tc");
}

// Partition the taps into MMA runs (see Group) and the runs into groups.  A group is grown run-by-run while its
// accumulator tiles fit TMEM (nTiles * BN <= 512 columns) and its stage fits the shared-memory budget.
struct Run {
    int n;                 // taps in the run
    int tap[8];
    int mode;              // 0: single / N-merge (shared A), 1: M-merge (shared G)
};

int build_groups(const WgradParams& w, int BN, int K, int nAc, WgParams& p) {
    const int NG = BN / 32;
    const int max_stage = 64 * 1024;
    const int m_merge = K < WM ? WM / K : 1;            // taps per MMA along M
    const int n_merge = BN <= 64 ? 256 / BN : 1;        // taps per MMA along N (wide tiles are efficient MMAs already)
    // ---- tap order (narrow N only): taps that share an A tile next to each other, ordered by G row, so that they
    // can be N-merged; otherwise the caller's order (kf-major: consecutive taps share the G tile) is kept ----
    int order[SEFD_MAX_TAPS], no = 0;
    if (n_merge > 1) {
        bool used[SEFD_MAX_TAPS] = {false};
        for (int i = 0; i < w.ntaps; ++i) {
            if (used[i]) continue;
            int same[SEFD_MAX_TAPS], ns = 0;
            for (int j2 = i; j2 < w.ntaps; ++j2)
                if (!used[j2] && w.a_off[j2] == w.a_off[i] && w.dt[j2] == w.dt[i]) { same[ns++] = j2; used[j2] = true; }
            for (int a = 0; a < ns; ++a)
                for (int b2 = a + 1; b2 < ns; ++b2)
                    if (w.g_off[same[b2]] < w.g_off[same[a]]) { const int t = same[a]; same[a] = same[b2]; same[b2] = t; }
            for (int a = 0; a < ns; ++a) order[no++] = same[a];
        }
    } else {
        for (int i = 0; i < w.ntaps; ++i) order[no++] = i;
    }
    // ---- runs ----
    Run runs[SEFD_MAX_TAPS];
    int nr = 0;
    for (int i = 0; i < no;) {
        Run r;
        memset(&r, 0, sizeof(r));
        r.tap[0] = order[i];
        r.n = 1;
        const int t0 = order[i];
        int j2 = i + 1;
        // N-merge: same A tile, G rows ascending by one
        while (j2 < no && r.n < n_merge && r.n < 8 && w.a_off[order[j2]] == w.a_off[t0] && w.dt[order[j2]] == w.dt[t0] &&
               w.g_off[order[j2]] == w.g_off[order[j2 - 1]] + 1) {
            r.tap[r.n++] = order[j2++];
        }
        if (r.n == 1 && m_merge > 1) {
            // M-merge: same G row, distinct A tiles (they get consecutive slots below)
            r.mode = 1;
            while (j2 < no && r.n < m_merge && r.n < 4 && w.g_off[order[j2]] == w.g_off[t0] &&
                   !(w.a_off[order[j2]] == w.a_off[order[j2 - 1]] && w.dt[order[j2]] == w.dt[order[j2 - 1]])) {
                r.tap[r.n++] = order[j2++];
            }
        }
        runs[nr++] = r;
        i = j2;
    }
    // runs that read the same G window next to each other (they then share the G slots of a group)
    if (n_merge > 1)
        for (int a = 1; a < nr; ++a)
            for (int b2 = a; b2 > 0 && w.g_off[runs[b2].tap[0]] < w.g_off[runs[b2 - 1].tap[0]]; --b2) {
                const Run t = runs[b2]; runs[b2] = runs[b2 - 1]; runs[b2 - 1] = t;
            }
    // ---- groups ----
    p.ngroups = 0;
    Group cur;
    memset(&cur, 0, sizeof(cur));
    auto add_run = [&](Group g, const Run& r, bool* ok) -> Group {
        *ok = true;
        int a_slot[8], g_slot[8];
        // G slots: an N-merged run needs r.n consecutive slots holding its rows in order - reuse such a window if the
        // group already has one, else append fresh slots; single taps / M-merged runs share any matching slot
        int q0 = -1;
        for (int q = 0; q + r.n <= g.nG && q0 < 0; ++q) {
            bool match = true;
            for (int i = 0; i < (r.mode == 0 ? r.n : 1); ++i) match = match && g.g_roff[q + i] == w.g_off[r.tap[i]];
            if (match) q0 = q;
        }
        if (r.mode == 1 && q0 < 0)
            for (int q = 0; q < g.nG; ++q)
                if (g.g_roff[q] == w.g_off[r.tap[0]]) q0 = q;
        if (q0 < 0) {
            const int need = r.mode == 0 ? r.n : 1;
            if (g.nG + need > MAX_OPS) { *ok = false; return g; }
            q0 = g.nG;
            for (int i = 0; i < need; ++i) g.g_roff[g.nG++] = w.g_off[r.tap[i]];
        }
        for (int i = 0; i < r.n; ++i) g_slot[i] = r.mode == 0 ? q0 + i : q0;
        for (int i = 0; i < r.n; ++i) {
            const int t = r.tap[i];
            int ia = -1;
            // an M-merged run needs its A tiles in fresh consecutive slots; otherwise slots are shared
            if (r.mode == 0)
                for (int a = 0; a < g.nA; ++a)
                    if (g.a_roff[a] == w.a_off[t] && g.a_toff[a] == w.dt[t]) ia = a;
            if (ia < 0) {
                if (g.nA >= MAX_OPS) { *ok = false; return g; }
                ia = g.nA; g.a_roff[ia] = w.a_off[t]; g.a_toff[ia] = w.dt[t]; ++g.nA;
            }
            a_slot[i] = ia;
        }
        (void)g_slot;
        if (g.nOps >= MAX_OPS) { *ok = false; return g; }
        const int o = g.nOps++;
        g.op_a[o] = a_slot[0] * nAc;
        g.op_g[o] = g_slot[0] * NG;
        g.op_col[o] = g.nTiles * BN;
        if (r.mode == 1) {
            if (g.nTiles + 1 > MAX_TILES) { *ok = false; return g; }
            g.op_n[o] = BN;
            g.tile_nsub[g.nTiles] = r.n;
            for (int i = 0; i < r.n; ++i) g.tile_slab[g.nTiles][i] = w.wslab[r.tap[i]];
            ++g.nTiles;
        } else {
            if (g.nTiles + r.n > MAX_TILES) { *ok = false; return g; }
            g.op_n[o] = r.n * BN;
            for (int i = 0; i < r.n; ++i) {
                g.tile_nsub[g.nTiles] = 1;
                g.tile_slab[g.nTiles][0] = w.wslab[r.tap[i]];
                ++g.nTiles;
            }
        }
        return g;
    };
    for (int i = 0; i < nr; ++i) {
        bool ok;
        Group trial = add_run(cur, runs[i], &ok);
        const int bytes = (trial.nA * nAc + trial.nG * NG) * CHB;
        if (!ok || trial.nTiles * BN > 512 || bytes > max_stage) {
            if (!cur.nOps) return -1;                   // a single run does not fit
            if (p.ngroups >= MAX_GROUPS) return -1;
            p.grp[p.ngroups++] = cur;
            memset(&cur, 0, sizeof(cur));
            --i;                                        // retry this run in a fresh group
            continue;
        }
        cur = trial;
    }
    if (cur.nOps) {
        if (p.ngroups >= MAX_GROUPS) return -1;
        p.grp[p.ngroups++] = cur;
    }
    return 0;
}

}  // namespace

void sefd_wgrad_tc_set_cta_limit(int n) { g_cta_limit = n; }

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
    SEFD_REQUIRE(build_groups(w, BN, K, p.a_sp, p) == 0, "wgrad_tc: tap grouping failed (K %d N %d BN %d taps %d)", K, N, BN, w.ntaps);
    int stage = 0;
    for (int g = 0; g < p.ngroups; ++g) {
        const int b = (p.grp[g].nA * p.a_sp + p.grp[g].nG * (BN / 32)) * CHB +
                      (K < WM ? 3 * CHB : 0);   // slack only when an A slot is narrower than the 4 chunks an MMA reads
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
    // units = tiles x splits are dealt round-robin to the 148 persistent CTAs: aim for just UNDER two full rounds
    // (300 units made three rounds, the third 3 % full: 2.03 rounds of work in the time of 3)
    const int ctas = g_cta_limit > 0 && g_cta_limit < 148 ? g_cta_limit : 148;
    long long splits = (2 * ctas) / tiles;
    if (splits < 1) splits = 1;
    if (splits > rows) splits = rows;
    if (splits > cap_floats / one) splits = cap_floats / one;
    if (splits > 2 * ctas) splits = 2 * ctas;
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
    if (w.g_tiled) {
        SEFD_REQUIRE(w.B == 1, "wgrad_tc: a tile-major gradient operand needs B = 1");
        const int mt = (w.Tg + 127) / 128, kbn = N / 32;
        cuuint64_t dims[5] = {32, 128, (cuuint64_t)kbn, (cuuint64_t)mt, (cuuint64_t)w.Fg};
        cuuint64_t str[4] = {128, 16384, (cuuint64_t)kbn * 16384, (cuuint64_t)mt * kbn * 16384};
        cuuint32_t box[5] = {32, (cuuint32_t)PB, (cuuint32_t)(BN / 32), 1, 1};
        SEFD_TRY(make_map(&g, w.g.p, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
        p.g_tiled = 1;
    } else {
        SEFD_TRY(make_pos_map(&g, w.g, w.Fg, w.Tg, w.B, BN / 32));
    }
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
    SEFD_REQUIRE(!w.g_tiled, "wgrad: a tile-major gradient operand needs the tensor-core engine");
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
