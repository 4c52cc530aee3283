// Fused tcgen05 LSTM recurrence for the time-major layer engine (lstm_seq.cuh): nn.LSTM inside SequenceModel
// (tools_for_model.py:741-748) and cfg.lstm = 'real' (models.py:96-105), gate order i, f, g, o.
//
// One CTA owns 128 consecutive sequences ("rows") for a RANGE of time steps; sequences are independent, so one launch runs
// the whole sequence: the only cross-step dependency is h_t (forward) / dG_{t-1} (backward), which the epilogue warps write
// to HBM / L2 and the TMA warp reads back for the next step (fence.proxy.async + mbarrier hand-over inside the kernel).
//
// Forward, per step and CTA:   G[128 x 4H] = [x_t | h_{t-1}] [128 x (I + H)] . Wcat^T   (TF32 tcgen05.mma, fp32 TMEM)
//   in chunks of 256 gate columns = i, f, g, o of 64 hidden units (interleaved gate layout), two TMEM accumulators so that
//   the LSTM cell of chunk q (8 epilogue warps: sigmoid / tanh, c_t, h_t, activated gates for the backward) overlaps the
//   MMAs of chunk q + 1.  No pre-activation tensor exists in HBM: x_t and h_{t-1} are the two K sources of one GEMM; the
//   x part of step t + 1 is issued before the hand-over wait, so it overlaps the last cell epilogue of step t.
// Backward, per step and CTA:  [dh_{t-1} | dx_t] [128 x (H + I)] = dG_t [128 x 4H] . [W_hh^T ; W_ih^T]
//   the epilogue turns the dh columns into dG_{t-1} with the cell backward (in place over the saved gates) - the A operand
//   of the next (earlier) step - and stores the dx columns, whose MMAs come last and overlap the hand-over; per-column sums
//   of dG (bias gradient) are reduced with a transposing shuffle butterfly into per-CTA slots (deterministic).
//
// Two cluster shapes:
//   CLM (many rows, FullSubNet's sub-band model): every CTA needs ALL weights every step (2.4 - 4.7 MB); per-CTA copies would
//     hit the chip-wide L2 -> SM bandwidth (~11 TB/s measured with tapgemm_tc) at ~45 % of the tensor rate.  CLM CTAs with
//     different row tiles walk the chunks in lockstep, each fetches 1 / CLM of every weight tile and TMA-multicasts it into
//     the shared memory of all of them (empty[] barriers count CLM consumers, tcgen05.commit multicasts the release).
//   CLN (<= 128 rows: FullSubNet's full-band model, lstm = 'real'): the chunks of ONE row tile are split over CLN CTAs, so a
//     step costs one chunk of MMA time instead of H / 64; h_t / dG_{t-1} are exchanged through L2 with a cluster-scope
//     mbarrier hand-over (remote arrive on every peer's barrier) once per step.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, then 8 (forward) / 16 (backward) epilogue warps (thread = sequence row,
// the warps that share a TMEM lane quarter split the chunk's columns).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "lstm_step_tc.cuh"
#include "prof.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128, KB = 32;
constexpr int NEPI_F = 8, NEPI_B = 12;          // epilogue warps: forward (MUFU-bound), backward (global-load-latency-bound: more warps in flight)
constexpr int A_BYTES = TM * KB * 4;            // 16 KB: 128 rows x 32 k
constexpr int MAXN = 256;                       // widest chunk (accumulator columns)
constexpr int W_BYTES = MAXN * KB * 4;          // 32 KB slot for a chunk's weight tile (narrower chunks use a prefix)
constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
constexpr int NSTAGE = 4;
constexpr int MAX_CHUNKS = 16;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void tma_load_3d_multicast(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned long long spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1ull << 24)) __trap();
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpf(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - 2.f * rcpf(ex2f(fminf(x, 10.f) * (2.f * LOG2E)) + 1.f); }

__device__ __forceinline__ void st_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_v8(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}

struct StepParams {
    int rows, T, t0, t1, I, H;  // steps [t0, t1): forward ascending, backward descending
    int nk;                     // k-blocks per chunk
    int kx;                     // forward: k-blocks that come from x (the rest from h)
    int nchunks;
    int chunk_n[MAX_CHUNKS];    // accumulator width of each chunk
    int chunk_row[MAX_CHUNKS];  // first row of the chunk in the weight operand (= first output column)
    int chunk_map[MAX_CHUNKS];  // which weight tensor map serves the chunk (box height differs for a narrower last chunk)
    int persist;                // the launch covers more than one step: hand h_t / dG_{t-1} over inside the kernel
    int round_tf32;
    float* gates;               // [T][rows][4H']
    float* hbuf;                // forward: [T + 1][rows][H], slot 0 = zeros, h_t at slot t + 1
    float* c;                   // [T][rows][H]
    const float* bias;          // forward: [4H']
    const float* dh_out;        // backward: [T][rows][H]
    float* dc;                  // backward: [rows][H] carried cell-state gradient
    float* dx;                  // backward: [T][rows][I] or null
    float* bias_part;           // backward: [grid][4H'] per-CTA column sums of dG
    int bias_accum;             // add to the slot instead of overwriting it (one-launch-per-step debugging mode)
};

struct Smem {
    uint64_t full[NSTAGE], empty[NSTAGE], tfull[2], tempty[2], ready;
    uint32_t tmem_ptr;
};

// ---- shared set-up / tear-down -----------------------------------------------------------------------------------------
template <int CLM, int CLN, int NE>
__device__ __forceinline__ uint32_t step_setup(unsigned char* smem, Smem* sb, int warp) {
    constexpr int CL = CLM * CLN;
    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&sb->full[s]), 1);
            mbar_init(smem_u32(&sb->empty[s]), CLM);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&sb->tfull[a]), 1);
            mbar_init(smem_u32(&sb->tempty[a]), NE);
        }
        mbar_init(smem_u32(&sb->ready), NE * CLN);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // every CTA's barriers exist before a peer multicasts into them / arrives on them
    tc_fence_after();
    return sb->tmem_ptr;
}
template <int CL>
__device__ __forceinline__ void step_teardown(uint32_t tmem_base, int warp) {
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // no CTA leaves while a peer may still write its shared memory / barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// the MMA issuer is the same in both directions: for every (step, own chunk) nk k-blocks into one TMEM accumulator
template <int CLM, int CLN>
__device__ __forceinline__ void mma_role(const StepParams& p, Smem* sb, uint32_t smem_base, uint32_t tmem_base, int nrank, int nsteps) {
    const uint16_t mc_mask = (uint16_t)((1u << CLM) - 1);
    int stage = 0, abuf = 0;
    uint32_t phase = 0, aphase = 0;
    for (int it = 0; it < nsteps; ++it) {
        for (int q = nrank; q < p.nchunks; q += CLN) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.chunk_n[q] >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            mbar_wait(smem_u32(&sb->tempty[abuf]), aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(abuf * MAXN);
            for (int kb = 0; kb < p.nk; ++kb) {
                mbar_wait(smem_u32(&sb->full[stage]), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + (uint32_t)(stage * STAGE_BYTES);
                const uint64_t ad = make_desc(sa), bd = make_desc(sa + A_BYTES);
#pragma unroll
                for (int k8 = 0; k8 < KB / 8; ++k8)
                    if (elect_one_sync()) tc_mma_tf32(d_tmem, ad + 2 * k8, bd + 2 * k8, idesc, (uint32_t)((kb | k8) != 0));
                if (elect_one_sync()) {
                    if (CLM > 1) tc_commit_multicast(smem_u32(&sb->empty[stage]), mc_mask);
                    else tc_commit(smem_u32(&sb->empty[stage]));
                }
                __syncwarp();
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) tc_commit(smem_u32(&sb->tfull[abuf]));
            __syncwarp();
            if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
    }
}

// epilogue -> producer hand-over of the step's output rows (generic-proxy global writes -> async-proxy TMA reads)
template <int CLN>
__device__ __forceinline__ void publish_step(Smem* sb, int lane) {
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        if (CLN > 1) {
#pragma unroll
            for (int c = 0; c < CLN; ++c) mbar_arrive_remote(smem_u32(&sb->ready), (uint32_t)c);
        } else {
            mbar_arrive(smem_u32(&sb->ready));
        }
    }
}
template <int CLN>
__device__ __forceinline__ void await_step(Smem* sb, uint32_t& phase) {
    if (CLN > 1) mbar_wait_cluster(smem_u32(&sb->ready), phase);
    else mbar_wait(smem_u32(&sb->ready), phase);
    phase ^= 1;
    asm volatile("fence.proxy.async;" ::: "memory");
}

// =====================================================================================================================
// forward
// =====================================================================================================================
template <int CLM, int CLN>
__global__ void __launch_bounds__(64 + 32 * NEPI_F, 1)
lstm_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2, const StepParams p) {
    constexpr int CL = CLM * CLN;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_bias = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);          // [4H'] <= 2048 floats
    Smem* sb = reinterpret_cast<Smem*>(s_bias + 2048);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
    const int mrank = CLM > 1 ? (int)rank : 0, nrank = CLN > 1 ? (int)rank : 0;
    const int row0 = (CLN > 1 ? blockIdx.x / CLN : blockIdx.x) * TM;
    for (int i = threadIdx.x; i < 4 * p.H; i += 64 + 32 * NEPI_F) s_bias[i] = p.bias[i];
    const uint32_t tmem_base = step_setup<CLM, CLN, NEPI_F>(smem, sb, warp);
    const uint32_t smem_base = smem_u32(smem);
    const uint16_t mc_mask = (uint16_t)((1u << CLM) - 1);

    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0;
        uint32_t phase = 0, rphase = 0;
        for (int t = p.t0; t < p.t1; ++t) {
            bool first_h = p.persist && t > p.t0;
            for (int q = nrank; q < p.nchunks; q += CLN) {
                const int wslice = p.chunk_n[q] / CLM;
                const uint32_t bytes = (uint32_t)(A_BYTES + p.chunk_n[q] * KB * 4);
                for (int kb = 0; kb < p.nk; ++kb) {
                    if (first_h && kb == p.kx) {            // h_{t-1} of this row tile is complete (all chunks, all peers)
                        await_step<CLN>(sb, rphase);
                        first_h = false;
                    }
                    mbar_wait(smem_u32(&sb->empty[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&sb->full[stage]);
                    const uint32_t sa = smem_base + (uint32_t)(stage * STAGE_BYTES);
                    if (elect_one_sync()) {
                        mbar_expect_tx(fb, bytes);
                        if (kb < p.kx) tma_load_4d(&tmX, fb, sa, 0, row0, kb, t);
                        else tma_load_4d(&tmH, fb, sa, 0, row0, kb - p.kx, t);                   // slot t = h_{t-1}
                        const uint32_t sw = sa + A_BYTES + (uint32_t)(mrank * wslice * KB * 4);
                        if (CLM > 1) tma_load_3d_multicast(&tmW, fb, sw, 0, p.chunk_row[q] + mrank * wslice, kb, mc_mask);
                        else tma_load_3d(&tmW, fb, sw, 0, p.chunk_row[q], kb);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        mma_role<CLM, CLN>(p, sb, smem_base, tmem_base, nrank, p.t1 - p.t0);
    } else {
        // ================= epilogue: LSTM cell (thread = row; the two warps of a lane quarter split the chunk's units) ======
        const int qd = warp & 3, half = (warp - 2) >> 2;
        const int rt = qd * 32 + lane;                     // row inside the tile
        const int row = row0 + rt;
        const bool rv = row < p.rows;
        const int H = p.H, UB = H >> 3;
        const long long mt = row0 / TM, MT = (p.rows + TM - 1) / TM;
        const bool tile_ok = mt < MT;                      // a cluster may carry CTAs without a row tile: they only run the protocol
        int abuf = 0;
        uint32_t aphase = 0;
        for (int t = p.t0; t < p.t1; ++t) {
            // tile-major state: gates [T][MT][UB][128][32], c [T][MT][UB][128][8]; h row-major [T + 1][rows][H]
            float* gt = p.gates + (((long long)t * MT + mt) * UB * TM + rt) * 32;
            const float* cp = p.c + (((long long)(t - 1) * MT + mt) * UB * TM + rt) * 8;       // read only when t > 0
            float* ct = p.c + (((long long)t * MT + mt) * UB * TM + rt) * 8;
            float* ht = p.hbuf + ((long long)(t + 1) * p.rows + row) * H;
            for (int q = nrank; q < p.nchunks; q += CLN) {
                float cv[4][8];
#pragma unroll
                for (int b = 0; b < 4; ++b) {              // the c_{t-1} blocks are in flight before the accumulator is ready
                    const int ub = q * 8 + 4 * half + b;
                    if (t > 0 && tile_ok) ld_v8(cp + (long long)ub * TM * 8, cv[b]);
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) cv[b][i] = 0.f;
                    }
                }
                mbar_wait(smem_u32(&sb->tfull[abuf]), aphase);
                tc_fence_after();
                const uint32_t tacc = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(abuf * MAXN);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int ubl = 4 * half + b, ub = q * 8 + ubl;
                    float v[32];                           // [gate][unit] of this 8-unit block
                    tmem_ld32(tacc + (uint32_t)(ubl * 32), v);
                    if (b == 3) {                          // this warp is done with the accumulator
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&sb->tempty[abuf]));
                    }
                    const float* bq = s_bias + ub * 32;
                    float hv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // sigmoid(x) = 1 / (1 + 2^(-x log2 e)), tanh(x) = 1 - 2 / (1 + 2^(2 x log2 e)); the four gate
                        // reciprocals share ONE rcp (arguments clamped so that the product of the four denominators stays
                        // finite: sigmoid(-20) = 2e-9, 1 - tanh(10) = 4e-9): 7 MUFU per hidden unit instead of 10
                        const float A = 1.f + ex2f(fminf(-(v[i] + bq[i]), 20.f) * LOG2E);
                        const float Bf = 1.f + ex2f(fminf(-(v[8 + i] + bq[8 + i]), 20.f) * LOG2E);
                        const float C = 1.f + ex2f(fminf(v[16 + i] + bq[16 + i], 10.f) * (2.f * LOG2E));
                        const float D = 1.f + ex2f(fminf(-(v[24 + i] + bq[24 + i]), 20.f) * LOG2E);
                        const float AB = A * Bf, CD = C * D;
                        const float r = rcpf(AB * CD);
                        const float rab = r * CD, rcd = r * AB;          // 1 / (A B), 1 / (C D)
                        const float a = rab * Bf, f = rab * A, o = rcd * C;
                        const float g = 1.f - 2.f * (rcd * D);
                        const float c = fmaf(f, cv[b][i], a * g);
                        float h = o * tanh_fast(c);
                        if (p.round_tf32) h = tf32_rn(h);
                        v[i] = a; v[8 + i] = f; v[16 + i] = g; v[24 + i] = o; cv[b][i] = c; hv[i] = h;
                    }
                    // pad rows of the last tile hold finite values (zero inputs): the tile-major stores need no guard
                    if (tile_ok) {
                        float* gq = gt + (long long)ub * TM * 32;
                        st_v8(gq, v); st_v8(gq + 8, v + 8); st_v8(gq + 16, v + 16); st_v8(gq + 24, v + 24);
                        st_v8(ct + (long long)ub * TM * 8, cv[b]);
                    }
                    if (rv) st_v8(ht + ub * 8, hv);
                }
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
            }
            if (p.persist && t + 1 < p.t1) publish_step<CLN>(sb, lane);
        }
    }
    step_teardown<CL>(tmem_base, warp);
}

// =====================================================================================================================
// backward
// =====================================================================================================================
// cell backward of one 8-unit block of one row (thread = row): consumes dh (recurrent part in d[], zeros for the last step),
// reads the tile-major gates / c / dc of step tc and the row-major dh_out, writes dG in place over the gates and the carried
// dc; v[32] = [gate][unit] gradient values (zeros for rows beyond the end) for the bias column sums
struct CellCtx {
    float* gates_t;        // tile-major base of step tc for this (tile, row): + ub * 128 * 32
    const float* c_t;      // + ub * 128 * 8
    const float* c_tm1;    // null when tc == 0
    const float* dh_out;   // row-major row of step tc (+ u0)
    float* dc;             // + ub * 128 * 8
    bool rv, first;        // first: step T - 1 (no carried dc yet)
    int round_tf32;
};
__device__ __forceinline__ void cell_bwd_block(const CellCtx& x, int ub, const float* d, float* v) {
    float gi[8], gf[8], gg[8], go[8], cc[8], cpv[8], dh[8], dcn[8];
    float* gq = x.gates_t + (long long)ub * TM * 32;
    ld_v8(gq, gi); ld_v8(gq + 8, gf); ld_v8(gq + 16, gg); ld_v8(gq + 24, go);
    ld_v8(x.c_t + (long long)ub * TM * 8, cc);
    if (x.c_tm1) ld_v8(x.c_tm1 + (long long)ub * TM * 8, cpv);
    if (!x.first) ld_v8(x.dc + (long long)ub * TM * 8, dcn);
    if (x.rv) ld_v8(x.dh_out + ub * 8, dh);
    float dcc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float cprev = x.c_tm1 ? cpv[i] : 0.f;
        const float dht = x.rv ? dh[i] + d[i] : 0.f;
        const float tc = tanh_fast(cc[i]);
        const float dcv = fmaf(dht * go[i], 1.f - tc * tc, x.first ? 0.f : dcn[i]);
        v[i] = dcv * gg[i] * gi[i] * (1.f - gi[i]);
        v[8 + i] = dcv * cprev * gf[i] * (1.f - gf[i]);
        v[16 + i] = dcv * gi[i] * (1.f - gg[i] * gg[i]);
        v[24 + i] = dht * tc * go[i] * (1.f - go[i]);
        dcc[i] = dcv * gf[i];
    }
    if (!x.rv) {                   // pad rows: zeros (they are read as GEMM / weight-gradient operands)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) dcc[i] = 0.f;
    }
    float o[8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = x.round_tf32 ? tf32_rn(v[8 * g + i]) : v[8 * g + i];
        st_v8(gq + 8 * g, o);
    }
    st_v8(x.dc + (long long)ub * TM * 8, dcc);
}
// column sums of v[32] over the warp's 32 rows (transposing butterfly: lane c ends with value index c) into the slot the
// lane owns: s_bsum[quarter][ub * 32 + lane]
__device__ __forceinline__ void bias_accumulate(float* v, float* bs, int ub, int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float sv = up ? v[i] : v[i + off], kv = up ? v[i + off] : v[i];
            v[i] = kv + __shfl_xor_sync(0xffffffffu, sv, off);
        }
    }
    bs[ub * 32 + lane] += v[0];
}

template <int CLM, int CLN>
__global__ void __launch_bounds__(64 + 32 * NEPI_B, 1)
lstm_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmG_unused,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2, const StepParams p) {
    constexpr int CL = CLM * CLN;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_bsum = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);           // [4 quarters][4H'] column sums of dG
    Smem* sb = reinterpret_cast<Smem*>(s_bsum + 4 * 4 * p.H);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
    const int mrank = CLM > 1 ? (int)rank : 0, nrank = CLN > 1 ? (int)rank : 0;
    const int mt = CLN > 1 ? blockIdx.x / CLN : blockIdx.x;
    const int row0 = mt * TM;
    const int H = p.H, N4 = 4 * H, UB = H >> 3;
    for (int i = threadIdx.x; i < 4 * N4; i += 64 + 32 * NEPI_B) s_bsum[i] = 0.f;
    const uint32_t tmem_base = step_setup<CLM, CLN, NEPI_B>(smem, sb, warp);
    const uint32_t smem_base = smem_u32(smem);
    const uint16_t mc_mask = (uint16_t)((1u << CLM) - 1);
    const int nsteps = p.t1 - p.t0;
    const bool prologue = p.t1 == p.T;        // this launch starts at the last step: dG_{T-1} is produced here first

    if (warp == 0) {
        // ================= TMA producer: A = dG_t (tile-major: contiguous 16 KB boxes), B = rows of [W_hh^T ; W_ih^T] ======
        int stage = 0;
        uint32_t phase = 0, rphase = 0;
        for (int t = p.t1 - 1; t >= p.t0; --t) {
            if (t == p.t1 - 1 ? prologue : p.persist != 0) await_step<CLN>(sb, rphase);      // dG_t was written inside this launch
            for (int q = nrank; q < p.nchunks; q += CLN) {
                const int wslice = p.chunk_n[q] / CLM;
                const uint32_t bytes = (uint32_t)(A_BYTES + p.chunk_n[q] * KB * 4);
                const CUtensorMap* wm = p.chunk_map[q] ? &tmW2 : &tmW;
                for (int kb = 0; kb < p.nk; ++kb) {
                    mbar_wait(smem_u32(&sb->empty[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&sb->full[stage]);
                    const uint32_t sa = smem_base + (uint32_t)(stage * STAGE_BYTES);
                    if (elect_one_sync()) {
                        mbar_expect_tx(fb, bytes);
                        tma_load_5d(&tmG, fb, sa, 0, 0, kb, mt, t);
                        const uint32_t sw = sa + A_BYTES + (uint32_t)(mrank * wslice * KB * 4);
                        if (CLM > 1) tma_load_3d_multicast(wm, fb, sw, 0, p.chunk_row[q] + mrank * wslice, kb, mc_mask);
                        else tma_load_3d(wm, fb, sw, 0, p.chunk_row[q], kb);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        mma_role<CLM, CLN>(p, sb, smem_base, tmem_base, nrank, nsteps);
    } else {
        // ================= epilogue: dh columns -> cell backward of step t - 1 (dG_{t-1} in place); dx columns -> store ====
        constexpr int NPART = NEPI_B / 4;                 // warps per TMEM lane quarter: they split the chunk's columns
        const int qd = warp & 3, part = (warp - 2) >> 2;
        const int rt = qd * 32 + lane;
        const int row = row0 + rt;
        const bool rv = row < p.rows;
        const long long MT = (p.rows + TM - 1) / TM;
        const bool tile_ok = mt < MT;                     // CTAs without a row tile only run the protocol
        float* bs = s_bsum + qd * N4;
        int abuf = 0;
        uint32_t aphase = 0;
        auto ctx_for = [&](int tc) {                      // cell context of step tc for this thread's row
            CellCtx x;
            x.gates_t = p.gates + (((long long)tc * MT + mt) * UB * TM + rt) * 32;
            x.c_t = p.c + (((long long)tc * MT + mt) * UB * TM + rt) * 8;
            x.c_tm1 = tc > 0 ? p.c + (((long long)(tc - 1) * MT + mt) * UB * TM + rt) * 8 : nullptr;
            x.dh_out = p.dh_out + ((long long)tc * p.rows + row) * H;
            x.dc = p.dc + ((long long)mt * UB * TM + rt) * 8;
            x.rv = rv; x.first = tc == p.T - 1; x.round_tf32 = p.round_tf32;
            return x;
        };
        if (prologue) {
            // step T - 1 has no recurrent gradient: cell backward of the dh chunks this CTA owns, straight from memory
            const CellCtx x = ctx_for(p.T - 1);
            float zero[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) zero[i] = 0.f;
            for (int q = nrank; q < p.nchunks; q += CLN) {
                if (p.chunk_row[q] >= H) break;
                const int n_dh = min(p.chunk_n[q], H - p.chunk_row[q]) >> 3;
#pragma unroll 1
                for (int s = part; s < n_dh && tile_ok; s += NPART) {
                    const int ub = (p.chunk_row[q] >> 3) + s;
                    float v[32];
                    cell_bwd_block(x, ub, zero, v);
                    bias_accumulate(v, bs, ub, lane);
                }
            }
            publish_step<CLN>(sb, lane);
        }
        for (int t = p.t1 - 1; t >= p.t0; --t) {
            const bool cell = t > 0;
            const CellCtx x = ctx_for(cell ? t - 1 : 0);
            float* dxp = p.dx ? p.dx + ((long long)t * p.rows + row) * p.I : nullptr;
            for (int q = nrank; q < p.nchunks; q += CLN) {
                mbar_wait(smem_u32(&sb->tfull[abuf]), aphase);
                tc_fence_after();
                const uint32_t tacc = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(abuf * MAXN);
                const int nsb_all = p.chunk_n[q] >> 3;        // 8-column sub-blocks, dealt round-robin to the NPART warps
                const int nsb = (nsb_all - part + NPART - 1) / NPART;
                if (nsb <= 0) {                               // nothing to do in a narrow chunk: still release the buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&sb->tempty[abuf]));
                }
#pragma unroll 1
                for (int s = 0; s < nsb; ++s) {
                    const int cl0 = (part + s * NPART) * 8;   // column inside the chunk
                    const int j0 = p.chunk_row[q] + cl0;      // output column: [0, H) = dh units, [H, H + I) = dx
                    float d[8];
                    tmem_ld8(tacc + (uint32_t)cl0, d);
                    tmem_ld_wait();
                    if (s == nsb - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&sb->tempty[abuf]));
                    }
                    if (j0 >= H) {
                        if (rv && dxp) st_v8(dxp + (j0 - H), d);
                        continue;
                    }
                    if (!cell || !tile_ok) continue;          // warp-uniform
                    float v[32];
                    cell_bwd_block(x, j0 >> 3, d, v);
                    bias_accumulate(v, bs, j0 >> 3, lane);
                }
                if (++abuf == 2) { abuf = 0; aphase ^= 1; }
                // dG_{t-1} of this CTA is complete after its last chunk with dh columns: publish before the dx chunks, whose
                // MMAs then overlap the hand-over (every CTA owns at least one dh chunk, see the host side)
                if (p.persist && t > p.t0 && p.chunk_row[q] < H && (q + CLN >= p.nchunks || p.chunk_row[q + CLN] >= H))
                    publish_step<CLN>(sb, lane);
            }
        }
        // per-CTA column sums -> this CTA's slot (columns it never touched stay zero)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NEPI_B) : "memory");
        float* slot = p.bias_part + (long long)blockIdx.x * N4;
        for (int i = threadIdx.x - 64; i < N4; i += 32 * NEPI_B)
            slot[i] = (p.bias_accum ? slot[i] : 0.f) + ((s_bsum[i] + s_bsum[N4 + i]) + (s_bsum[2 * N4 + i] + s_bsum[3 * N4 + i]));
    }
    step_teardown<CL>(tmem_base, warp);
}

int smem_bytes(int H) { return NSTAGE * STAGE_BYTES + (16 * H * 4 > 8192 ? 16 * H * 4 : 8192) + (int)sizeof(Smem) + 64; }   // both directions

int make_seq_map(CUtensorMap* m, const float* base, int rows, int C, int steps) {
    // [steps][rows][C] as (c_inner 32, row, c_block, step); box = 128 rows of one 32-channel block
    cuuint64_t dims[4] = {32, (cuuint64_t)rows, (cuuint64_t)(C / 32), (cuuint64_t)steps};
    cuuint64_t str[3] = {(cuuint64_t)C * 4, 128, (cuuint64_t)rows * C * 4};
    cuuint32_t box[4] = {32, TM, 1, 1};
    return make_map(m, base, 4, dims, str, box);
}
int make_tiled_map(CUtensorMap* m, const float* base, int rows, int C, int steps) {
    // tile-major [steps][row tiles][C / 32][128][32]: one box = one contiguous 16 KB block
    const int mt = (rows + TM - 1) / TM, kbn = C / 32;
    cuuint64_t dims[5] = {32, TM, (cuuint64_t)kbn, (cuuint64_t)mt, (cuuint64_t)steps};
    cuuint64_t str[4] = {128, 16384, (cuuint64_t)kbn * 16384, (cuuint64_t)mt * kbn * 16384};
    cuuint32_t box[5] = {32, TM, 1, 1, 1};
    return make_map(m, base, 5, dims, str, box);
}
int make_w_map3(CUtensorMap* m, const float* base, int nrows, int K, int box_rows) {
    // [nrows][K] K-major as (k_inner 32, row, k_block)
    cuuint64_t dims[3] = {32, (cuuint64_t)nrows, (cuuint64_t)(K / 32)};
    cuuint64_t str[2] = {(cuuint64_t)K * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    return make_map(m, base, 3, dims, str, box);
}

typedef void (*StepKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const StepParams);

struct Shape {
    int clm, cln;
};
StepKernel kernel_for(bool fwd, Shape s) {
#define PICK(M, N) return fwd ? (StepKernel)lstm_fwd_tc_kernel<M, N> : (StepKernel)lstm_bwd_tc_kernel<M, N>
    if (s.cln == 8) PICK(1, 8);
    if (s.cln == 4) PICK(1, 4);
    if (s.cln == 2) PICK(1, 2);
    if (s.clm == 4) PICK(4, 1);
    if (s.clm == 2) PICK(2, 1);
    PICK(1, 1);
#undef PICK
}

cudaLaunchConfig_t make_cfg(int grid, int cl, int smem, cudaStream_t st, cudaLaunchAttribute* attr, int nthreads) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(nthreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}

// Cluster shape.  Many row tiles: weight multicast over CLM = 4 (else 2, else 1) row tiles - the largest whose clusters
// are all co-resident (a partial second wave would double the time of a persistent launch).  One row tile: split its
// `nsplit` chunks over CLN = 8 / 4 / 2 CTAs.
Shape pick_shape(bool fwd, int m_tiles, int nsplit, int smem) {
    static const int forced = getenv("SEFD_LSTM_CLUSTER") ? atoi(getenv("SEFD_LSTM_CLUSTER")) : 0;
    const bool split = m_tiles == 1 && nsplit > 1 && forced != 1;
    for (int cl = 8; cl >= 2; cl >>= 1) {
        if (split ? (nsplit % cl != 0) : (cl > 4)) continue;
        if (!split && forced && cl != forced) continue;
        Shape s = split ? Shape{1, cl} : Shape{cl, 1};
        StepKernel k = kernel_for(fwd, s);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaLaunchAttribute attr[1];
        const int clusters = split ? 1 : (m_tiles + cl - 1) / cl;
        cudaLaunchConfig_t cfg = make_cfg(clusters * cl, cl, smem, nullptr, attr, fwd ? 64 + 32 * NEPI_F : 64 + 32 * NEPI_B);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) == cudaSuccess && n >= clusters) return s;
        cudaGetLastError();
    }
    cudaFuncSetAttribute(kernel_for(fwd, Shape{1, 1}), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return Shape{1, 1};
}

int launch_step(bool fwd, StepKernel k, Shape s, int m_tiles, int smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b,
                const CUtensorMap& c, const CUtensorMap& d, const StepParams& p, int* grid_out) {
    const int cl = s.clm * s.cln;
    const int grid = s.cln > 1 ? m_tiles * s.cln : (m_tiles + cl - 1) / cl * cl;
    if (grid_out) *grid_out = grid;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = make_cfg(grid, cl, smem, st, attr, fwd ? 64 + 32 * NEPI_F : 64 + 32 * NEPI_B);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, a, b, c, d, p);
    if (e != cudaSuccess) {
        sefd_set_error("lstm_step_tc launch: %s", cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

bool persist_mode() {
    static const int persist = getenv("SEFD_LSTM_PERSIST") == nullptr || atoi(getenv("SEFD_LSTM_PERSIST")) != 0;
    return persist != 0;
}

}  // namespace

bool sefd_lstm_step_tc_eligible(int I, int H) { return I % 32 == 0 && H % 64 == 0 && H <= 512 && I >= 32 && I <= 1024; }
bool sefd_lstm_step_tc_has_backward() { return true; }

int sefd_lstm_step_bias_blocks(int rows) { return (rows + TM - 1) / TM + 8; }

int sefd_lstm_step_tc_forward(const SeqLstmFwdParams& f, cudaStream_t st) {
    const int H = f.w.H, I = f.w.I;
    SEFD_REQUIRE(sefd_lstm_step_tc_eligible(I, H), "lstm_step_tc: I=%d H=%d unsupported", I, H);
    SEFD_REQUIRE(f.w.Wcat_nk != nullptr && f.h_zero_slot, "lstm_step_tc: needs the concatenated weight operand and h with a zero slot in front");
    const int m_tiles = (f.rows + TM - 1) / TM, smem = smem_bytes(H);
    const Shape s = pick_shape(true, m_tiles, H / 64, smem);
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.rows = f.rows; p.T = f.T; p.I = I; p.H = H;
    p.kx = I / KB; p.nk = (I + H) / KB;
    p.nchunks = H / 64;
    for (int q = 0; q < p.nchunks; ++q) { p.chunk_n[q] = 256; p.chunk_row[q] = q * 256; }
    p.round_tf32 = f.round_h;
    p.gates = f.gates; p.hbuf = f.h - (long long)f.rows * H; p.c = f.c; p.bias = f.w.bias;
    CUtensorMap mx, mh, mw;
    SEFD_TRY(make_seq_map(&mx, f.x, f.rows, I, f.T));
    SEFD_TRY(make_seq_map(&mh, p.hbuf, f.rows, H, f.T + 1));
    SEFD_TRY(make_w_map3(&mw, f.w.Wcat_nk, 4 * H, I + H, 256 / s.clm));
    const double flops = 2.0 * f.rows * 4.0 * H * (I + H), bytes = 4.0 * f.rows * (I + 4.0 * H + 4.0 * H);
    StepKernel k = kernel_for(true, s);
    auto go = [&](int t0, int t1) -> int {
        p.t0 = t0; p.t1 = t1; p.persist = t1 - t0 > 1;
        sefd_prof_label("lstm_fwd_tc rows%d I%d H%d steps%d clm%d cln%d", f.rows, I, H, t1 - t0, s.clm, s.cln);
        SefdProfScope prof(SEFD_PROF_LSTM, flops * (t1 - t0), bytes * (t1 - t0), st);
        SEFD_TRY(launch_step(true, k, s, m_tiles, smem, st, mx, mh, mw, mw, p, nullptr));
        return sefd_check_launch("lstm_fwd_tc");
    };
    if (persist_mode()) return go(0, f.T);
    for (int t = 0; t < f.T; ++t) SEFD_TRY(go(t, t + 1));
    return 0;
}

int sefd_lstm_step_tc_backward(SeqLstmBwdParams& b, cudaStream_t st) {
    const int H = b.w.H, I = b.w.I, N4 = 4 * H;
    SEFD_REQUIRE(sefd_lstm_step_tc_eligible(I, H), "lstm_step_tc: I=%d H=%d unsupported", I, H);
    SEFD_REQUIRE(b.w.Wih_kn == b.w.Whh_kn + (long long)H * N4, "lstm_step_tc backward: W_ih^T must follow W_hh^T in memory");
    const int nb0 = 0;      // step T - 1 (no recurrent gradient) is the prologue of the launch that contains it
    b.dx_done = 0;
    const int m_tiles = (b.rows + TM - 1) / TM, smem = smem_bytes(H);
    StepParams p;
    memset(&p, 0, sizeof(p));
    // One row tile: its output columns are split over the cluster, every CTA gets dh chunks (H / 8 wide) and, when the dx
    // columns can be cut the same way, dx chunks (I / 8 wide) behind them.  Many row tiles: 256-wide chunks (a narrower
    // last one), weights multicast over the row tiles of a cluster.
    const bool can_split = H % 128 == 0;
    Shape s = pick_shape(false, m_tiles, can_split ? 8 : 1, smem);
    bool fuse_dx = b.dx != nullptr;
    int narrow = 0, wide = 256;
    if (s.cln > 1) {
        if (I % 128) fuse_dx = false;
        wide = H / 8;
        for (int q = 0; q < 8; ++q) { p.chunk_n[p.nchunks] = H / 8; p.chunk_row[p.nchunks] = q * (H / 8); p.chunk_map[p.nchunks++] = 0; }
        if (fuse_dx) {
            narrow = I / 8 != H / 8 ? I / 8 : 0;
            for (int q = 0; q < 8; ++q) { p.chunk_n[p.nchunks] = I / 8; p.chunk_row[p.nchunks] = H + q * (I / 8); p.chunk_map[p.nchunks++] = narrow ? 1 : 0; }
        }
    } else {
        const int Nout = H + (fuse_dx ? I : 0);
        for (int j = 0; j < Nout; j += 256) {
            const int n = Nout - j < 256 ? Nout - j : 256;
            if (n < 256) narrow = n;
            p.chunk_n[p.nchunks] = n; p.chunk_row[p.nchunks] = j; p.chunk_map[p.nchunks++] = n < 256 ? 1 : 0;
        }
        if (narrow && (narrow % (8 * s.clm) || narrow % 16)) {      // the narrow chunk cannot be sliced: no multicast
            SEFD_REQUIRE(narrow % 16 == 0, "lstm_step_tc backward: %d output columns unsupported", Nout);
            s = Shape{1, 1};
            cudaFuncSetAttribute(kernel_for(false, s), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        }
    }
    const int Nout = H + (fuse_dx ? I : 0);
    p.rows = b.rows; p.T = b.T; p.I = I; p.H = H;
    p.nk = N4 / KB;
    p.round_tf32 = b.round_tf32;
    p.gates = b.gates; p.c = const_cast<float*>(b.c); p.dh_out = b.dh_out; p.dc = b.dc; p.dx = fuse_dx ? b.dx : nullptr;
    p.bias_part = b.bias_part + (long long)nb0 * N4;
    CUtensorMap mg, mw, mw2;
    SEFD_TRY(make_tiled_map(&mg, b.gates, b.rows, N4, b.T));
    SEFD_TRY(make_w_map3(&mw, b.w.Whh_kn, Nout, N4, (wide < Nout ? wide : Nout) / s.clm));
    if (narrow) SEFD_TRY(make_w_map3(&mw2, b.w.Whh_kn, Nout, N4, narrow / s.clm));
    else mw2 = mw;
    const double flops = 2.0 * b.rows * (double)N4 * Nout, bytes = 4.0 * b.rows * (2.0 * N4 + 5.0 * H + (fuse_dx ? I : 0));
    StepKernel k = kernel_for(false, s);
    int grid = 0;
    auto go = [&](int t0, int t1) -> int {
        p.t0 = t0; p.t1 = t1; p.persist = t1 - t0 > 1;
        sefd_prof_label("lstm_bwd_tc rows%d I%d H%d Nout%d steps%d clm%d cln%d", b.rows, I, H, Nout, t1 - t0, s.clm, s.cln);
        SefdProfScope prof(SEFD_PROF_LSTM, flops * (t1 - t0), bytes * (t1 - t0), st);
        SEFD_TRY(launch_step(false, k, s, m_tiles, smem, st, mg, mg, mw, mw2, p, &grid));
        return sefd_check_launch("lstm_bwd_tc");
    };
    // iteration t consumes dG_t and produces dG_{t-1} (t >= 1) and dx_t; the bias sums of all iterations share one slot set
    if (persist_mode()) {
        SEFD_TRY(go(0, b.T));
    } else {
        for (int t = b.T - 1; t >= 0; --t) {
            p.bias_accum = t != b.T - 1;
            SEFD_TRY(go(t, t + 1));
        }
    }
    b.bias_blocks = nb0 + grid;
    b.dx_done = fuse_dx ? 1 : 0;
    return 0;
}
