// Cluster-split recurrence of the complex LSTM (tools_for_model.py:141-181; nn.LSTM, H = 128, gate order i, f, g, o):
// the north-star form of lstm.cu's recurrence.  A thread-block CLUSTER of 4 CTAs owns a group of R sequence rows of one
// LSTM for all T steps; CTA j owns the hidden units [32 j, 32 j + 32), i.e. 128 of the 512 gate rows of W_hh:
//   * its 64 KB weight slice is staged ONCE with bulk async copies (cp.async.bulk, TMA engine) into shared memory and from
//     there into REGISTERS (one gate row = 128 weights per thread): the step loop reads no weights from memory at all
//     (lstm.cu keeps 40 of every 128 weights in shared memory: its step is shared-memory-wavefront bound);
//   * every step each CTA computes its 128 gate pre-activations for the R rows (thread = gate row x row pair, h_{t-1}
//     broadcast from shared memory), applies the cell for its 32 units and writes its slice of h_t into the h buffer of
//     ALL FOUR CTAs through distributed shared memory with asynchronous stores that complete on the RECEIVER's mbarrier
//     (st.async ... mbarrier::complete_tx::bytes), double-buffered: no cluster barrier and no release fence in the step loop
//     (a barrier.cluster per step measured 0.44 us of a 1.46 us step: its release waits for the step's global stores);
//   * the weight matrix is used R times per step instead of once.
// Backward mirrors it: cell backward for the CTA's units, dh_{t-1} partial sums over the CTA's 128 gate rows (W_hh columns
// in registers), cross-CTA reduction through distributed shared memory.
// Gate pre-activations of all steps come from the tap-GEMM (input projections), exactly like lstm.cu; same buffers, same
// layouts, same arithmetic order inside a dot product up to the split over CTAs.
#include <stdlib.h>

#include "lstm.cuh"
#include "prof.cuh"

namespace {

constexpr int H = 128, G4 = 512, CL = 4, UPC = H / CL;   // units per CTA
constexpr int NT = 256;
constexpr int WP = H + 4;                               // staging row pitch (floats): rows 4 banks apart -> 4-way instead of 32-way conflicts

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The release of barrier.cluster.arrive waits for EVERY earlier write of the CTA to be performed, global stores included
// (measured: 0.44 us of a 1.46 us step).  The step loops therefore arrive right after their distributed-shared-memory
// stores and issue the step's global stores between arrive and wait: they get a whole step to drain before the next release.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster(uint32_t local_addr, uint32_t cta, float v) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(r), "f"(v) : "memory");
}
// asynchronous 4-byte store into the shared memory of CTA `cta` that completes (by 4 bytes) on the mbarrier at `local_bar`'s
// offset in THAT CTA: data and its arrival notice travel together, the receiver only waits on its own mbarrier - no cluster
// barrier and no release fence on the step's critical path
__device__ __forceinline__ void st_async_cluster(uint32_t local_addr, uint32_t local_bar, uint32_t cta, float v) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(local_bar), "r"(cta));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb)
                 : "memory");
}
// 16-byte variant: four consecutive floats per store (a quarter of the mbarrier transactions)
__device__ __forceinline__ void st_async_cluster4(uint32_t local_addr, uint32_t local_bar, uint32_t cta, float a, float b, float c, float d) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(local_bar), "r"(cta));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(ra),
                 "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(rb)
                 : "memory");
}
// bulk async copy global -> shared (TMA engine), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect_(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned long long spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (spin > (1ull << 24)) __trap();
    }
}

// stage this CTA's weight slice (4 gates x 32 units x 128) into shared memory: rows gate * 128 + 32 j + u of W_hh
__device__ __forceinline__ void stage_weights(const float* W, int j, float* Wst, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init_(smem_addr(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_(smem_addr(bar), 4 * UPC * H * 4);
        for (int g = 0; g < 4; ++g)
            for (int u = 0; u < UPC; ++u)       // one 512-byte row per copy (padded pitch in shared memory)
                bulk_g2s(smem_addr(Wst + (g * UPC + u) * WP), W + (size_t)(g * H + UPC * j + u) * H, H * 4, smem_addr(bar));
    }
    __syncthreads();
    mbar_wait_(smem_addr(bar), 0);
}

// ---------------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) lstm_cluster_fwd_kernel(const LstmFwdParams p) {
    extern __shared__ __align__(128) float sm[];
    float* Wst = sm;                       // [4][32][128] staging (64 KB), dead after the register load
    float* hs = Wst + 4 * UPC * WP;        // [2][R][128] h_{t-1} (double buffered; written by all four CTAs)
    float* gs = hs + 2 * R * H;            // [R][4][32] activated gates of this CTA's units
    uint64_t* bar = reinterpret_cast<uint64_t*>(gs + R * 4 * UPC);
    uint64_t* hb = bar + 1;                // [2]: h buffer b is complete (R * 128 floats have landed from the four CTAs)
    const int tid = threadIdx.x, j = (int)cta_rank();
    const int groups_per_lstm = p.rows / R;
    const int grp = blockIdx.x / CL, lstm = grp / groups_per_lstm, row0 = (grp % groups_per_lstm) * R;
    if (tid == 0) {
        mbar_init_(smem_addr(&hb[0]), 1);
        mbar_init_(smem_addr(&hb[1]), 1);
    }
    const int q = tid & 127, gate = q >> 5, u = q & 31, rh = tid >> 7;          // gate row (gate, u); rows [rh * R/2, (rh + 1) * R/2)
    constexpr int RP = R / 2;
    stage_weights(p.Whh + (size_t)lstm * G4 * H, j, Wst, bar);
    float w[H];
#pragma unroll
    for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 v = *reinterpret_cast<const float4*>(Wst + (gate * UPC + u) * WP + 4 * k4);
        w[4 * k4] = v.x; w[4 * k4 + 1] = v.y; w[4 * k4 + 2] = v.z; w[4 * k4 + 3] = v.w;
    }
    for (int i = tid; i < 2 * R * H; i += NT) hs[i] = 0.f;
    const size_t rbase = (size_t)lstm * p.rows + row0;
    float* Gp[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) Gp[r] = p.G + (rbase + rh * RP + r) * p.T * G4 + gate * H + UPC * j + u;
    // cell role (threads < R * 32): row cr, unit cu of this CTA
    const int cr = tid >> 5, cu = tid & 31;
    const bool cell = tid < R * UPC;
    const size_t hbase = (rbase + (cell ? cr : 0)) * p.T * H + UPC * j + cu;
    float c = 0.f;
    float pre[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) pre[r] = Gp[r][0];
    cluster_sync();                        // every CTA's h buffers are zeroed and its barriers initialised before a peer writes
    for (int t = 0; t < p.T; ++t) {
        if (tid == 0 && t + 1 < p.T) mbar_expect_(smem_addr(&hb[(t + 1) & 1]), R * H * 4);     // arm the buffer h_t lands in
        const float* hc = hs + (t & 1) * R * H + rh * RP * H;
        float acc[RP], acc2[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) { acc[r] = pre[r]; acc2[r] = 0.f; }
        if (t + 1 < p.T) {
#pragma unroll
            for (int r = 0; r < RP; ++r) pre[r] = Gp[r][(size_t)(t + 1) * G4];
        }
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                const float4 h4 = *reinterpret_cast<const float4*>(hc + r * H + 4 * k4);
                acc[r] = fmaf(w[4 * k4 + 0], h4.x, acc[r]);
                acc2[r] = fmaf(w[4 * k4 + 1], h4.y, acc2[r]);
                acc[r] = fmaf(w[4 * k4 + 2], h4.z, acc[r]);
                acc2[r] = fmaf(w[4 * k4 + 3], h4.w, acc2[r]);
            }
        }
        float act[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) {
            const float pa = acc[r] + acc2[r];
            act[r] = gate == 2 ? tanhf(pa) : sigmoidf_(pa);
            gs[((rh * RP + r) * 4 + gate) * UPC + u] = act[r];
        }
        __syncthreads();
        float h = 0.f;
        if (cell) {
            const float ig = gs[(cr * 4 + 0) * UPC + cu], fg = gs[(cr * 4 + 1) * UPC + cu];
            const float gg = gs[(cr * 4 + 2) * UPC + cu], og = gs[(cr * 4 + 3) * UPC + cu];
            c = fmaf(fg, c, ig * gg);
            h = og * tanhf(c);
            // four neighbouring units travel in one 16-byte store: lane 4m collects the values of lanes 4m .. 4m + 3
            const float h1 = __shfl_down_sync(0xffffffffu, h, 1), h2 = __shfl_down_sync(0xffffffffu, h, 2), h3 = __shfl_down_sync(0xffffffffu, h, 3);
            if (t + 1 < p.T && (cu & 3) == 0) {
                const uint32_t dst = smem_addr(hs + ((t + 1) & 1) * R * H + cr * H + UPC * j + cu);
                const uint32_t nb = smem_addr(&hb[(t + 1) & 1]);
#pragma unroll
                for (int peer = 0; peer < CL; ++peer) st_async_cluster4(dst, nb, (uint32_t)peer, h, h1, h2, h3);
            }
        }
#pragma unroll
        for (int r = 0; r < RP; ++r) Gp[r][(size_t)t * G4] = act[r];
        if (cell) {
            p.Hh[hbase + (size_t)t * H] = h;
            p.Cc[hbase + (size_t)t * H] = c;
        }
        // h_t is complete in this CTA once all R * 128 values have landed; the senders read gs before sending, so gs may be
        // overwritten by whoever passes this wait
        // (buffer 1 is first filled for step 1, buffer 0 for step 2: fill number n = t / 2 either way)
        if (t + 1 < p.T) mbar_wait_(smem_addr(&hb[(t + 1) & 1]), (uint32_t)((t >> 1) & 1));
    }
    cluster_sync();                        // nobody leaves while a peer could still be addressing it
}

// ---------------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) lstm_cluster_bwd_kernel(const LstmBwdParams p) {
    extern __shared__ __align__(128) float sm[];
    float* Wst = sm;                       // [4][32][128] staging
    float* dgs = Wst + 4 * UPC * WP;       // [R][128]: gate gradients of this CTA's 128 gate rows (index gate * 32 + u)
    float* red = dgs + R * G4 / CL;        // [2][CL][R][32]: partial dh of this CTA's units from every CTA (double buffered)
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * CL * R * UPC);
    uint64_t* rb = bar + 1;                // [2]: reduction buffer b holds the four CTAs' partial sums
    const int tid = threadIdx.x, j = (int)cta_rank();
    const int groups_per_lstm = p.rows / R;
    const int grp = blockIdx.x / CL, lstm = grp / groups_per_lstm, row0 = (grp % groups_per_lstm) * R;
    if (tid == 0) {
        mbar_init_(smem_addr(&rb[0]), 1);
        mbar_init_(smem_addr(&rb[1]), 1);
    }
    constexpr int RP = R / 2;
    const int k = tid & 127, rh = tid >> 7;                 // mat-vec role: hidden unit k (all 128), rows [rh * RP, ...)
    stage_weights(p.Whh + (size_t)lstm * G4 * H, j, Wst, bar);
    float w[H];                                             // column k of this CTA's 128 gate rows: w[n] = W[row n][k]
#pragma unroll
    for (int n = 0; n < H; ++n) w[n] = Wst[n * WP + k];     // lanes read consecutive k: conflict-free
    for (int i = tid; i < 2 * CL * R * UPC; i += NT) red[i] = 0.f;
    const size_t rbase = (size_t)lstm * p.rows + row0;
    // cell role (threads < R * 32): row cr, unit cu of this CTA
    const int cr = tid >> 5, cu = tid & 31;
    const bool cell = tid < R * UPC;
    const size_t row = rbase + (cell ? cr : 0);
    const float* Gp = p.G + row * p.T * G4 + UPC * j + cu;
    const float* Cp = p.Cc + row * p.T * H + UPC * j + cu;
    const float* dHp = p.dH + row * p.T * H + UPC * j + cu;
    float* dGp = p.dG + row * p.T * G4 + UPC * j + cu;
    float dc_next = 0.f;
    float n_i = 0.f, n_f = 0.f, n_g = 0.f, n_o = 0.f, n_c = 0.f, n_cp = 0.f, n_dh = 0.f;     // operands of the next step, prefetched
    if (cell) {
        const int t = p.T - 1;
        n_i = Gp[(size_t)t * G4]; n_f = Gp[(size_t)t * G4 + H]; n_g = Gp[(size_t)t * G4 + 2 * H]; n_o = Gp[(size_t)t * G4 + 3 * H];
        n_c = Cp[(size_t)t * H]; n_cp = t > 0 ? Cp[(size_t)(t - 1) * H] : 0.f; n_dh = dHp[(size_t)t * H];
    }
    cluster_sync();
    int it = 0;                            // iteration counter: buffer (it & 1) receives this iteration's partial sums
    for (int t = p.T - 1; t >= 0; --t, ++it) {
        float o_i = 0.f, o_f = 0.f, o_g = 0.f, o_o = 0.f;
        if (cell) {
            const float ig = n_i, fg = n_f, gg = n_g, og = n_o, ct = n_c, cprev = n_cp, dho = n_dh;
            if (t > 0) {
                const int s = t - 1;
                n_i = Gp[(size_t)s * G4]; n_f = Gp[(size_t)s * G4 + H]; n_g = Gp[(size_t)s * G4 + 2 * H]; n_o = Gp[(size_t)s * G4 + 3 * H];
                n_c = Cp[(size_t)s * H]; n_cp = s > 0 ? Cp[(size_t)(s - 1) * H] : 0.f; n_dh = dHp[(size_t)s * H];
            }
            // recurrent gradient: the four CTAs' partial sums for this unit (written by the previous iteration)
            const float* rp = red + ((it + 1) & 1) * CL * R * UPC + cr * UPC + cu;     // previous iteration's buffer (zeros at first)
            const float dh = dho + ((rp[0] + rp[R * UPC]) + (rp[2 * R * UPC] + rp[3 * R * UPC]));
            const float tc = tanhf(ct);
            const float dc = fmaf(dh * og, 1.f - tc * tc, dc_next);
            float dpi = dc * gg * ig * (1.f - ig);
            float dpf = dc * cprev * fg * (1.f - fg);
            float dpg = dc * ig * (1.f - gg * gg);
            float dpo = dh * tc * og * (1.f - og);
            dc_next = dc * fg;
            dgs[cr * H + 0 * UPC + cu] = dpi; dgs[cr * H + 1 * UPC + cu] = dpf;
            dgs[cr * H + 2 * UPC + cu] = dpg; dgs[cr * H + 3 * UPC + cu] = dpo;
            if (p.round_tf32) { dpi = tf32_rn(dpi); dpf = tf32_rn(dpf); dpg = tf32_rn(dpg); dpo = tf32_rn(dpo); }
            o_i = dpi; o_f = dpf; o_g = dpg; o_o = dpo;
        }
        __syncthreads();
        if (t > 0) {
            float acc[RP], acc2[RP];
#pragma unroll
            for (int r = 0; r < RP; ++r) acc[r] = acc2[r] = 0.f;
            const float* dg = dgs + rh * RP * H;
#pragma unroll
            for (int n4 = 0; n4 < H / 4; ++n4) {
#pragma unroll
                for (int r = 0; r < RP; ++r) {
                    const float4 d4 = *reinterpret_cast<const float4*>(dg + r * H + 4 * n4);
                    acc[r] = fmaf(w[4 * n4 + 0], d4.x, acc[r]);
                    acc2[r] = fmaf(w[4 * n4 + 1], d4.y, acc2[r]);
                    acc[r] = fmaf(w[4 * n4 + 2], d4.z, acc[r]);
                    acc2[r] = fmaf(w[4 * n4 + 3], d4.w, acc2[r]);
                }
            }
            // partial dh_{t-1}[row][k] of this CTA's gate rows -> slot j of the CTA that owns unit k
            const int owner = k >> 5, ku = k & 31;
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                const uint32_t dst = smem_addr(red + (it & 1) * CL * R * UPC + j * R * UPC + (rh * RP + r) * UPC + ku);
                st_cluster(dst, (uint32_t)owner, acc[r] + acc2[r]);
            }
        }
        // (measured: here the plain distributed-shared-memory stores + split cluster barrier beat the st.async / mbarrier
        // hand-over of the forward kernel, 1.30 vs 1.6 us per step)
        cluster_arrive();                  // releases the partial sums; the gate-gradient stores below drain during the next step
        if (cell) {
            dGp[(size_t)t * G4] = o_i; dGp[(size_t)t * G4 + H] = o_f; dGp[(size_t)t * G4 + 2 * H] = o_g; dGp[(size_t)t * G4 + 3 * H] = o_o;
        }
        cluster_wait();
    }
    cluster_sync();
}

template <int R>
int launch_fwd(const LstmFwdParams& p, cudaStream_t st) {
    const size_t smem = sizeof(float) * (4 * UPC * WP + 2 * R * H + R * 4 * UPC) + 32;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(lstm_cluster_fwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int nl = p.nl ? p.nl : 2;
    sefd_absorb_stale_error();
    lstm_cluster_fwd_kernel<R><<<nl * (p.rows / R) * CL, NT, smem, st>>>(p);
    return sefd_check_launch("lstm_cluster_fwd");
}
template <int R>
int launch_bwd(const LstmBwdParams& p, cudaStream_t st) {
    const size_t smem = sizeof(float) * (4 * UPC * WP + R * G4 / CL + 2 * CL * R * UPC) + 32;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(lstm_cluster_bwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int nl = p.nl ? p.nl : 2;
    sefd_absorb_stale_error();
    lstm_cluster_bwd_kernel<R><<<nl * (p.rows / R) * CL, NT, smem, st>>>(p);
    return sefd_check_launch("lstm_cluster_bwd");
}

bool enabled() {
    static const int on = getenv("SEFD_LSTM_CLUSTERED") == nullptr || atoi(getenv("SEFD_LSTM_CLUSTERED")) != 0;
    return on != 0;
}

}  // namespace

// rows per cluster: 4 when the clusters of all groups fit on the device at once, else 8 (fewer, fatter clusters)
int sefd_lstm_cluster_rows(int rows, int nl) {
    if (!enabled()) return 0;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (rows % 4 == 0 && nl * (rows / 4) * CL <= sms - 16) return 4;     // cluster placement strands a few SMs: keep a margin
    if (rows % 8 == 0 && nl * (rows / 8) * CL <= 2 * sms) return 8;
    return 0;
}

int sefd_lstm_cluster_fwd(const LstmFwdParams& p, int R, cudaStream_t st) { return R == 4 ? launch_fwd<4>(p, st) : launch_fwd<8>(p, st); }
int sefd_lstm_cluster_bwd(const LstmBwdParams& p, int R, cudaStream_t st) { return R == 4 ? launch_bwd<4>(p, st) : launch_bwd<8>(p, st); }
