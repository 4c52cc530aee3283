// Time-major nn.LSTM layer engine for FullSubNet's SequenceModel (tools_for_model.py:726-795) and cfg.lstm = 'real'
// (models.py:96-105): any hidden size that is a multiple of 64, any number of independent sequences ("rows").
//
// Layout: every per-step tensor is [T][rows][C] (time-major), so that step t of ALL sequences is one dense
// [rows x C] matrix: the recurrence h_{t-1} W_hh^T is a plain GEMM per step and the tile of a GEMM CTA is 128
// consecutive sequences.  The 4H gate columns are stored INTERLEAVED in blocks of 8 hidden units,
//     n' = (u / 8) * 32 + gate * 8 + (u % 8)        (reference order n = gate * H + u, gates i, f, g, o)
// so that every 32-column block (= one 128-byte TMA / MMA k-block) holds i, f, g, o of the same 8 units: a 256-column
// accumulator tile of the recurrent GEMM covers 64 units with all their gates and the LSTM cell can run in the GEMM
// epilogue (lstm_step_tc.cu); packed weights / bias use the same column order.
//
// The fused kernels keep the per-step state they alone touch TILE-MAJOR (`tiled` flag below): rows are cut into tiles of
// 128 and each 32-column block of a tile is one contiguous 16 KB box,
//     gates / dG : [T][row tiles][4H / 32][128][32]        c : [T][row tiles][H / 8][128][8]        dc : [row tiles][H / 8][128][8]
// so that a thread-per-row epilogue reads / writes whole 128-byte lines that are contiguous across the warp (the row-major
// form made every access a 32-byte sector at a 6 KB stride: DRAM ran at a third of its rate) and the backward's A operand
// dG_t is fetched by TMA as contiguous boxes.  Buffers must be sized for rows rounded up to a multiple of 128.
#pragma once
#include "common.cuh"

struct SeqLstmWeights {
    const float* Wih_nk;   // [4H'][I]   (n' rows, k contiguous)  forward operand of the tensor-core engine / dgrad "W"
    const float* Wih_kn;   // [I][4H']                            forward operand of the fp32 engine / dgrad "Wnk"
    const float* Whh_nk;   // [4H'][H]
    const float* Whh_kn;   // [H][4H']
    const float* bias;     // [4H'] = b_ih + b_hh
    const float* Wcat_nk;  // [4H'][I + H] = [W_ih | W_hh] rows (operand of the fused forward step kernel) or null
    // the fused backward step kernel reads [W_hh^T ; W_ih^T] = [(H + I)][4H']: Wih_kn must FOLLOW Whh_kn directly in memory
    int I, H;              // I = input width as stored (padded to a multiple of 32)
};

struct SeqLstmPackParams {
    const float *w_ih, *w_hh, *b_ih, *b_hh;   // reference layouts: [4H][I_real], [4H][H], [4H], [4H]
    int I_real, I, H;
    float *Wih_nk, *Wih_kn, *Whh_nk, *Whh_kn, *bias;
    float* Wcat_nk;        // optional
    int round_tf32;
    int kd;                // input-column permutation of layer 0 (0 / 1: none): the stored column k' = d * (I / kd) + c holds the
                           // reference's column c * kd + d (DCCRN's [T, B, C * D] LSTM input read as D blocks of C channels)
};
int sefd_seqlstm_pack(const SeqLstmPackParams& p, cudaStream_t st);

// dW[n][k] (reference layout [4H][K_real]) = sum_s part[s][k][n'(n)]   (partials [nsplit][K][4H'] of sefd_wgrad)
int sefd_seqlstm_fold_wgrad(const float* part, int nsplit, long long split_stride, int K, int K_real, int H, float* dW,
                            cudaStream_t st, int kd = 1);
// db_ih[n] = db_hh[n] = sum_blk part[blk][n'(n)]
int sefd_seqlstm_fold_bias(const float* part, int nblk, int H, float* db_ih, float* db_hh, cudaStream_t st);

struct SeqLstmFwdParams {
    const float* x;        // [T][rows][I] layer input (tf32-rounded when the tensor-core engine is on)
    SeqLstmWeights w;
    float* gates;          // [T][rows][4H'] out: activated gates (kept for the backward)
    float* h;              // [T][rows][H]
    float* c;              // [T][rows][H]
    int rows, T;
    int round_h;           // h feeds tensor-core GEMMs: round to tf32 while writing
    int h_zero_slot;       // h is preceded by one all-zero step ([-1] = h_{-1} = 0): the fused kernel reads h_{t-1} uniformly
    int tiled;             // out: gates / c were written tile-major (the fused path ran); the backward must be told
};
int sefd_seqlstm_forward(const SeqLstmFwdParams& p, cudaStream_t st);

struct SeqLstmBwdParams {
    SeqLstmWeights w;
    float* gates;          // in: activated gates; out: gradient w.r.t. the gate pre-activations (in place)
    const float* c;        // [T][rows][H]
    const float* dh_out;   // [T][rows][H] gradient arriving at every h_t from above
    float* dh_rec;         // [rows][H] scratch: recurrent gradient
    float* dc;             // [rows][H] scratch: cell-state gradient carried backwards
    float* bias_part;      // [sefd_seqlstm_bias_blocks(rows)][4H'] per-block column sums of dG over all steps
    int bias_blocks;       // out: how many of those slots the path that ran has written (sefd_seqlstm_fold_bias sums them)
    float* dx;             // optional [T][rows][I]: gradient w.r.t. the layer input; when the fused kernel runs it is
    int dx_done;           // written by the step epilogue and dx_done is set (else the caller computes it from dG)
    int rows, T;
    int round_tf32;        // dG feeds tensor-core GEMMs
    int tiled;             // in: layout the forward reported (tile-major: the fused backward runs, dG stays tile-major)
};
int sefd_seqlstm_bias_blocks(int rows);
// the generic cell backward of ONE step (t = T - 1: no recurrent gradient, initialises dc and the bias slots); *nblk = slots used
int sefd_seqlstm_cell_bwd_step(SeqLstmBwdParams& p, int t, int* nblk, cudaStream_t st);
int sefd_seqlstm_backward(const SeqLstmBwdParams& p, cudaStream_t st);

// nn.LSTM(dropout = p) between stacked layers (tools_for_model.py:746): dst = src * m, m = 0 or 1 / (1 - p).
// mask != null: injected multiplier tensor (tests); else Philox4x32-10 keyed by (seed, stream) and the element index.
int sefd_dropout_apply(const float* src, float* dst, long long n, float p, const float* mask, unsigned long long seed,
                       unsigned int stream_id, int round_tf32, cudaStream_t st);

// 1: fused tcgen05 step kernels (GEMM + LSTM cell in the epilogue), 0: generic tap-GEMM per step + cell kernels
int sefd_seqlstm_fused_enabled();
