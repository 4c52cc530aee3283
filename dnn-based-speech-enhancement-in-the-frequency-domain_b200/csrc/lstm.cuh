#pragma once
#include "common.cuh"

// Row = one sequence. rows = 2B per LSTM (real-part inputs then imag-part inputs); two LSTMs (real_lstm,
// imag_lstm) are laid out back to back in every buffer.
struct LstmFwdParams {
    const float* Whh;   // [2][512][128]  weight_hh_l0 of real_lstm, imag_lstm
    float* G;           // [2][rows][T][512]  in: x W_ih^T + b_ih + b_hh ; out: activated gates i,f,g,o
    float* Hh;          // [2][rows][T][128]
    float* Cc;          // [2][rows][T][128]
    int rows, T;
    int nl;             // LSTMs laid out back to back (0 = 2: the complex pair; 1: CRN's single nn.LSTM)
};
struct LstmBwdParams {
    const float* Whh;
    const float* G;     // activated gates from the forward
    const float* Cc;
    const float* dH;    // [2][rows][T][128]  gradient arriving at every h_t from above
    float* dG;          // [2][rows][T][512]  gradient w.r.t. the gate pre-activations
    int rows, T;
    int round_tf32;     // dG feeds tensor-core GEMMs
    int nl;
};
int sefd_lstm_fwd_launch(const LstmFwdParams& p, cudaStream_t st);
int sefd_lstm_bwd_launch(const LstmBwdParams& p, cudaStream_t st);

// cluster-split variant (lstm_cluster.cu): rows per cluster (4 or 8) when it applies to (rows, nl), else 0
int sefd_lstm_cluster_rows(int rows, int nl);
int sefd_lstm_cluster_fwd(const LstmFwdParams& p, int R, cudaStream_t st);
int sefd_lstm_cluster_bwd(const LstmBwdParams& p, int R, cudaStream_t st);
