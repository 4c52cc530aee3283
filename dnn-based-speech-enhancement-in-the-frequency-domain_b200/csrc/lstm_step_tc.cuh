// Fused tcgen05 LSTM step kernels (lstm_step_tc.cu): one launch per time step computes the gate pre-activations of ALL
// sequences as a TF32 GEMM [rows x (I + H)] x [(I + H) x 4H] (x_t and h_{t-1} are the two K sources, no pre-activation
// tensor goes through HBM) and runs the LSTM cell in the epilogue; the backward step is the transposed GEMM
// dG_t [rows x 4H] x [4H x (I + H)] whose epilogue stores dx_t and turns dh_{t-1} into dG_{t-1}.
#pragma once
#include "lstm_seq.cuh"

bool sefd_lstm_step_tc_eligible(int I, int H);
int sefd_lstm_step_bias_blocks(int rows);
bool sefd_lstm_step_tc_has_backward();
int sefd_lstm_step_tc_forward(const SeqLstmFwdParams& p, cudaStream_t st);
int sefd_lstm_step_tc_backward(SeqLstmBwdParams& p, cudaStream_t st);
