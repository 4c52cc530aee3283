#include "lstm_step_tc.cuh"

bool sefd_lstm_step_tc_eligible(int, int) { return false; }
int sefd_lstm_step_bias_blocks(int) { return 0; }
int sefd_lstm_step_tc_forward(const SeqLstmFwdParams&, cudaStream_t) {
    sefd_set_error("lstm_step_tc: not built");
    return -1;
}
int sefd_lstm_step_tc_backward(const SeqLstmBwdParams&, cudaStream_t) {
    sefd_set_error("lstm_step_tc: not built");
    return -1;
}
