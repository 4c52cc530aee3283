// Optional per-kernel-category device timing with CUDA events on the launching stream (used by bench.py for
// the roofline figures; off by default, costs nothing when off).
#pragma once
#include <cuda_runtime.h>

enum SefdProfCat {
    SEFD_PROF_TAPGEMM = 0,   // conv / convT / linear forward and data gradients
    SEFD_PROF_WGRAD = 1,     // weight gradients
    SEFD_PROF_BN = 2,        // BatchNorm + PReLU forward / backward passes
    SEFD_PROF_LSTM = 3,      // recurrent kernels
    SEFD_PROF_STFT = 4,      // STFT, mask + ISTFT and its adjoint, loss passes
    SEFD_PROF_MISC = 5,      // weight packing / folding, small reductions, Adam
    SEFD_PROF_SKINNY = 6,    // CUDA-core kernels of the 2-channel ends (encoder 0, decoder 5): HBM-bound
    SEFD_PROF_NCAT = 7
};

bool sefd_prof_on();
// cudaGetLastError() is per host thread and keeps a NON-sticky error (a failed launch configuration, a rejected attribute, a
// query that returned "not ready") until somebody reads it: left pending by other code in the process it would be blamed on
// the next kernel this library launches.  Every launch scope reads it first; a pending error is counted and kept
// (sefd_stale_cuda_errors / sefd_last_stale_cuda_error) instead of failing an unrelated call.  Sticky errors are unaffected.
void sefd_absorb_stale_error();
void sefd_prof_push(int cat, double flops, double bytes, cudaStream_t st, bool begin);
void sefd_prof_label(const char* fmt, ...);   // label for the NEXT pushed record (ignored when profiling is off)

struct SefdProfScope {
    int cat;
    cudaStream_t st;
    bool on;
    SefdProfScope(int c, double flops, double bytes, cudaStream_t s) : cat(c), st(s), on(sefd_prof_on()) {
        sefd_absorb_stale_error();
        if (on) sefd_prof_push(cat, flops, bytes, st, true);
    }
    ~SefdProfScope() {
        if (on) sefd_prof_push(cat, 0, 0, st, false);
    }
};
