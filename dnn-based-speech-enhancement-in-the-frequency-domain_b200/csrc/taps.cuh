// Tap lists and tensor-view helpers shared by the model orchestration and the op-level C entry points.
#pragma once
#include <string.h>
#include "common.cuh"

static inline TapSrc src4(const float* p, int F, int Ts, int Ctot, int C) {
    TapSrc s;
    s.p = p;
    s.sT = Ctot;
    s.sF = (long long)Ts * Ctot;
    s.sB = (long long)F * Ts * Ctot;
    s.C = C;
    return s;
}
static inline TapDst dst4(float* p, int F, int Ts, int Ctot, int N) {
    TapDst d;
    d.p = p;
    d.sT = Ctot;
    d.sF = (long long)Ts * Ctot;
    d.sB = (long long)F * Ts * Ctot;
    d.N = N;
    return d;
}
static inline TapSrc no_src() {
    TapSrc s;
    memset(&s, 0, sizeof(s));
    return s;
}
static inline TapDst no_dst() {
    TapDst d;
    memset(&d, 0, sizeof(d));
    return d;
}


// kernel (5,2), stride (2,1): the 10 (kf, kt) taps.  "down" = stride-2 gather over F (conv forward, convT
// data gradient); "up" = the two output-row phases of the stride-2 scatter (convT forward, conv data gradient).
static inline void conv_taps_down(TapGemmParams& g, int dt_sign /* -1: dt = kt-1 (conv fwd); +1: dt = +kt (convT dgrad) */) {
    g.ntaps = 10;
    for (int kf = 0; kf < 5; ++kf)
        for (int kt = 0; kt < 2; ++kt) {
            const int i = kf * 2 + kt;
            g.df[i] = kf - 2;
            g.dt[i] = dt_sign < 0 ? kt - 1 : kt;
            g.wslab[i] = i;
        }
    g.fi_mul = 2;
    g.fo_mul = 1;
    g.fo_off = 0;
}
static inline void conv_taps_up(TapGemmParams& g, int phase, int mode /* 0: convT fwd (dt = -kt); 1: conv dgrad (dt = 1-kt) */) {
    g.ntaps = 0;
    for (int kf = phase; kf < 5; kf += 2)
        for (int kt = 0; kt < 2; ++kt) {
            const int i = g.ntaps++;
            g.df[i] = (2 + phase - kf) / 2;      // even: kf 0,2,4 -> +1,0,-1 ; odd: kf 1,3 -> +1,0
            g.dt[i] = mode == 0 ? -kt : 1 - kt;
            g.wslab[i] = kf * 2 + kt;
        }
    g.fi_mul = 1;
    g.fo_mul = 2;
    g.fo_off = phase;
}

