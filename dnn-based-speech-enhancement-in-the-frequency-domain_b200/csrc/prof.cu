#include <vector>

#include "../../include/sefd.h"
#include "common.cuh"
#include "prof.cuh"

namespace {
struct Rec {
    int cat;
    cudaEvent_t e0, e1;
    double flops, bytes;
};
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

bool sefd_prof_on() { return g_on; }

void sefd_prof_push(int cat, double flops, double bytes, cudaStream_t st, bool begin) {
    if (begin) {
        Rec r;
        r.cat = cat;
        r.flops = flops;
        r.bytes = bytes;
        r.e0 = get_event();
        r.e1 = get_event();
        cudaEventRecord(r.e0, st);
        g_recs.push_back(r);
    } else {
        // scopes nest strictly (LIFO) on the host, so the innermost open record of this category is the last one
        for (int i = (int)g_recs.size() - 1; i >= 0; --i)
            if (g_recs[i].cat == cat) {
                cudaEventRecord(g_recs[i].e1, st);
                break;
            }
    }
}

extern "C" {

int sefd_prof_enable(int on) {
    g_on = on != 0;
    return 0;
}

int sefd_prof_reset(void) {
    for (Rec& r : g_recs) {
        g_pool.push_back(r.e0);
        g_pool.push_back(r.e1);
    }
    g_recs.clear();
    return 0;
}

int sefd_prof_get(int cat, double* ms, long long* launches, double* flops, double* bytes) {
    SEFD_REQUIRE(cat >= 0 && cat < SEFD_PROF_NCAT, "prof_get: bad category %d", cat);
    cudaError_t e = cudaDeviceSynchronize();
    SEFD_REQUIRE(e == cudaSuccess, "prof_get: %s", cudaGetErrorString(e));
    double t = 0, f = 0, b = 0;
    long long n = 0;
    for (const Rec& r : g_recs)
        if (r.cat == cat) {
            float dt = 0.f;
            cudaEventElapsedTime(&dt, r.e0, r.e1);
            t += dt;
            f += r.flops;
            b += r.bytes;
            ++n;
        }
    *ms = t; *launches = n; *flops = f; *bytes = b;
    return 0;
}

}  // extern "C"
