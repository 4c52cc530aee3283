#include <stdarg.h>
#include <stdio.h>

#include <vector>

#include "../../include/sefd.h"
#include "common.cuh"
#include "prof.cuh"

namespace {
struct Rec {
    int cat;
    cudaEvent_t e0, e1;
    double flops, bytes;
    char label[96];
};
char g_label[96] = "";
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

namespace {
int g_stale = 0;
char g_stale_msg[160] = "";
}  // namespace
void sefd_absorb_stale_error() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ++g_stale;
        snprintf(g_stale_msg, sizeof(g_stale_msg), "%s (%s)", cudaGetErrorString(e), cudaGetErrorName(e));
    }
}
extern "C" int sefd_stale_cuda_errors(void) { return g_stale; }
extern "C" const char* sefd_last_stale_cuda_error(void) { return g_stale_msg; }

void sefd_prof_label(const char* fmt, ...) {
    if (!g_on) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_label, sizeof(g_label), fmt, ap);
    va_end(ap);
}

void sefd_prof_push(int cat, double flops, double bytes, cudaStream_t st, bool begin) {
    if (begin) {
        Rec r;
        r.cat = cat;
        r.flops = flops;
        r.bytes = bytes;
        snprintf(r.label, sizeof(r.label), "%s", g_label);
        g_label[0] = 0;
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

int sefd_prof_dump(const char* path) {
    cudaError_t e = cudaDeviceSynchronize();
    SEFD_REQUIRE(e == cudaSuccess, "prof_dump: %s", cudaGetErrorString(e));
    FILE* f = fopen(path, "w");
    SEFD_REQUIRE(f != nullptr, "prof_dump: cannot open %s", path);
    fprintf(f, "idx,category,label,ms,gflop,gbyte,tflops,gbs\n");
    int i = 0;
    for (const Rec& r : g_recs) {
        float dt = 0.f;
        cudaEventElapsedTime(&dt, r.e0, r.e1);
        fprintf(f, "%d,%d,%s,%.4f,%.3f,%.4f,%.2f,%.1f\n", i++, r.cat, r.label, dt, r.flops / 1e9, r.bytes / 1e9,
                dt > 0 ? r.flops / 1e9 / dt : 0.0, dt > 0 ? r.bytes / 1e6 / dt : 0.0);
    }
    fclose(f);
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
