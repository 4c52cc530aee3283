// CPU unit test of the index algebra in fft512.cuh (the same source the CUDA kernels compile).
// Emulates the 64 threads of one transform group; a barrier = finishing the loop over threads.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fft512.cuh"

template <bool INV>
static void run(std::vector<float2>& s, const std::vector<float2>& tw) {
    float2 v[64][8];
    for (int t = 0; t < 64; ++t) fft512_pass1<INV>(s.data(), tw.data(), t, v[t]);
    for (int t = 0; t < 64; ++t) fft512_scatter1(s.data(), t, v[t]);
    for (int t = 0; t < 64; ++t) fft512_pass2<INV>(s.data(), tw.data(), t, v[t]);
    for (int t = 0; t < 64; ++t) fft512_scatter2(s.data(), t, v[t]);
    for (int t = 0; t < 64; ++t) fft512_pass3<INV>(s.data(), t, v[t]);
    for (int t = 0; t < 64; ++t) fft512_scatter3(s.data(), t, v[t]);
}

int main() {
    const int N = 512;
    std::vector<float2> tw(N), x(N);
    for (int j = 0; j < N; ++j) tw[j] = make_float2((float)cos(2 * M_PI * j / N), (float)-sin(2 * M_PI * j / N));
    srand(1);
    for (int j = 0; j < N; ++j) x[j] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    double worst = 0;
    for (int inv = 0; inv < 2; ++inv) {
        std::vector<float2> s(N);
        for (int j = 0; j < N; ++j) s[fft_at(j)] = x[j];
        if (inv) run<true>(s, tw); else run<false>(s, tw);
        for (int k = 0; k < N; ++k) {
            std::complex<double> acc = 0;
            for (int n = 0; n < N; ++n)
                acc += std::complex<double>(x[n].x, x[n].y) * std::polar(1.0, (inv ? 2 : -2) * M_PI * ((long)n * k % N) / N);
            double e = std::abs(acc - std::complex<double>(s[fft_at(k)].x, s[fft_at(k)].y));
            if (e > worst) worst = e;
        }
    }
    printf("max abs err %.3e\n", worst);
    return worst < 2e-4 ? 0 : 1;
}
