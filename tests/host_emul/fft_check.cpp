// CPU check of the Stockham pass index math in ofdm_uhd_b200/csrc/fft.cuh:
// every pass is run for all thread ids in sequence (ping-pong buffers make that exact).
#define OFDM_HOST_EMUL
#include "../../ofdm_uhd_b200/csrc/fft.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <cmath>

template <int N, int S> double check() {
    using P = FftPlan<N>;
    constexpr int T = N / P::E;
    std::vector<float2> tw(fft_twiddle_elems(N), make_float2(0.f, 0.f)), x(N), out(N), A(fft_smem_elems<N>()), B(fft_smem_elems<N>());
    fft_fill_twiddles<N>(tw.data());
    for (int i = 0; i < N; ++i) x[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    auto ld = [&](int i, int) { return x[i]; };
    auto st = [&](int i, float2 v, int) { out[i] = v; };
    constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2];
    for (int t = 0; t < T; ++t) fft_pass<N, R0, 1, S>(t, tw.data(), ld, SmemOut{A.data()});
    if (P::NP == 2) {
        for (int t = 0; t < T; ++t) fft_pass<N, R1, R0, S>(t, tw.data(), SmemIn{A.data()}, st);
    } else {
        for (int t = 0; t < T; ++t) fft_pass<N, R1, R0, S>(t, tw.data(), SmemIn{A.data()}, SmemOut{B.data()});
        for (int t = 0; t < T; ++t) fft_pass<N, (R2 > 1 ? R2 : 2), R0 * R1, S>(t, tw.data(), SmemIn{B.data()}, st);
    }
    double num = 0, den = 0;
    for (int k = 0; k < N; ++k) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n)
            acc += std::complex<double>(x[n].x, x[n].y) * std::polar(1.0, S * 2 * M_PI * (double)((long)k * n % N) / N);
        num += std::norm(acc - std::complex<double>(out[k].x, out[k].y));
        den += std::norm(acc);
    }
    return sqrt(num / den);
}

// warp plans (FftPlanW: 32 points per thread, radix-32 first pass, padding i + i/32)


template <int N, int S> double check_w() {
    using P = typename FftPlanW<N>::type;
    constexpr int T = N / P::E;
    std::vector<float2> tw(fft_twiddle_elems(N), make_float2(0.f, 0.f)), x(N), out(N), A(FFT_PAD32(N) + 2);
    fft_fill_twiddles<N, P>(tw.data());
    for (int i = 0; i < N; ++i) x[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    auto ld = [&](int i, int) { return x[i]; };
    auto st = [&](int i, float2 v, int) { out[i] = v; };
    for (int t = 0; t < T; ++t) fft_pass<N, P::R[0], 1, S, decltype(ld), SmemOut32, false, P>(t, tw.data(), ld, SmemOut32{A.data()});
    for (int t = 0; t < T; ++t) fft_pass<N, P::R[1], P::R[0], S, SmemIn32, decltype(st), false, P>(t, tw.data(), SmemIn32{A.data()}, st);
    double num = 0, den = 0;
    for (int k = 0; k < N; ++k) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n)
            acc += std::complex<double>(x[n].x, x[n].y) * std::polar(1.0, S * 2 * M_PI * (double)((long)k * n % N) / N);
        num += std::norm(acc - std::complex<double>(out[k].x, out[k].y));
        den += std::norm(acc);
    }
    return sqrt(num / den);
}

int main() {
    double worst = 0;
#define RUN(N) { double a = check<N, -1>(), b = check<N, 1>(); printf("N=%d fwd %.3g inv %.3g\n", N, a, b); worst = fmax(worst, fmax(a, b)); }
    RUN(64) RUN(128) RUN(256) RUN(512) RUN(1024) RUN(2048) RUN(4096)
#define RUNW(N) { double a = check_w<N, -1>(), b = check_w<N, 1>(); printf("warp plan N=%d fwd %.3g inv %.3g\n", N, a, b); worst = fmax(worst, fmax(a, b)); }
    RUNW(512) RUNW(1024)
    printf("worst %.3g\n", worst);
    return worst < 2e-6 ? 0 : 1;
}
