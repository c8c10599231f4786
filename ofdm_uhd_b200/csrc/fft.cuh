// Shared-memory Stockham FFT core (no cuFFT), radix 4/8/16 register butterflies.
//
// One FFT of length N is executed by T = N/E cooperating threads, each owning E
// points per pass.  The first pass pulls its points through a Load functor (global
// memory, window/derotation/constellation mapping fused there) and the last pass
// hands its results to a Store functor (scaling, cyclic prefix, |X|^2, ... fused
// there), so only the NP-1 inter-pass exchanges touch shared memory.  Both the
// first-pass loads and the last-pass stores are stride-1 across the threads of a
// warp (index j + r*N/R), i.e. coalesced.
//
// Replaces gr.fft_vcc (FFTW) at /root/reference/ofdm.py:112 (backward, ifftshift in the
// Load functor), ofdm_receiver.py~:126 (forward + shift) and the sensing FFTs at
// secondary_tx.py:165 / usrp_fft_save.py:61.
#pragma once
#include "common.cuh"

#define FFT_PAD(i) ((i) + ((i) >> 4))

template <int N> struct FftPlan;
template <> struct FftPlan<64>   { static constexpr int E = 8,  NP = 2; static constexpr int R[4] = {8, 8, 1, 1}; };
template <> struct FftPlan<128>  { static constexpr int E = 16, NP = 2; static constexpr int R[4] = {8, 16, 1, 1}; };
template <> struct FftPlan<256>  { static constexpr int E = 16, NP = 2; static constexpr int R[4] = {16, 16, 1, 1}; };
template <> struct FftPlan<512>  { static constexpr int E = 8,  NP = 3; static constexpr int R[4] = {8, 8, 8, 1}; };
template <> struct FftPlan<1024> { static constexpr int E = 16, NP = 3; static constexpr int R[4] = {4, 16, 16, 1}; };
template <> struct FftPlan<2048> { static constexpr int E = 16, NP = 3; static constexpr int R[4] = {8, 16, 16, 1}; };
template <> struct FftPlan<4096> { static constexpr int E = 16, NP = 3; static constexpr int R[4] = {16, 16, 16, 1}; };
// Warp-wide plans: 32 points per thread, radix-32 register butterflies, ONE shared-memory exchange per transform and
// at most 32 threads per transform (a warp: __syncwarp() instead of block barriers).  1024 = 32 x 32.
// 512 = 32 x 16 (16 threads per transform, two transforms per warp; the radix-16 pass runs two butterflies per thread).
struct FftPlanW1024 { static constexpr int E = 32, NP = 2; static constexpr int R[4] = {32, 32, 1, 1}; };
struct FftPlanW512 { static constexpr int E = 32, NP = 2; static constexpr int R[4] = {32, 16, 1, 1}; };
template <int N> struct FftPlanW;
template <> struct FftPlanW<512> { using type = FftPlanW512; };
template <> struct FftPlanW<1024> { using type = FftPlanW1024; };
#define FFT_PAD32(i) ((i) + ((i) >> 5))

// ---- twiddle table, one section per pass, laid out for the threads that read it -------------------------------
// A pass of radix R after NS = 2^s points of earlier radices needs, per thread and butterfly, the log2(R) power-of-two
// twiddles w^(r * (j mod NS) * N/(NS*R)), r = 1, 2, 4[, 8].  Read from the plain table exp(-2*pi*i*k/N) those loads are
// strided by r (or hit NS distinct lines) and cost up to 16 L1 wavefronts each -- on kernels whose limit is the
// L1/shared-memory data pipe.  Here section s holds them as [butterfly q][li][thread]: every load is 32 consecutive
// float2.  Section s starts at s * N/2 (a section has at most N/4 * 4 / 2 entries); passes of the forward and of the
// reversed radix order never share an NS with different radices, so one table serves both.
HD constexpr int fft_ilog2(int v) { return v <= 1 ? 0 : 1 + fft_ilog2(v >> 1); }
HD constexpr int fft_twiddle_elems(int N) { return 12 * (N / 2); }

template <int N, class P = FftPlan<N>> inline void fft_fill_twiddles(float2* t) {   // t: fft_twiddle_elems(N) entries, zero-initialised
    constexpr int E = P::E, T = N / E;
    auto fill = [&](int R, int NS) {
        if (NS <= 1 || R <= 1) return;
        const int s = fft_ilog2(NS), LG = fft_ilog2(R);
        for (int q = 0; q < E / R; ++q)
            for (int li = 0; li < LG; ++li)
                for (int tid = 0; tid < T; ++tid) {
                    const int k = (tid + q * T) & (NS - 1);
                    const long idx = (long)(1 << li) * k * (N / (NS * R));
                    const double a = 2.0 * 3.14159265358979323846 * (double)idx / (double)N;
                    t[(size_t)s * (N / 2) + (size_t)(q * LG + li) * T + tid] = make_float2((float)cos(a), (float)(-sin(a)));
                }
    };
    int NS = 1;
    for (int p = 0; p < P::NP; ++p) { fill(P::R[p], NS); NS *= P::R[p]; }
    NS = 1;
    for (int p = P::NP - 1; p >= 0; --p) { fill(P::R[p], NS); NS *= P::R[p]; }
}
inline bool fft_fill_twiddles_n(int N, float2* t) {
    switch (N) {
        case 64: fft_fill_twiddles<64>(t); return true;
        case 128: fft_fill_twiddles<128>(t); return true;
        case 256: fft_fill_twiddles<256>(t); return true;
        case 512: fft_fill_twiddles<512>(t); return true;
        case 1024: fft_fill_twiddles<1024>(t); return true;
        case 2048: fft_fill_twiddles<2048>(t); return true;
        case 4096: fft_fill_twiddles<4096>(t); return true;
    }
    return false;
}

template <int N> HD constexpr int fft_threads() { return N / FftPlan<N>::E; }
template <int N> HD constexpr int fft_smem_elems() { return FFT_PAD(N) + 1; }   // float2 elements per buffer

// One butterfly of the register DFTs: (lo, hi) = e +- w o with the compile-time twiddle w = e^{S*j*2*pi*K/R}, written as
// w = g (1 + j t) (|cos| >= |sin|) or w = g (u + j): the scaled product is one packed FMA and the two outputs are one
// packed FMA each (common.cuh).  w = 1 and w = +-j stay plain adds.
HD constexpr double fft_cos32(int q) {                 // cos(2 pi q / 32), q in [0, 16]
    constexpr double C[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                             0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                             0.19509032201612826785, 0.0};
    return q <= 8 ? C[q] : -C[16 - q];
}
template <int R, int K, int S> HD void tw_butterfly(float2 e, float2 o, float2& lo, float2& hi) {
    constexpr int Q = (32 / R) * K;                   // position on the 32-point circle, 0 <= Q < 16
    if constexpr (Q == 0) {
        lo = cadd(e, o); hi = csub(e, o);
    } else if constexpr (Q == 8) {
        const float2 t = S < 0 ? make_float2(o.y, -o.x) : make_float2(-o.y, o.x);
        lo = cadd(e, t); hi = csub(e, t);
    } else {
        constexpr double c = fft_cos32(Q), s = (double)S * fft_cos32(Q <= 8 ? 8 - Q : Q - 8);      // w = c + j s
        if constexpr (Q <= 4 || Q >= 12) {            // |cos| >= |sin|
            constexpr float g = (float)c, t = (float)(s / c);
            const float2 p = cmul_1jt(o, t);
            lo = caxpy(e, g, p); hi = caxpy(e, -g, p);
        } else {
            constexpr float g = (float)s, u = (float)(c / s);
            const float2 p = cmul_uj(o, u);
            lo = caxpy(e, g, p); hi = caxpy(e, -g, p);
        }
    }
}

template <int R, int S> struct DftReg;
template <int S> struct DftReg<2, S> {
    HDM static void run(float2* v) { float2 a = v[0], b = v[1]; v[0] = cadd(a, b); v[1] = csub(a, b); }
};
template <int R, int S, int K> struct DftCombine {
    HDM static void run(float2* v, const float2* e, const float2* o) {
        tw_butterfly<R, K, S>(e[K], o[K], v[K], v[K + R / 2]);
        DftCombine<R, S, K + 1>::run(v, e, o);
    }
};
template <int S> struct DftCombine<4, S, 2> { HDM static void run(float2*, const float2*, const float2*) {} };
template <int S> struct DftCombine<8, S, 4> { HDM static void run(float2*, const float2*, const float2*) {} };
template <int S> struct DftCombine<16, S, 8> { HDM static void run(float2*, const float2*, const float2*) {} };
template <int S> struct DftCombine<32, S, 16> { HDM static void run(float2*, const float2*, const float2*) {} };

// in-register DFT of size R (decimation in time, natural order in and out)
template <int R, int S> struct DftReg {
    HDM static void run(float2* v) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        DftReg<R / 2, S>::run(e);
        DftReg<R / 2, S>::run(o);
        DftCombine<R, S, 0>::run(v, e, o);
    }
};

// One Stockham pass of radix R with Ns = product of the previous radices.
//   in(j + r*N/R), r<R  ->  twiddle by w^{r*(j mod Ns)*N/(Ns*R)}  ->  DFT_R  ->  out((j/Ns)*Ns*R + j mod Ns + r*Ns)
// Functors also receive the compile-time register slot q*R + r of the point, so a caller can keep
// first-pass inputs / last-pass outputs in a register array (slot <-> index tid + q*T + r*N/R).
// BAR: __syncthreads() between the loads and the stores of the pass, for passes that read and write the SAME buffer
// (allowed only when a thread's E points form one butterfly, E == R, so that every load precedes every store).
// WPRE (warp plans: the transform's threads sit in one warp): ALL E points of the thread are read first, then a
// __syncwarp(), then the butterflies and their stores -- the outputs may overwrite the buffer the inputs came from.
template <int N, int R, int NS, int S, class In, class Out, bool BAR = false, class P = FftPlan<N>, bool WPRE = false>
HD void fft_pass(int tid, const float2* __restrict__ tw, In in, Out out) {
    static_assert(!BAR || P::E == R, "in-place pass needs one butterfly per thread");
    constexpr int E = P::E;
    constexpr int T = N / E;
    float2 pre[WPRE ? E : 1];
    if (WPRE) {
#pragma unroll
        for (int q = 0; q < E / R; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) pre[WPRE ? q * R + r : 0] = in(tid + q * T + r * (N / R), q * R + r);
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
#pragma unroll
    for (int q = 0; q < E / R; ++q) {
        const int j = tid + q * T;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = WPRE ? pre[WPRE ? q * R + r : 0] : in(j + r * (N / R), q * R + r);
#if defined(__CUDA_ARCH__)
        if (BAR) __syncthreads();
#endif
        const int k = j & (NS - 1);
        if (NS > 1) {
            // twiddles w^r, r < R: only the power-of-two exponents are loaded (3-4 coalesced loads from this pass's
            // section of the table); the others are products of two loaded ones
            float2 w[R];
            constexpr int LG = fft_ilog2(R);
            const float2* twp = tw + fft_ilog2(NS) * (N / 2) + q * LG * T + tid;    // this pass, this butterfly, this thread
#pragma unroll
            for (int li = 0; li < LG; ++li) {
                w[1 << li] = LDG(twp + li * T);
                if (S > 0) w[1 << li].y = -w[1 << li].y;
            }
#pragma unroll
            for (int r = 3; r < R; ++r)
                if (r & (r - 1)) {                                   // not a power of two
                    const int hi = (r >= 16) ? 16 : ((r >= 8) ? 8 : ((r >= 4) ? 4 : 2));
                    w[r] = cmul(w[hi], w[r - hi]);
                }
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
        }
        DftReg<R, S>::run(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out(j0 + r * NS, v[r], q * R + r);
    }
}

struct SmemIn {
    const float2* p;
    HDM float2 operator()(int i, int) const { return p[FFT_PAD(i)]; }
};
struct SmemOut {
    float2* p;
    HDM void operator()(int i, float2 v, int) const { p[FFT_PAD(i)] = v; }
};
// exchange buffer of the warp plans (padding i + i/32: conflict-free 8-byte accesses in both directions)
struct SmemIn32 {
    const float2* p;
    HDM float2 operator()(int i, int) const { return p[FFT_PAD32(i)]; }
};
struct SmemOut32 {
    float2* p;
    HDM void operator()(int i, float2 v, int) const { p[FFT_PAD32(i)] = v; }
};
