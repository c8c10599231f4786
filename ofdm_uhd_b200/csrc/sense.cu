// Spectrum sensing: stream_to_vector -> fft_vcc(N, True, blackmanharris(N)[, shift]) ->
// complex_to_mag_squared -> bin_statistics_f  (secondary_tx.py:163-202, sensing_and_tramsmitting.py:185-234,
// predictive_sense.py:72-123, usrp_fft_save.py:58-62) and the sense_loop decision (secondary_tx.py:237-266,306-331).
#include "internal.h"
#include <limits.h>
#include "fft.cuh"

struct SenseParams {
    const float2* x;
    const float* win;
    const float2* tw;
    int64_t n_dwell;       // work items
    int per, tune_delay, dwell_delay;
    int shift;
    float* maxhold;        // [n_dwell][N] or null
    float2* spectra;       // [n_frames][N] or null (dwell == frame)
};

template <int N>
struct SenseLoad {
    const float2* frame;
    const float* win;
    __device__ __forceinline__ float2 operator()(int idx, int) const {
        float2 v = LDG(frame + idx);
        float w = LDG(win + idx);
        return make_float2(fmul_rn(v.x, w), fmul_rn(v.y, w));
    }
};

// keeps the |X|^2 max-hold of the thread's own E bins in registers
template <int N, int E>
struct SenseStore {
    float* mx;             // per-thread register array [E], indexed by the pass's register slot
    float2* spec;          // or raw spectrum row
    int shift;
    __device__ __forceinline__ void operator()(int idx, float2 v, int slot) const {
        if (spec) {
            int k = shift ? ((idx + N / 2) & (N - 1)) : idx;
            spec[k] = v;
        } else {
            float p = fadd_rn(fmul_rn(v.x, v.x), fmul_rn(v.y, v.y));
            mx[slot] = fmaxf(mx[slot], p);
        }
    }
};

template <int N, int G>
__global__ void __launch_bounds__(G * (N / FftPlan<N>::E)) sense_kernel(const SenseParams p) {
    using P = FftPlan<N>;
    constexpr int E = P::E;
    constexpr int T = N / E;
    constexpr int SB = fft_smem_elems<N>();
    extern __shared__ float2 smem[];
    const int g = threadIdx.x / T;
    const int tid = threadIdx.x - g * T;
    float2* bufA = smem + (size_t)g * 2 * SB;
    float2* bufB = bufA + SB;
    auto bar = [] { __syncthreads(); };
    constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2];
    for (int64_t base = (int64_t)blockIdx.x * G; base < p.n_dwell; base += (int64_t)gridDim.x * G) {
        const int64_t d = base + g;
        const bool active = d < p.n_dwell;
        float mx[E];
#pragma unroll
        for (int i = 0; i < E; ++i) mx[i] = 0.f;         // bin_statistics_f starts its max at 0
        for (int fr = 0; fr < p.dwell_delay; ++fr) {
            const int64_t frame = d * p.per + p.tune_delay + fr;
            SenseLoad<N> ld{p.x + (active ? frame : 0) * N, p.win};
            SenseStore<N, E> st{mx, (active && p.spectra) ? p.spectra + frame * N : nullptr, p.shift};
            if (active) fft_pass<N, R0, 1, -1>(tid, p.tw, ld, SmemOut{bufA});
            bar();
            if constexpr (P::NP == 2) {
                if (active) fft_pass<N, R1, R0, -1>(tid, p.tw, SmemIn{bufA}, st);
                bar();
            } else {
                if (active) fft_pass<N, R1, R0, -1>(tid, p.tw, SmemIn{bufA}, SmemOut{bufB});
                bar();
                if (active) fft_pass<N, R2, R0 * R1, -1>(tid, p.tw, SmemIn{bufB}, st);
            }
        }
        if (active && p.maxhold) {
            float* row = p.maxhold + d * N;
            constexpr int RL = P::R[P::NP - 1];            // last-pass radix: slot q*RL + r <-> bin tid + q*T + r*N/RL
#pragma unroll
            for (int s = 0; s < E; ++s) {
                int idx = tid + (s / RL) * T + (s % RL) * (N / RL);
                int k = p.shift ? ((idx + N / 2) & (N - 1)) : idx;
                row[k] = mx[s];
            }
        }
    }
}

template <int N, int G>
static int launch_sense_n(ofdm_sense_handle* s, const SenseParams& p, cudaStream_t st) {
    constexpr int T = N / FftPlan<N>::E;
    size_t smem = ((size_t)G * 2 * fft_smem_elems<N>()) * sizeof(float2);
    OFDM_SET_MAX_SMEM((sense_kernel<N, G>), smem, s->device);
    const int sms = s->sms;
    int64_t want = (p.n_dwell + G - 1) / G;
    int64_t cap = (int64_t)sms * 16;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) return OFDM_OK;
    sense_kernel<N, G><<<grid, G * T, smem, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

int launch_sense(ofdm_sense_handle* s, const float2* x, int64_t n_frames, int shift, int32_t tune_delay,
                 int32_t dwell_delay, float* maxhold, float2* spectra, cudaStream_t st) {
    SenseParams p;
    p.x = x; p.win = s->d_win; p.tw = s->d_tw; p.per = tune_delay + dwell_delay; p.tune_delay = tune_delay;
    p.dwell_delay = dwell_delay; p.shift = shift; p.maxhold = maxhold; p.spectra = spectra;
    p.n_dwell = n_frames / p.per;
    switch (s->N) {
        case 64:   return launch_sense_n<64, 8>(s, p, st);
        case 128:  return launch_sense_n<128, 8>(s, p, st);
        case 256:  return launch_sense_n<256, 8>(s, p, st);
        case 512:  return launch_sense_n<512, 4>(s, p, st);
        case 1024: return launch_sense_n<1024, 4>(s, p, st);
        case 2048: return launch_sense_n<2048, 2>(s, p, st);
        case 4096: return launch_sense_n<4096, 1>(s, p, st);
    }
    return OFDM_E_INVAL;
}

// sense_loop decision: one thread per bin, Python-float (double) arithmetic.
__global__ void sense_decide_kernel(const float* __restrict__ maxhold, int N, int n_avg, double thr,
                                    double* __restrict__ avg_inorder, uint8_t* __restrict__ free_bits,
                                    char* __restrict__ hex) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;       // nibble index in frequency order
    if (q >= N / 4) return;
    int v = 0;
    for (int j = 0; j < 4; ++j) {
        int io = 4 * q + j;                               // in-order bin
        int k = (io + N / 2) % N;                         // bin of the unshifted FFT (halves swapped)
        double acc = 0.0;
        for (int a = 0; a < n_avg; ++a) acc = acc + (double)maxhold[(size_t)a * N + k];
        double avg = acc / (double)n_avg;
        int fr = (avg > thr) ? 0 : 1;
        if (avg_inorder) avg_inorder[io] = avg;
        if (free_bits) free_bits[io] = (uint8_t)fr;
        v |= fr << j;                                     // first bit = LSB (hex_conv)
    }
    if (hex) hex[q] = "0123456789ABCDEF"[v];
}

int launch_sense_decide(ofdm_sense_handle* s, const float* maxhold, int32_t n_avg, double threshold,
                        double* avg_inorder, uint8_t* free_bits, char* hex, cudaStream_t st) {
    int nq = s->N / 4;
    sense_decide_kernel<<<(nq + 127) / 128, 128, 0, st>>>(maxhold, s->N, n_avg, threshold, avg_inorder, free_bits, hex);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// sense_loop hop decision (secondary_tx.py:268-295): occupied bins in thrshold_inorder[ri-16 : ri+16] (Python slice
// semantics: negative bounds wrap once, then clamp) and the centre (+8) of the quietest 17-bin window over
// i in [200, N-217), each window summed left to right in double like the reference's inner loop, first strict
// minimum below 50 wins.  out = {busy, index or -1, window length}.
__global__ void __launch_bounds__(1024) sense_hop_kernel(const double* __restrict__ avg_inorder,
                                                         const uint8_t* __restrict__ free_bits, int N, int ri,
                                                         int32_t* __restrict__ out) {
    __shared__ double s_p[32];
    __shared__ int s_i[32];
    __shared__ int s_busy;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_busy = 0;
    __syncthreads();
    int start = ri - 16, stop = ri + 16;
    if (start < 0) { start += N; if (start < 0) start = 0; }
    if (stop < 0) { stop += N; if (stop < 0) stop = 0; }
    if (start > N) start = N;
    if (stop > N) stop = N;
    int busy = 0;
    for (int i = start + tid; i < stop; i += 1024) busy += (free_bits[i] == 0);
    if (busy) atomicAdd(&s_busy, busy);
    double best = 50.0;
    int bi = INT_MAX;
    for (int i = 200 + tid; i < N - 217; i += 1024) {
        double power = 0.0;
        for (int j = 0; j < 17; ++j) power = power + avg_inorder[i + j];
        if (power < best) { best = power; bi = i; }          // i ascends within a thread: first minimum kept
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_p[w] = best; s_i[w] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < 32; ++k)
            if (s_p[k] < best || (s_p[k] == best && s_i[k] < bi)) { best = s_p[k]; bi = s_i[k]; }
        out[0] = s_busy;
        out[1] = (bi == INT_MAX) ? -1 : bi + 8;
        out[2] = stop > start ? stop - start : 0;
    }
}

int launch_sense_hop(ofdm_sense_handle* s, const double* avg_inorder, const uint8_t* free_bits, int32_t required_index,
                     int32_t* out, cudaStream_t st) {
    sense_hop_kernel<<<1, 1024, 0, st>>>(avg_inorder, free_bits, s->N, required_index, out);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}
