// K_RX2 fused: the whole of upstream ofdm_sync_pn (Schmidl-Cox metric + peak_detector_fb) as ONE streaming
// kernel -- the filtered stream is read once and only trigger indices are written (no timing-metric array).
// Reference wiring: ofdm_receiver.py~:97-101; math: SURVEY.md A.6-A.7; same arithmetic (operation order,
// float64 accumulation, float32 rounding points) as sync_metric_kernel + peak_detect_kernel in rx_front.cu,
// which remain the stage-level entry points (ofdm_rx_sync_metric / ofdm_rx_peak_detect).
//
// One warp walks one contiguous segment, 32*K samples per step with K consecutive samples per lane, where
// 32*K = N/2: a step is exactly one van Herk block, so
//   * y[n - N/2] is the same lane's sample of the previous step (kept in registers, never re-read),
//   * a window sum = [same lane's later elements + later lanes of the previous step] + [earlier lanes + own
//     elements of this step]: two warp scans per sum, no shared memory, no subtraction,
//   * the cp-wide average of the metric is a float64 prefix difference through a per-warp ring in shared memory,
//   * the detector (IIR average scan, threshold ballot, run / arg-max state machine) consumes the K metric values
//     straight from registers.
// A segment starts OFDM_PEAK_WARM samples early (+2 priming steps for the sums) and runs past its end until an
// open run closes; a run belongs to the segment it starts in.
#include "internal.h"
#include "common.cuh"
#include <limits.h>

struct StreamParams {
    const float2* y;
    int64_t n;
    int cp;
    float tapf;
    int64_t seg_len, n_seg;
    int seg_cap;
    int32_t* seg_count;
    int64_t* seg_trig;
    int64_t* first_nan;
    uint32_t* status;
};

constexpr int SS_WARPS = 4;

template <int K>
__global__ void __launch_bounds__(SS_WARPS * 32, 3) sync_stream_kernel(const StreamParams p) {
    constexpr int SZ = 32 * K;                       // samples per step = N/2
    extern __shared__ double s_ring[];               // [SS_WARPS][2*SZ] prefix sums of the metric
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* ring = s_ring + (size_t)wib * 2 * SZ;
    const int64_t seg = (int64_t)blockIdx.x * SS_WARPS + wib;
    if (seg >= p.n_seg) return;
    const int64_t s0 = seg * p.seg_len;
    const int64_t s1 = (s0 + p.seg_len < p.n) ? s0 + p.seg_len : p.n;
    int64_t w0 = s0 - OFDM_PEAK_WARM - 2 * SZ;
    if (w0 < 0) w0 = 0;
    w0 -= w0 % SZ;
    const int cp = p.cp;
    const double tap = (double)p.tapf;
    const bool vec_ok = (((uintptr_t)p.y) & 15) == 0;

    // detector constants
    const double a1 = (double)0.001f, a2 = 1.0 - a1;
    double a2k = 1.0;
#pragma unroll
    for (int k = 0; k < K; ++k) a2k *= a2;
    double pw[5];
    pw[0] = a2k;
#pragma unroll
    for (int k = 1; k < 5; ++k) pw[k] = pw[k - 1] * pw[k - 1];
    const double p32 = pw[4] * pw[4];
    double plane = 1.0;
    for (int k = 0; k < lane; ++k) plane *= a2k;

    // streaming state
    float2 yprev[K];                                 // this lane's samples of the previous step (= y[n - N/2])
    float xp[3][K];                                  // previous step's products (Re c, Im c, |y|^2)
    double pbwd[3] = {0.0, 0.0, 0.0};                // previous step: sum over the later lanes
#pragma unroll
    for (int i = 0; i < K; ++i) {
        yprev[i] = make_float2(0.f, 0.f);
        xp[0][i] = xp[1][i] = xp[2][i] = 0.f;
    }
    if (w0 > 0) {                                    // mid-stream start: the delayed samples exist
#pragma unroll
        for (int i = 0; i < K; ++i) yprev[i] = p.y[w0 - SZ + lane * K + i];
    }
    for (int i = lane; i < 2 * SZ; i += 32) ring[i] = 0.0;
    __syncwarp();
    double carry2 = 0.0;                             // prefix of the metric up to the previous step
    double carry = 0.0;                              // detector average after the last consumed sample
    int state = 0, count = 0;
    float peak = -INFINITY;
    int64_t ind = 0, run_start = 0;
    const int prime = (w0 > 0) ? 2 : 0;              // steps whose sums still miss history
    int step = 0;

    for (int64_t i0 = w0; i0 < p.n; i0 += SZ, ++step) {
        if (i0 >= s1 && state == 0) break;
        const int64_t b0 = i0 + (int64_t)lane * K;
        float2 yv[K];
        if (vec_ok && i0 + SZ <= p.n && (K % 2) == 0) {
            const float4* q = (const float4*)(p.y + b0);
#pragma unroll
            for (int i = 0; i < K / 2; ++i) {
                const float4 t = __ldg(q + i);
                yv[2 * i] = make_float2(t.x, t.y);
                yv[2 * i + 1] = make_float2(t.z, t.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; ++i) yv[i] = (b0 + i < p.n) ? p.y[b0 + i] : make_float2(0.f, 0.f);
        }
        // products of this step
        float x[3][K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float2 c = cmulc_x(yv[i], yprev[i]);          // y[n] * conj(y[n - N/2])
            x[0][i] = c.x; x[1][i] = c.y; x[2][i] = norm_x(yv[i]);
            yprev[i] = yv[i];
        }
        // three moving sums of width N/2 = SZ (van Herk: previous step's tail + this step's head)
        float PR[3][K];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double pre[K];
            double run = 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) { run += (double)x[a][i]; pre[i] = run; }
            // exclusive scans of the lane totals: earlier lanes (fwd), later lanes (bwd)
            double fi = run, bi = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double of = __shfl_up_sync(0xffffffffu, fi, d);
                const double ob = __shfl_down_sync(0xffffffffu, bi, d);
                if (lane >= d) fi += of;
                if (lane + d < 32) bi += ob;
            }
            double fwd = __shfl_up_sync(0xffffffffu, fi, 1);
            double bwd = __shfl_down_sync(0xffffffffu, bi, 1);
            if (lane == 0) fwd = 0.0;
            if (lane == 31) bwd = 0.0;
            // previous step's part of each window: elements i+1.. of this lane, then the later lanes
            double tail[K];
            double sfx = pbwd[a];
            tail[K - 1] = sfx;
#pragma unroll
            for (int i = K - 2; i >= 0; --i) { sfx += (double)xp[a][i + 1]; tail[i] = sfx; }
#pragma unroll
            for (int i = 0; i < K; ++i) {
                PR[a][i] = (float)(tail[i] + (pre[i] + fwd));
                xp[a][i] = x[a][i];
            }
            pbwd[a] = bwd;
        }
        // normalised metric and its cp-wide average (float64 prefix difference)
        float Mt[K];
        double mloc[K];
        double mrun = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float num = fadd_rn(fmul_rn(PR[0][i], PR[0][i]), fmul_rn(PR[1][i], PR[1][i]));
            const float den = fmul_rn(PR[2][i], PR[2][i]);
            // mid-stream start: the first step's sums miss the previous block -> keep the prefix finite
            Mt[i] = (prime && step == 0) ? 0.f : fdiv_rn(num, den);
            mrun += (double)Mt[i];
            mloc[i] = mrun;
        }
        double mi = mrun;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, mi, d);
            if (lane >= d) mi += o;
        }
        double mex = __shfl_up_sync(0xffffffffu, mi, 1);
        if (lane == 0) mex = 0.0;
        const double mtot = __shfl_sync(0xffffffffu, mi, 31);
        const int rb = (step & 1) * SZ;                          // ring half of this step
#pragma unroll
        for (int i = 0; i < K; ++i) { mloc[i] += carry2 + mex; ring[rb + lane * K + i] = mloc[i]; }
        carry2 += mtot;
        __syncwarp();
        float v[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int e = lane * K + i - cp;                     // position of n - cp relative to this step
            double prevS;
            if (e >= 0) prevS = ring[rb + e];
            else prevS = ring[(rb ^ SZ) + SZ + e];               // previous step's half (zeros before the stream)
            const float s = (float)((mloc[i] - prevS) * tap);
            v[i] = fadd_rn(s, -1.0f);
        }
        __syncwarp();
        if (step < prime) continue;                              // priming: sums not yet valid, detector idle
        // first NaN of the metric (poisons the detector for the rest of the stream, C.1)
        if (v[K - 1] != v[K - 1]) {                              // a NaN poisons the prefix: the lane's last value shows it
            int first = K - 1;
#pragma unroll
            for (int i = K - 2; i >= 0; --i)
                if (v[i] != v[i]) first = i;
            if (b0 + first < p.n) atomicMin((unsigned long long*)p.first_nan, (unsigned long long)(b0 + first));
        }
        // ---- peak_detector_fb: IIR average as an affine scan over lanes, threshold bits, run state machine ----
        double vd[K];
#pragma unroll
        for (int i = 0; i < K; ++i) vd[i] = a1 * (double)v[i];
        double loc = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) loc = a2 * loc + vd[i];
        double b = loc;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double o = __shfl_up_sync(0xffffffffu, b, 1 << k);
            if (lane >= (1 << k)) b = b + pw[k] * o;
        }
        double prev = __shfl_up_sync(0xffffffffu, b, 1);
        if (lane == 0) prev = 0.0;
        prev = prev + plane * carry;
        carry = __shfl_sync(0xffffffffu, b, 31) + p32 * carry;
        unsigned mk = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float thr = fmul_rn((float)prev, 0.2f);
            if (b0 + i < p.n && v[i] > thr) mk |= 1u << i;
            prev = a2 * prev + vd[i];
        }
        const unsigned any = __ballot_sync(0xffffffffu, mk != 0);
        if (state == 0 && any == 0) continue;
        // lane summaries for the bulk path: a lane fully inside a run only contributes its maximum
        float lmax = v[0];
        int larg = 0;
#pragma unroll
        for (int i = 1; i < K; ++i)
            if (v[i] > lmax) { lmax = v[i]; larg = i; }
        constexpr unsigned FULL = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
        for (int l = (state == 0 ? __ffs(any) - 1 : 0); l < 32; ++l) {
            const unsigned m = __shfl_sync(0xffffffffu, mk, l);
            if (state == 0 && m == 0) continue;
            const int64_t base = i0 + (int64_t)l * K;
            if (state == 1 && m == FULL && base + K <= p.n) {
                const float bm = __shfl_sync(0xffffffffu, lmax, l);
                const int ba = __shfl_sync(0xffffffffu, larg, l);
                if (bm > peak) { peak = bm; ind = base + ba; }
                continue;
            }
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const float vi = __shfl_sync(0xffffffffu, v[i], l);
                const int64_t idx = base + i;
                if (idx >= p.n) break;
                const bool bit = (m >> i) & 1u;
                if (state == 0) {
                    if (bit) { state = 1; peak = vi; ind = idx; run_start = idx; }
                } else if (vi > peak) {
                    peak = vi; ind = idx;
                } else if (!bit) {
                    if (run_start >= s0 && run_start < s1) {
                        if (count < p.seg_cap) {
                            if (lane == 0) p.seg_trig[seg * p.seg_cap + count] = ind;
                        } else if (lane == 0) {
                            atomicOr(p.status, OFDM_ST_SEG_OVERFLOW);
                        }
                        ++count;
                    }
                    state = 0;
                }
            }
        }
    }
    if (lane == 0) p.seg_count[seg] = count < p.seg_cap ? count : p.seg_cap;
}

__global__ void stream_init_kernel(int64_t* first_nan) { *first_nan = LLONG_MAX; }

template <int K>
static int launch_stream_k(const StreamParams& p, cudaStream_t st) {
    const size_t smem = sizeof(double) * SS_WARPS * 2 * 32 * K;
    static bool attr_done = false;
    if (!attr_done) {
        OFDM_CUDA_CHECK(cudaFuncSetAttribute(sync_stream_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    sync_stream_kernel<K><<<(unsigned)((p.n_seg + SS_WARPS - 1) / SS_WARPS), SS_WARPS * 32, smem, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// returns 1 if the fused kernel does not apply (the caller then runs the two-kernel path): layouts outside
// 32*K = N/2, cp > N/2, or -- unless force -- streams too short to give every SM a few warps (one warp per
// >= 65 536-sample segment), where the tile-parallel metric kernel is the faster choice
int launch_sync_stream(ofdm_handle* h, const float2* y, int64_t n, ofdm_rx_io* io, RxWorkspace* ws, int force,
                       cudaStream_t st) {
    const int K = h->N / 64;                                     // 32*K = N/2
    if (!(K == 2 || K == 4 || K == 8) || h->cp > 32 * K || ws->n_seg == 0) return 1;
    if (!force && ws->n_seg < 148 * 8) return 1;
    StreamParams p;
    p.y = y; p.n = n; p.cp = h->cp; p.tapf = (float)(1.0 / (double)h->cp);
    p.seg_len = ws->seg_len; p.n_seg = ws->n_seg; p.seg_cap = (int)ws->seg_cap;
    p.seg_count = ws->seg_count; p.seg_trig = ws->seg_trig; p.first_nan = ws->first_nan; p.status = io->status;
    stream_init_kernel<<<1, 1, 0, st>>>(p.first_nan);
    OFDM_LAUNCH_CHECK();
    switch (K) {
        case 2: return launch_stream_k<2>(p, st);
        case 4: return launch_stream_k<4>(p, st);
        default: return launch_stream_k<8>(p, st);
    }
}
