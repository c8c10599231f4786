// K_RX2: the whole of upstream ofdm_sync_pn (Schmidl-Cox metric + peak_detector_fb) as two streaming kernels.
// Reference wiring: ofdm_receiver.py~:97-101; math: SURVEY.md A.6-A.7; same arithmetic (operation order,
// float64 accumulation, float32 rounding points) as sync_metric_kernel + peak_detect_kernel in rx_front.cu,
// which remain the stage-level entry points (ofdm_rx_sync_metric / ofdm_rx_peak_detect) and serve N = 64 / cp > N/2.
//
// Both kernels run one warp per CTA (the metric kernel of N >= 1024: N/512 warps, see metric_multi_kernel) (segment index and loop bounds are then CTA-uniform, so the compiler can
// prove every shuffle convergent) and walk a contiguous piece of the stream, 32*K samples per step with K
// consecutive samples per lane, where 32*K = N/2: a step is exactly one van Herk block, so
//   * y[n - N/2] is the same lane's sample of the previous step (kept in registers, never re-read),
//   * a window sum = [same lane's later elements + later lanes of the previous step] + [earlier lanes + own
//     elements of this step]: both exclusive lane scans of a sum from one recursive-doubling pass (bfly_scan: five
//     exchanges), no shared memory, no subtraction,
//   * the cp-wide average of the metric is a float64 prefix difference through a per-warp ring in shared memory,
//   * the detector (IIR average scan, threshold ballot, run / arg-max state machine) consumes the K metric values
//     straight from registers.
// The window sums and the detector have very different shapes -- the sums need ~170 registers and no history
// beyond one block, the detector ~60 registers but a 24 576-sample warm-up of its IIR average -- so as two kernels
// each runs at its own occupancy, the expensive half skips the warm-up overlap entirely, and the 4 B/sample
// hand-over is noise next to the instruction-issue / XU-pipe limit both halves sit on (a fused single kernel
// measured 5.3 ms on the 640 M-sample bench capture, the pair 3.4 ms).
//   metric_chunk_kernel  one warp per chunk of MC_STEPS blocks (+1 priming block): y -> M = |P|^2 / R^2
//   metric_multi_kernel  the same for N/2 = MW * 256: MW warps per CTA share every block
//   detect_seg_kernel    one warp per detector segment: M -> cp-average - 1 -> peak_detector_fb -> triggers;
//                        a segment starts OFDM_PEAK_WARM samples early and runs past its end until an open run
//                        closes; a run belongs to the segment it starts in.
#include "internal.h"
#include "common.cuh"
#include <limits.h>
#include <stdlib.h>
#include <string.h>

struct StreamParams {
    const float2* y;
    const int64_t* soff;          // stream offsets (nullptr: one stream of n samples); tables below are per stream
    int max_frames;
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

// Warp scans of doubles.  The Kogge-Stone pair (OFDM_METRIC_SCAN=ks, what the round measured before bfly_scan; the
// detector's scans too): a conditional add after a shuffle compiles to DADD + two FSEL; multiplying the shuffled value by
// a per-lane 1.0 / 0.0 mask inside one DFMA is exact (x*1 = x, x*0 = 0 for finite x) and a single instruction.
__device__ __forceinline__ double shfl_up_d(double x, int d) {
    return __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(x), d),
                            __shfl_up_sync(0xffffffffu, __double2loint(x), d));
}
__device__ __forceinline__ double shfl_down_d(double x, int d) {
    return __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(x), d),
                            __shfl_down_sync(0xffffffffu, __double2loint(x), d));
}
__device__ __forceinline__ double shfl_xor_d(double x, int d) {
    return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(x), d),
                            __shfl_xor_sync(0xffffffffu, __double2loint(x), d));
}
// Both exclusive lane scans of one double per lane by recursive doubling: after level k every lane holds the total of
// its 2^(k+1)-lane group, and the partner group's total (ONE exchange) belongs to the lanes in front of this lane if the
// lane sits in the upper half of the group, to the lanes behind it otherwise.  Five exchanges serve both directions and
// leave them exclusive -- the Kogge-Stone pair takes ten and two more to shift.  The routing is an integer AND on the two
// halves of the exchanged double (x + 0.0 is exact, and a NaN / Inf behind a lane never reaches the sum of the lanes in
// front of it, which a multiply by a 0.0 mask would let through); the ANDs run on the ALU pipe, which these kernels leave
// idle.  Additions only (an all-zero group sums to exactly 0); `tot` is the same bit pattern on every lane (IEEE
// addition commutes).
__device__ __forceinline__ void bfly_scan(const double run, const int lane, double& fwd, double& bwd, double& tot) {
    double s = run, f = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(s), 1 << k);
        const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(s), 1 << k);
        const int up = -((lane >> k) & 1);                 // all ones: the partner group is in front of this lane
        f += __hiloint2double(hi & up, lo & up);
        b += __hiloint2double(hi & ~up, lo & ~up);
        s += __hiloint2double(hi, lo);
    }
    fwd = f; bwd = b; tot = s;
}

// 256-bit global accesses (sm_100): a lane owns K consecutive samples, i.e. lanes sit 8K bytes apart, and the L1 data
// pipe -- which the shuffles share -- is charged per 128-byte line touched by an instruction, so a 64-byte run per
// lane costs half as many wavefronts as two LDG.256 than as four LDG.128.
__device__ __forceinline__ void ldg256(const void* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// In-lane inclusive prefix and "elements after i" suffix sums of 8 doubles as depth-3 trees instead of 7-deep chains:
// a warp issues in order, so the chains' latency, not the FP64 rate, paced the metric kernels.
__device__ __forceinline__ void prefix8_tree(const double (&x)[8], double (&pre)[8]) {
    const double s01 = x[0] + x[1], s23 = x[2] + x[3], s45 = x[4] + x[5], s67 = x[6] + x[7];
    const double s0123 = s01 + s23, s4567 = s45 + s67;
    pre[0] = x[0]; pre[1] = s01; pre[2] = s01 + x[2]; pre[3] = s0123;
    pre[4] = s0123 + x[4]; pre[5] = s0123 + s45; pre[6] = pre[5] + x[6]; pre[7] = s0123 + s4567;
}
// tail[i] = b + sum of x[i+1 .. 7]   (b: what follows this lane)
__device__ __forceinline__ void suffix8_tree(const double (&x)[8], const double b, double (&tail)[8]) {
    const double q67 = x[6] + x[7], q45 = x[4] + x[5], q23 = x[2] + x[3];
    const double q4567 = q45 + q67;
    tail[7] = b; tail[6] = b + x[7]; tail[5] = b + q67; tail[4] = b + (x[5] + q67);
    tail[3] = b + q4567; tail[2] = b + (x[3] + q4567); tail[1] = b + (q23 + q4567);
    tail[0] = b + (x[1] + (q23 + q4567));
}

// one step's worth of per-lane history (ping-ponged between two instances so nothing is copied per step)
template <int K>
struct StepHist {
    float2 y[K];                                     // samples            (y[n - N/2] of the next step)
    float x[3][K];                                   // products Re c, Im c, |y|^2
    double bwd[3];                                   // sum of the products of the later lanes
};

constexpr int MC_STEPS = 32;       // blocks per chunk (+1 priming block) of a short capture; long ones take MC_STEPS_LONG
constexpr int MC_STEPS_LONG = 64;

template <int K, bool BF = false>          // BF: bfly_scan instead of the Kogge-Stone scan pair
struct MetricCtx {
    static constexpr int SZ = 32 * K;
    const float2* y;
    float* mt;
    int64_t n;
    int lane;
    bool vec_ok, st_ok;
    double mu[5], md[5];

    __device__ __forceinline__ void load(float2 (&dst)[K], const int64_t i0) const {
        const int64_t b0 = i0 + (int64_t)lane * K;
        if (i0 >= 0 && i0 + SZ <= n && vec_ok) {                 // warp-uniform
            if constexpr (K % 4 == 0) {
#pragma unroll
                for (int i = 0; i < K; i += 4) {
                    float t[8];
                    ldg256(y + b0 + i, t);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[i + j] = make_float2(t[2 * j], t[2 * j + 1]);
                }
            } else {
                const float4* q = (const float4*)(y + b0);
#pragma unroll
                for (int i = 0; i < K / 2; ++i) {
                    const float4 t = __ldg(q + i);
                    dst[2 * i] = make_float2(t.x, t.y);
                    dst[2 * i + 1] = make_float2(t.z, t.w);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; ++i) dst[i] = (b0 + i >= 0 && b0 + i < n) ? y[b0 + i] : make_float2(0.f, 0.f);
        }
    }
    __device__ __forceinline__ void products(const StepHist<K>& prev, StepHist<K>& cur) const {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float2 c = cmulc_x(cur.y[i], prev.y[i]);      // y[n] * conj(y[n - N/2])
            cur.x[0][i] = c.x; cur.x[1][i] = c.y; cur.x[2][i] = norm_x(cur.y[i]);
        }
    }
    // priming block: only what the next block needs of it (products and the sums of the later lanes)
    __device__ __forceinline__ void prime(const StepHist<K>& prev, StepHist<K>& cur) const {
        products(prev, cur);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double run = 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) run += (double)cur.x[a][i];
            if constexpr (BF) {
                double fwd, tot;
                bfly_scan(run, lane, fwd, cur.bwd[a], tot);
            } else {
                double bi = run;
#pragma unroll
                for (int k = 0; k < 5; ++k) bi = fma(shfl_down_d(bi, 1 << k), md[k], bi);
                cur.bwd[a] = shfl_down_d(bi, 1) * md[0];
            }
        }
    }
    __device__ __forceinline__ void step(StepHist<K>& prev, StepHist<K>& cur, const int64_t i0, const bool more) {
        products(prev, cur);
        if (more) load(prev.y, i0 + SZ);                         // next block's samples into the dead buffer
        float PR[3][K];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double pre[K], tail[K];
            double run;
            if constexpr (K == 8) {
                double xd[8], pd[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { xd[i] = (double)cur.x[a][i]; pd[i] = (double)prev.x[a][i]; }
                prefix8_tree(xd, pre);
                run = pre[7];
                suffix8_tree(pd, prev.bwd[a], tail);
            } else {
                run = 0.0;
#pragma unroll
                for (int i = 0; i < K; ++i) { run += (double)cur.x[a][i]; pre[i] = run; }
                double sfx = prev.bwd[a];
                tail[K - 1] = sfx;
#pragma unroll
                for (int i = K - 2; i >= 0; --i) { sfx += (double)prev.x[a][i + 1]; tail[i] = sfx; }
            }
            double fwd;
            if constexpr (BF) {
                double tot;
                bfly_scan(run, lane, fwd, cur.bwd[a], tot);
            } else {
                double fi = run, bi = run;                       // inclusive scans: earlier lanes, later lanes
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    fi = fma(shfl_up_d(fi, 1 << k), mu[k], fi);
                    bi = fma(shfl_down_d(bi, 1 << k), md[k], bi);
                }
                fwd = shfl_up_d(fi, 1) * mu[0];
                cur.bwd[a] = shfl_down_d(bi, 1) * md[0];
            }
#pragma unroll
            for (int i = 0; i < K; ++i) PR[a][i] = (float)(tail[i] + (pre[i] + fwd));
        }
        float q[K], num[K], den[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            num[i] = fadd_rn(fmul_rn(PR[0][i], PR[0][i]), fmul_rn(PR[1][i], PR[1][i]));
            den[i] = fmul_rn(PR[2][i], PR[2][i]);
        }
        fdiv_block<K>(num, den, q);                              // all lanes of the warp are here (CTA-uniform loops)
        const int64_t b0 = i0 + (int64_t)lane * K;
        if (i0 + SZ <= n && st_ok) {
            if constexpr (K == 8) {
                stg256(mt + b0, q);
            } else {
#pragma unroll
                for (int i = 0; i < K; i += 4) *(float4*)(mt + b0 + i) = make_float4(q[i], q[i + 1], q[i + 2], q[i + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (b0 + i < n) mt[b0 + i] = q[i];
        }
    }

    // ---- N/2 = MW * 256: MW warps of one CTA share a van Herk block, warp w owning its w-th quarter ---------------
    // A window is still [tail of the previous block] + [head of this one]; inside a warp the sums are the ones of step(),
    // and what the other warps contribute is a warp-uniform offset: head += totals of the earlier warps of THIS block,
    // tail += totals of the later warps of the PREVIOUS block.  s_tot[slot][array][warp] holds the per-warp totals of
    // a block; slots rotate over three blocks, so one __syncthreads per step orders writers and readers.
    template <int MW>
    __device__ __forceinline__ void prime_multi(const StepHist<K>& prev, StepHist<K>& cur, double (*s_tot)[3][MW], const int slot,
                                                const int w) const {
        products(prev, cur);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double run = 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) run += (double)cur.x[a][i];
            double bi;                                       // the warp's total
            if constexpr (BF) {
                double fwd;
                bfly_scan(run, lane, fwd, cur.bwd[a], bi);
            } else {
                bi = run;
#pragma unroll
                for (int k = 0; k < 5; ++k) bi = fma(shfl_down_d(bi, 1 << k), md[k], bi);
                cur.bwd[a] = shfl_down_d(bi, 1) * md[0];
            }
            if (lane == 0) s_tot[slot][a][w] = bi;           // lane 0's inclusive backward scan = the warp's total
        }
    }
    template <int MW>
    __device__ __forceinline__ void step_multi(StepHist<K>& prev, StepHist<K>& cur, const int64_t i0, const bool more,
                                               double (*s_tot)[3][MW], const int slot, const int pslot, const int w) {
        products(prev, cur);
        if (more) load(prev.y, i0 + (int64_t)MW * SZ);           // the next block's samples of this warp into the dead buffer
        float PR[3][K];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            // one array at a time (its 32 doubles of prefixes / tails are the register budget), a block barrier per array
            double pre[K], tail[K];
            double run;
            if constexpr (K == 8) {
                double xd[8], pd[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { xd[i] = (double)cur.x[a][i]; pd[i] = (double)prev.x[a][i]; }
                prefix8_tree(xd, pre);
                run = pre[7];
                suffix8_tree(pd, prev.bwd[a], tail);
            } else {
                run = 0.0;
#pragma unroll
                for (int i = 0; i < K; ++i) { run += (double)cur.x[a][i]; pre[i] = run; }
                double sfx = prev.bwd[a];
                tail[K - 1] = sfx;
#pragma unroll
                for (int i = K - 2; i >= 0; --i) { sfx += (double)prev.x[a][i + 1]; tail[i] = sfx; }
            }
            double fwd, bi;
            if constexpr (BF) {
                bfly_scan(run, lane, fwd, cur.bwd[a], bi);
            } else {
                double fi = run;
                bi = run;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    fi = fma(shfl_up_d(fi, 1 << k), mu[k], fi);
                    bi = fma(shfl_down_d(bi, 1 << k), md[k], bi);
                }
                fwd = shfl_up_d(fi, 1) * mu[0];
                cur.bwd[a] = shfl_down_d(bi, 1) * md[0];
            }
            if (lane == 0) s_tot[slot][a][w] = bi;
            __syncthreads();
            double head = 0.0, tl = 0.0;
#pragma unroll
            for (int j = 0; j < MW; ++j) {
                if (j < w) head += s_tot[slot][a][j];
                if (j > w) tl += s_tot[pslot][a][j];
            }
#pragma unroll
            for (int i = 0; i < K; ++i) PR[a][i] = (float)((tail[i] + tl) + ((pre[i] + fwd) + head));
        }
        float q[K], num[K], den[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            num[i] = fadd_rn(fmul_rn(PR[0][i], PR[0][i]), fmul_rn(PR[1][i], PR[1][i]));
            den[i] = fmul_rn(PR[2][i], PR[2][i]);
        }
        fdiv_block<K>(num, den, q);                              // all lanes of the warp are here (CTA-uniform loops)
        const int64_t b0 = i0 + (int64_t)lane * K;
        if (i0 + SZ <= n && st_ok) {
            if constexpr (K == 8) {
                stg256(mt + b0, q);
            } else {
#pragma unroll
                for (int i = 0; i < K; i += 4) *(float4*)(mt + b0 + i) = make_float4(q[i], q[i + 1], q[i + 2], q[i + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (b0 + i < n) mt[b0 + i] = q[i];
        }
    }
};

template <int K, bool BF>
__global__ void __launch_bounds__(32, 12) metric_chunk_kernel(const float2* __restrict__ y, float* __restrict__ mt, const int64_t n_single,
                                                               const int64_t* __restrict__ soff, const int mc_steps) {
    constexpr int SZ = 32 * K;
    static_assert(K % 4 == 0 || K == 2, "K");
    int64_t s_a, n;
    stream_span(soff, blockIdx.y, n_single, s_a, n);        // zero history in front of every stream
    y += s_a;
    mt += s_a;
    MetricCtx<K, BF> c;
    c.y = y; c.mt = mt; c.n = n; c.lane = threadIdx.x;
    c.vec_ok = (((uintptr_t)y) & 31) == 0;
    c.st_ok = (((uintptr_t)mt) & 31) == 0 && (K % 4) == 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        c.mu[k] = (c.lane >= (1 << k)) ? 1.0 : 0.0;
        c.md[k] = (c.lane + (1 << k) < 32) ? 1.0 : 0.0;
    }
    const int64_t i_begin = (int64_t)blockIdx.x * mc_steps * SZ;
    if (i_begin >= n) return;
    const int64_t i_end = (i_begin + (int64_t)mc_steps * SZ < n) ? i_begin + (int64_t)mc_steps * SZ : n;
    StepHist<K> A, B;
    if (i_begin == 0) {                              // zero history in front of the stream
#pragma unroll
        for (int i = 0; i < K; ++i) {
            A.y[i] = make_float2(0.f, 0.f);
            A.x[0][i] = A.x[1][i] = A.x[2][i] = 0.f;
        }
        A.bwd[0] = A.bwd[1] = A.bwd[2] = 0.0;
    } else {                                         // the block in front of the chunk, from the two blocks before it
        c.load(B.y, i_begin - 2 * SZ);
        c.load(A.y, i_begin - SZ);
        c.prime(B, A);
    }
    c.load(B.y, i_begin);
    for (int64_t i0 = i_begin; i0 < i_end;) {
        c.step(A, B, i0, i0 + SZ < i_end);
        i0 += SZ;
        if (i0 >= i_end) break;
        c.step(B, A, i0, i0 + SZ < i_end);
        i0 += SZ;
    }
}

// ---------------------------------------------------------------------------------------------
// N/2 = MW * 256 (N = 1024, 2048, 4096): one CTA of MW warps walks MC_STEPS van Herk blocks of N/2 samples (+1 priming
// block); warp w owns the w-th 256-sample quarter of every block, with its history in registers exactly like the
// one-warp kernel -- the stream is read once.  (Round 1 ran these layouts with one warp per 256-sample sub-step and the
// previous block's products recomputed from y[n-W] and y[n-2W]: three stream reads per sample.)
// ---------------------------------------------------------------------------------------------
template <int K, int MW, bool BF>
__global__ void __launch_bounds__(32 * MW, (10 / MW) > 0 ? (10 / MW) : 1) metric_multi_kernel(const float2* __restrict__ y, float* __restrict__ mt,
                                                                                             const int64_t n_single,
                                                                                             const int64_t* __restrict__ soff, const int mc_steps) {
    constexpr int SZ = 32 * K;
    constexpr int64_t W = (int64_t)MW * SZ;
    __shared__ double s_tot[3][3][MW];
    int64_t s_a, n;
    stream_span(soff, blockIdx.y, n_single, s_a, n);
    y += s_a;
    mt += s_a;
    MetricCtx<K, BF> c;
    c.y = y; c.mt = mt; c.n = n; c.lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    c.vec_ok = (((uintptr_t)y) & 31) == 0;
    c.st_ok = (((uintptr_t)mt) & 31) == 0 && (K % 4) == 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        c.mu[k] = (c.lane >= (1 << k)) ? 1.0 : 0.0;
        c.md[k] = (c.lane + (1 << k) < 32) ? 1.0 : 0.0;
    }
    const int64_t b_begin = (int64_t)blockIdx.x * mc_steps * W;          // first sample of the chunk
    if (b_begin >= n) return;
    const int64_t b_end = (b_begin + (int64_t)mc_steps * W < n) ? b_begin + (int64_t)mc_steps * W : n;
    const int64_t off = (int64_t)w * SZ;                                  // this warp's quarter
    StepHist<K> A, B;
    if (b_begin == 0) {                              // zero history in front of the stream
#pragma unroll
        for (int i = 0; i < K; ++i) {
            A.y[i] = make_float2(0.f, 0.f);
            A.x[0][i] = A.x[1][i] = A.x[2][i] = 0.f;
        }
        A.bwd[0] = A.bwd[1] = A.bwd[2] = 0.0;
        if (c.lane == 0) { s_tot[2][0][w] = 0.0; s_tot[2][1][w] = 0.0; s_tot[2][2][w] = 0.0; }
    } else {                                         // the block in front of the chunk, from the two blocks before it
        c.load(B.y, b_begin - 2 * W + off);
        c.load(A.y, b_begin - W + off);
        c.template prime_multi<MW>(B, A, s_tot, 2, w);
    }
    c.load(B.y, b_begin + off);
    int slot = 0, pslot = 2;
    for (int64_t i0 = b_begin; i0 < b_end;) {
        c.template step_multi<MW>(A, B, i0 + off, i0 + W < b_end, s_tot, slot, pslot, w);
        i0 += W;
        pslot = slot; slot = (slot == 2) ? 0 : slot + 1;
        if (i0 >= b_end) break;
        c.template step_multi<MW>(B, A, i0 + off, i0 + W < b_end, s_tot, slot, pslot, w);
        i0 += W;
        pslot = slot; slot = (slot == 2) ? 0 : slot + 1;
    }
}

template <int K>
struct DetectCtx {
    static constexpr int SZ = 32 * K;
    const StreamParams& p;
    const float* mt;
    double* ring;
    int lane;
    int64_t seg, s0, s1;
    bool vec_ok, cp_aligned;
    int prime, cpq, ring_mask;                       // ring_mask = steps in the ring - 1 (power of two)
    double tap, a1, a2, p32, plane;
    double pw[5], mu[5];
    double carry2 = 0.0;                             // prefix of the metric up to the previous step
    double carry = 0.0;                              // detector average after the last consumed sample
    int state = 0, count = 0;
    float peak = -INFINITY;
    int64_t ind = 0, run_start = 0;

    __device__ __forceinline__ DetectCtx(const StreamParams& p_) : p(p_) {}

    __device__ __forceinline__ void load(float (&dst)[K], const int64_t i0) const {
        const int64_t b0 = i0 + (int64_t)lane * K;
        if (i0 + SZ <= p.n && vec_ok) {
            if constexpr (K == 8) {
                ldg256(mt + b0, dst);
            } else {
#pragma unroll
                for (int i = 0; i < K; i += 4) {
                    const float4 t = __ldg((const float4*)(mt + b0 + i));
                    dst[i] = t.x; dst[i + 1] = t.y; dst[i + 2] = t.z; dst[i + 3] = t.w;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; ++i) dst[i] = (b0 + i < p.n) ? mt[b0 + i] : 0.f;
        }
    }

    // q holds this step's metric on entry and the next step's on return
    __device__ __forceinline__ void step(float (&q)[K], const int64_t i0, const int stepno) {
        const int64_t b0 = i0 + (int64_t)lane * K;
        const bool full = i0 + SZ <= p.n;                        // warp-uniform
        int nvalid = K;
        if (!full) {
            const int64_t r = p.n - b0;
            nvalid = r < 0 ? 0 : (r > K ? K : (int)r);
        }
        double mloc[K];
        double mrun = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) { mrun += (double)q[i]; mloc[i] = mrun; }
        if (i0 + SZ < p.n) load(q, i0 + SZ);
        double mi = mrun;
#pragma unroll
        for (int k = 0; k < 5; ++k) mi = fma(shfl_up_d(mi, 1 << k), mu[k], mi);
        const double mex = fma(shfl_up_d(mi, 1), mu[0], carry2);
        carry2 += __shfl_sync(0xffffffffu, mi, 31);
        // ring layout [half][i][lane]: element i of lane l sits at i*32 + l, so every store and every load below is
        // stride-1 across the warp (a lane-major layout has lanes 64 B apart: 4-way bank conflicts on both sides,
        // which made the shared-memory pipe -- shared with the shuffles -- this kernel's bottleneck)
        const int rb = (stepno & ring_mask) * SZ;                // ring slot of this step
#pragma unroll
        for (int i = 0; i < K; ++i) {
            mloc[i] += mex;
            ring[rb + i * 32 + lane] = mloc[i];
        }
        __syncwarp();
        float v[K];
        if (cp_aligned) {
            // n - cp sits cp/K lanes back, same element index: in this step's slot or sb steps earlier
            const int ql = lane - cpq;
            const int sb = (ql < 0) ? ((31 - ql) >> 5) : 0;
            const double* src = ring + ((stepno - sb) & ring_mask) * SZ + ql + 32 * sb;
#pragma unroll
            for (int i = 0; i < K; ++i) v[i] = fadd_rn((float)((mloc[i] - src[i * 32]) * tap), -1.0f);
        } else {
#pragma unroll
            for (int i = 0; i < K; ++i) {
                int e = lane * K + i - p.cp;                     // position of n - cp relative to this step
                const int sb = (e < 0) ? ((SZ - 1 - e) / SZ) : 0;
                e += sb * SZ;
                const double prevS = ring[((stepno - sb) & ring_mask) * SZ + (e % K) * 32 + e / K];
                v[i] = fadd_rn((float)((mloc[i] - prevS) * tap), -1.0f);
            }
        }
        __syncwarp();
        if (stepno < prime) return;                              // priming: the cp window reaches before the start
        if (v[K - 1] != v[K - 1]) {                              // a NaN poisons the prefix: the lane's last value shows it
            int first = K - 1;
#pragma unroll
            for (int i = K - 2; i >= 0; --i)
                if (v[i] != v[i]) first = i;
            if (first < nvalid) atomicMin((unsigned long long*)p.first_nan, (unsigned long long)(b0 + first));
        }
        double vd[K];
#pragma unroll
        for (int i = 0; i < K; ++i) vd[i] = a1 * (double)v[i];
        double loc = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) loc = a2 * loc + vd[i];
        double b = loc;
#pragma unroll
        for (int k = 0; k < 5; ++k) b = fma(shfl_up_d(b, 1 << k), pw[k], b);
        double pavg = shfl_up_d(b, 1) * mu[0];
        pavg = pavg + plane * carry;
        carry = __shfl_sync(0xffffffffu, b, 31) + p32 * carry;
        unsigned mk = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float thr = fmul_rn((float)pavg, 0.2f);
            if (v[i] > thr) mk |= 1u << i;
            pavg = a2 * pavg + vd[i];
        }
        constexpr unsigned FULLM = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
        if (!full) mk &= (nvalid >= K) ? FULLM : ((1u << nvalid) - 1u);
        const unsigned any = __ballot_sync(0xffffffffu, mk != 0);
        if (state == 0 && any == 0) return;
        float lmax = v[0];
        int larg = 0;
#pragma unroll
        for (int i = 1; i < K; ++i)
            if (v[i] > lmax) { lmax = v[i]; larg = i; }
        for (int l = (state == 0 ? __ffs(any) - 1 : 0); l < 32; ++l) {
            const unsigned m = __shfl_sync(0xffffffffu, mk, l);
            if (state == 0 && m == 0) continue;
            const int64_t base = i0 + (int64_t)l * K;
            if (state == 1 && m == FULLM && base + K <= p.n) {
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
        state = (int)__reduce_or_sync(0xffffffffu, (unsigned)state);
    }
};

template <int K>
__global__ void __launch_bounds__(32, 24) detect_seg_kernel(const StreamParams p_all, const float* __restrict__ mt, const int ring_steps) {
    constexpr int SZ = 32 * K;
    extern __shared__ __align__(16) double s_ring[];
    // the tables of this CTA's stream (blockIdx.y), built field by field (not as a modified copy of the kernel
    // parameter: see demod_kernel in rx_demod.cu); the stream starts with a zero detector average
    StreamParams p;
    int64_t s_a;
    stream_span(p_all.soff, blockIdx.y, p_all.n, s_a, p.n);
    mt += s_a;
    p.y = p_all.y; p.soff = p_all.soff; p.max_frames = p_all.max_frames; p.cp = p_all.cp; p.tapf = p_all.tapf;
    p.seg_len = p_all.seg_len; p.seg_cap = p_all.seg_cap;
    p.seg_count = p_all.seg_count + (int64_t)blockIdx.y * (p_all.n_seg + 1);
    p.seg_trig = p_all.seg_trig + (int64_t)blockIdx.y * p_all.n_seg * p_all.seg_cap;
    p.first_nan = p_all.first_nan + blockIdx.y;
    p.status = p_all.status + blockIdx.y;
    p.n_seg = p.n > 0 ? (p.n + p.seg_len - 1) / p.seg_len : 0;
    DetectCtx<K> c(p);
    c.mt = mt;
    c.lane = threadIdx.x;
    c.ring = s_ring;
    c.seg = (int64_t)blockIdx.x;
    if (c.seg >= p.n_seg) return;
    c.s0 = c.seg * p.seg_len;
    c.s1 = (c.s0 + p.seg_len < p.n) ? c.s0 + p.seg_len : p.n;
    const int back = (p.cp + SZ - 1) / SZ;           // steps the cp window reaches back
    int64_t w0 = c.s0 - OFDM_PEAK_WARM - (int64_t)back * SZ;
    if (w0 < 0) w0 = 0;
    w0 -= w0 % SZ;
    c.ring_mask = ring_steps - 1;
    c.tap = (double)p.tapf;
    c.vec_ok = (((uintptr_t)mt) & 31) == 0 && (K % 4) == 0;
    c.cp_aligned = (p.cp % K) == 0;
    c.cpq = p.cp / K;
    c.a1 = (double)0.001f;
    c.a2 = 1.0 - c.a1;
    double a2k = 1.0;
#pragma unroll
    for (int k = 0; k < K; ++k) a2k *= c.a2;
    c.pw[0] = a2k;
#pragma unroll
    for (int k = 1; k < 5; ++k) c.pw[k] = c.pw[k - 1] * c.pw[k - 1];
    c.p32 = c.pw[4] * c.pw[4];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        c.mu[k] = (c.lane >= (1 << k)) ? 1.0 : 0.0;
        c.pw[k] *= c.mu[k];
    }
    c.plane = 1.0;
    for (int k = 0; k < c.lane; ++k) c.plane *= a2k;
    for (int i = c.lane; i < ring_steps * SZ; i += 32) c.ring[i] = 0.0;
    __syncwarp();
    c.prime = (w0 > 0) ? back : 0;                   // steps whose cp window reaches in front of w0
    float q[K];
    int step = 0;
    if (w0 < p.n) c.load(q, w0);
    for (int64_t i0 = w0; i0 < p.n; i0 += SZ, ++step) {
        if (i0 >= c.s1 && c.state == 0) break;
        c.step(q, i0, step);
    }
    if (c.lane == 0) p.seg_count[c.seg] = c.count < p.seg_cap ? c.count : p.seg_cap;
}

__global__ void stream_init_kernel(int64_t* first_nan, int S) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) first_nan[s] = LLONG_MAX;
}

template <int K, bool BF>
static void launch_metric(const dim3 g, const int m, const StreamParams& p, float* mt, const int steps, cudaStream_t st) {
    if (m == 1) metric_chunk_kernel<K, BF><<<g, 32, 0, st>>>(p.y, mt, p.n, p.soff, steps);
    else if constexpr (K == 8) {
        if (m == 2) metric_multi_kernel<8, 2, BF><<<g, 64, 0, st>>>(p.y, mt, p.n, p.soff, steps);
        else if (m == 4) metric_multi_kernel<8, 4, BF><<<g, 128, 0, st>>>(p.y, mt, p.n, p.soff, steps);
        else metric_multi_kernel<8, 8, BF><<<g, 256, 0, st>>>(p.y, mt, p.n, p.soff, steps);
    }
}

template <int K>
static int launch_split_k(const StreamParams& p, float* mt, int m, int S, int parts, int sms, cudaStream_t st) {
    constexpr int SZ = 32 * K;
    if (!(parts & 1)) {
    } else {
        // chunk length: a chunk re-runs one priming block, so long captures take 64-block chunks (1.5 % extra work instead
        // of 3 %) as long as that still leaves eight waves of chunks; short ones keep 32 blocks and fill the GPU sooner
        const int64_t W = (int64_t)m * SZ;
        int steps = MC_STEPS_LONG;
        if ((p.n + W * steps - 1) / (W * steps) * S < (int64_t)sms * 12 * 8) steps = MC_STEPS;
        const int64_t chunks = (p.n + W * steps - 1) / (W * steps);
        const dim3 g((unsigned)chunks, S);
        // A/B switch for measurements (read once): OFDM_METRIC_SCAN=ks selects the Kogge-Stone scan pair
        static const bool ks = [] { const char* e = getenv("OFDM_METRIC_SCAN"); return e && !strcmp(e, "ks"); }();
        if ((m != 1 && K != 8) || !(m == 1 || m == 2 || m == 4 || m == 8)) {
            ofdm_set_error("sync: no metric kernel for N/2 = %d x %d", m, SZ);
            return OFDM_E_INVAL;
        }
        if (ks) launch_metric<K, false>(g, m, p, mt, steps, st);
        else launch_metric<K, true>(g, m, p, mt, steps, st);
    }
    OFDM_LAUNCH_CHECK();
    if (!(parts & 2)) return OFDM_OK;
    int ring_steps = 2;
    while (ring_steps < (p.cp + SZ - 1) / SZ + 1) ring_steps *= 2;
    detect_seg_kernel<K><<<dim3((unsigned)p.n_seg, S), 32, sizeof(double) * ring_steps * SZ, st>>>(p, mt, ring_steps);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// returns 1 if the streaming kernels do not apply (the caller then runs the tile-parallel stage kernels, single
// stream only): N = 64 or cp > N/2.  Every stream length takes the streaming pair (one warp per >= 65 536-sample
// detector segment: short streams simply use few warps).
// `parts`: bit 0 the metric kernel, bit 1 the detector kernel (3 = both; the stage-timing entry point runs them apart).
int launch_sync_stream(ofdm_handle* h, const float2* y, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, int parts,
                       cudaStream_t st) {
    // a step is 32*K samples: N/2 for N <= 512, else 256 with m = N/512 sub-steps per N/2-wide block
    const int K = h->N >= 512 ? 8 : h->N / 64;
    const int m = h->N >= 512 ? h->N / 512 : 1;
    if (!(K == 2 || K == 4 || K == 8) || m > 8 || h->cp > h->N / 2 || ws->n_seg == 0 || !ws->mf) return 1;
    StreamParams p;
    p.y = y; p.soff = ss.off; p.max_frames = io->max_frames; p.n = ss.n_max; p.cp = h->cp; p.tapf = (float)(1.0 / (double)h->cp);
    p.seg_len = ws->seg_len; p.n_seg = ws->n_seg; p.seg_cap = (int)ws->seg_cap;
    p.seg_count = ws->seg_count; p.seg_trig = ws->seg_trig; p.first_nan = ws->first_nan; p.status = io->status;
    if (parts & 2) {                                   // the detector owns first_nan
        stream_init_kernel<<<(ss.S + 255) / 256, 256, 0, st>>>(p.first_nan, ss.S);
        OFDM_LAUNCH_CHECK();
    }
    switch (K) {
        case 2: return launch_split_k<2>(p, ws->mf, 1, ss.S, parts, h->sms, st);
        case 4: return launch_split_k<4>(p, ws->mf, 1, ss.S, parts, h->sms, st);
        default: return launch_split_k<8>(p, ws->mf, m, ss.S, parts, h->sms, st);
    }
}
