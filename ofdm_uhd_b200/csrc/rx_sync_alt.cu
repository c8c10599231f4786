// The two other synchronisers the reference names (ofdm_receiver.py~:89-107): SYNC == "pnac" (upstream
// ofdm_sync_pnac: PN cross-correlation + delayed auto-correlation, Tufvesson et al.) and SYNC == "ml" (upstream
// ofdm_sync_ml: van de Beek's cyclic-prefix correlator gated by a known-symbol cross-correlation).  The reference
// hard-codes SYNC = "pn", so neither branch is ever taken there; they are restated here (oracle: sync_pnac / sync_ml in
// oracle/ofdm_oracle.py) as plain, untuned kernels -- correctness and the same interface, not throughput:
//   xcorr            gr.fir_filter_ccc with the conjugated, reversed known symbol: the overlap-save kernel of the
//                    channel filter with another frequency response (rx_front.cu)
//   alt_point        the sample-wise products in front of the moving sums
//   window_sum       gr.fir_filter_fff / _ccf with w equal taps: float64 prefix sums per tile, rounded once
//   pnac_flag / ml_detect   threshold_ff / peak_detector_fb(0.2, 0.25, 30, 0.0005) -> one flag byte per sample
//   evt_count / scan / write  ordered compaction of the flagged samples into (index, angle) lists
#include "internal.h"
#include "common.cuh"
#include <limits.h>
#include <math.h>
#include <complex>
#include <vector>

// ---- sample-wise products -------------------------------------------------------------------------------------
// MODE 1 (pnac): a[n] = |cc[n]|^2.   MODE 2 (ml): a[n] = |y[n]|^2 + |y[n-N]|^2,  mix[n] = conj(y[n-N]) * y[n].
template <int MODE>
__global__ void __launch_bounds__(256) alt_point_kernel(const float2* __restrict__ in, int64_t n, int N, float* __restrict__ a,
                                                        float2* __restrict__ mix) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 v = in[i];
        if (MODE == 1) {
            a[i] = norm_x(v);
        } else {
            const float2 d = (i >= N) ? in[i - N] : make_float2(0.f, 0.f);
            a[i] = fadd_rn(norm_x(v), norm_x(d));
            mix[i] = cmul_x(make_float2(d.x, -d.y), v);            // multiply_cc(conj(delayed), input)
        }
    }
}

// ---- moving sum of w equal taps -------------------------------------------------------------------------------
// out[n] = float32( sum_{k<w} tap * v[n-k] ), products and sum in float64: a tile of WS_T outputs plus a halo of w inputs,
// inclusive float64 prefix over both (block scan), window = difference of two prefixes.
constexpr int WS_THREADS = 256, WS_PER = 8, WS_T = WS_THREADS * WS_PER;

__device__ __forceinline__ double ws_block_incl_scan(double v, double* s_w) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    double base = 0.0;
    for (int j = 0; j < w; ++j) base += s_w[j];
    __syncthreads();
    return inc + base;
}

template <bool CPLX>
__global__ void __launch_bounds__(WS_THREADS) window_sum_kernel(const float* __restrict__ v, int64_t n, int w, double tap,
                                                                float* __restrict__ out) {
    extern __shared__ double s_pre[];                  // [rounds * WS_T + 1] inclusive prefixes, s_pre[0] = 0
    __shared__ double s_w[WS_THREADS / 32];
    const int64_t t0 = (int64_t)blockIdx.x * WS_T;     // first output of the tile
    const int total = WS_T + w;                        // inputs t0 - w .. t0 + WS_T - 1
    const int rounds = (total + WS_T - 1) / WS_T;
    for (int comp = 0; comp < (CPLX ? 2 : 1); ++comp) {
        double carry = 0.0;
        if (threadIdx.x == 0) s_pre[0] = 0.0;
        for (int r = 0; r < rounds; ++r) {
            // each thread: WS_PER consecutive inputs
            const int e0 = r * WS_T + threadIdx.x * WS_PER;
            double loc[WS_PER];
            double run = 0.0;
#pragma unroll
            for (int i = 0; i < WS_PER; ++i) {
                const int e = e0 + i;
                const int64_t g = t0 - w + e;
                double x = 0.0;
                if (e < total && g >= 0 && g < n) x = (double)(CPLX ? v[2 * g + comp] : v[g]) * tap;
                run += x;
                loc[i] = run;
            }
            const double incl = ws_block_incl_scan(run, s_w);
            const double base = carry + incl - run;
#pragma unroll
            for (int i = 0; i < WS_PER; ++i)
                if (e0 + i < total) s_pre[e0 + i + 1] = base + loc[i];
            __syncthreads();
            carry = s_pre[min((r + 1) * WS_T, total)];
            __syncthreads();
        }
        for (int o = threadIdx.x; o < WS_T; o += WS_THREADS) {
            const int64_t g = t0 + o;
            if (g < n) {
                const float r = (float)(s_pre[o + w + 1] - s_pre[o + 1]);      // inputs (g - w, g]
                if (CPLX) out[2 * g + comp] = r; else out[g] = r;
            }
        }
        __syncthreads();
    }
}

static int launch_window_sum(const float* v, int64_t n, int w, double tap, float* out, bool cplx, int device, cudaStream_t st) {
    const int total = WS_T + w;
    const size_t smem = sizeof(double) * (size_t)(((total + WS_T - 1) / WS_T) * WS_T + 2);
    const unsigned grid = (unsigned)((n + WS_T - 1) / WS_T);
    if (cplx) {
        OFDM_SET_MAX_SMEM((window_sum_kernel<true>), smem, device);
        window_sum_kernel<true><<<grid, WS_THREADS, smem, st>>>(v, n, w, tap, out);
    } else {
        OFDM_SET_MAX_SMEM((window_sum_kernel<false>), smem, device);
        window_sum_kernel<false><<<grid, WS_THREADS, smem, st>>>(v, n, w, tap, out);
    }
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// ---- pnac: threshold_ff(0, 0, 0) of |corr|^2 - power ------------------------------------------------------------
__device__ __forceinline__ float pnac_compare(const float2* cc, const float* power, int64_t i, int half, float2* corr_out) {
    const float2 v = cc[i];
    const float2 d = (i >= half) ? cc[i - half] : make_float2(0.f, 0.f);
    const float2 corr = cmul_x(v, make_float2(d.x, -d.y));          // multiply_cc(cc, conj(delayed cc))
    if (corr_out) *corr_out = corr;
    return fsub_rn(norm_x(corr), power[i]);
}

__global__ void __launch_bounds__(256) pnac_flag_kernel(const float2* __restrict__ cc, const float* __restrict__ power, int64_t n,
                                                        int half, uint8_t* __restrict__ flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float c = pnac_compare(cc, power, i, half, nullptr);
        int64_t j = i;
        while (!(c > 0.f) && !(c < 0.f) && j > 0) {                  // exactly on the threshold (or NaN): the previous output holds
            --j;
            c = pnac_compare(cc, power, j, half, nullptr);
        }
        flag[i] = (c > 0.f) ? 1 : 0;
    }
}

// ---- ml: peak_detector_fb(0.2, 0.25, 30, 0.0005) of theta = |ms2| - energy ---------------------------------------
// One thread per segment, sequential like the block itself; every sample updates the IIR average exactly once, so a
// segment that starts ML_WARM samples early has the true average (0.9995^65536 ~ 6e-15) when it reaches its own
// samples, and it owns the runs that START in it (it walks on past its end until an open run closes).
constexpr int ML_SEG = 16384, ML_WARM = 65536;

__global__ void __launch_bounds__(64) ml_detect_kernel(const float2* __restrict__ ms2, const float* __restrict__ energy, int64_t n,
                                                       uint8_t* __restrict__ flag) {
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s0 = seg * ML_SEG;
    if (s0 >= n) return;
    const int64_t s1 = (s0 + ML_SEG < n) ? s0 + ML_SEG : n;
    int64_t i = s0 - ML_WARM;
    if (i < 0) i = 0;
    const double a1 = (double)0.0005f, a2 = 1.0 - a1;
    double avg = 0.0;
    int state = 0;
    float peak = -INFINITY;
    int64_t ind = 0, run_start = 0;
    while (i < n) {
        if (i >= s1 && state == 0) break;
        const float2 m = ms2[i];
        const float x = fsub_rn(__fsqrt_rn(norm_x(m)), energy[i]);
        if (state == 0) {
            if (x > fmul_rn((float)avg, 0.2f)) { state = 1; run_start = i; continue; }      // no sample consumed
            avg = a1 * (double)x + a2 * avg;
            ++i;
        } else if (x > peak) {
            peak = x; ind = i;
            avg = a1 * (double)x + a2 * avg;
            ++i;
        } else if (x > fmul_rn((float)avg, 0.25f)) {
            avg = a1 * (double)x + a2 * avg;
            ++i;
        } else {
            if (run_start >= s0 && run_start < s1) flag[ind] = 1;
            state = 0;
            peak = -INFINITY;
        }
    }
}

// ---- ordered compaction of the flagged samples ------------------------------------------------------------------
__global__ void __launch_bounds__(1024) evt_count_kernel(const uint8_t* __restrict__ flag, int64_t n, int32_t* __restrict__ blk_cnt) {
    __shared__ int s_c[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int c = (i < n && flag[i]) ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        c = s_c[threadIdx.x];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (threadIdx.x == 0) blk_cnt[blockIdx.x] = c;
    }
}

// exclusive scan of the block counts in place; out_n = min(total, cap), overflow bit in status
__global__ void __launch_bounds__(1024) evt_scan_kernel(int32_t* __restrict__ blk_cnt, int64_t nblk, int32_t* __restrict__ out_n, int cap,
                                                        uint32_t* __restrict__ status) {
    __shared__ int s_w[33];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblk; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int v = i < nblk ? blk_cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        if (w == 0) {
            const int t = s_w[lane];
            int ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, ti, d);
                if (lane >= d) ti += o;
            }
            s_w[lane] = ti - t;
            if (lane == 31) s_w[32] = ti;
        }
        __syncthreads();
        if (i < nblk) blk_cnt[i] = s_carry + s_w[w] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int total = s_carry;
        if (total > cap) { total = cap; atomicOr(status, OFDM_ST_TRIG_OVERFLOW); }
        *out_n = total;
    }
}

// MODE 1 (pnac): angle = arg(cc[i] conj(cc[i - N/2])).  MODE 2 (ml): angle = arg(ms2[i]).
template <int MODE>
__global__ void __launch_bounds__(1024) evt_write_kernel(const uint8_t* __restrict__ flag, int64_t n, const int32_t* __restrict__ blk_off,
                                                         int cap, const float2* __restrict__ c, int half,
                                                         int64_t* __restrict__ idx_out, float* __restrict__ ang_out) {
    __shared__ int s_w[33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int f = (i < n && flag[i]) ? 1 : 0;
    int inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int t = s_w[lane];
        int ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += o;
        }
        s_w[lane] = ti - t;
    }
    __syncthreads();
    if (!f) return;
    const int pos = blk_off[blockIdx.x] + s_w[w] + inc - 1;
    if (pos >= cap) return;
    float2 z;
    if (MODE == 1) {
        const float2 v = c[i];
        const float2 d = (i >= half) ? c[i - half] : make_float2(0.f, 0.f);
        z = cmul_x(v, make_float2(d.x, -d.y));
    } else {
        z = c[i];
    }
    idx_out[pos] = i;
    ang_out[pos] = (float)atan2((double)z.y, (double)z.x);
}

// ml: the timing flags = detector flags gated by the known-symbol correlation
__global__ void __launch_bounds__(256) ml_gate_kernel(const uint8_t* __restrict__ flag, const float2* __restrict__ kc,
                                                      const float* __restrict__ energy, int64_t n, uint8_t* __restrict__ timing) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t t = 0;
        if (flag[i]) t = fdiv_rn(norm_x(kc[i]), energy[i]) > 0.1f ? 1 : 0;
        timing[i] = t;
    }
}

// ---- the known symbol's correlator response (host, once per handle and tap count) ---------------------------------
static int make_xcorr_response(ofdm_handle* h, int ntaps, float2** d_out, int* nos_out) {
    // ks0time = ifft(ifftshift(padded known symbol)) (ofdm_receiver.py~:80-87), complex64; taps = conj(ks0time[:ntaps]) reversed
    const int N = h->N;
    std::vector<float> ks(h->occ);
    if (cudaMemcpy(ks.data(), h->d_ks, sizeof(float) * h->occ, cudaMemcpyDeviceToHost) != cudaSuccess) return OFDM_E_CUDA;
    std::vector<std::complex<float>> kst(N);
    for (int nn = 0; nn < N; ++nn) {
        std::complex<double> acc = 0;
        for (int i = 0; i < h->occ; ++i) {
            if (ks[i] == 0.f) continue;
            const long k = ((long)(h->zl + i) + N / 2) % N;
            acc += (double)ks[i] * std::polar(1.0, 2 * M_PI * (double)(k * nn % N) / N);
        }
        kst[nn] = std::complex<float>((float)(acc.real() / N), (float)(acc.imag() / N));
    }
    int NOS = 2048;
    while (NOS < 2 * ntaps) NOS *= 2;
    if (NOS > 4096) { ofdm_set_error("sync_alt: a %d-tap known-symbol correlator does not fit the 4096-point overlap-save", ntaps); return OFDM_E_INVAL; }
    std::vector<std::complex<double>> taps(ntaps);
    for (int k = 0; k < ntaps; ++k) taps[k] = std::conj(std::complex<double>(kst[ntaps - 1 - k]));
    std::vector<float2> H(NOS);
    for (int k = 0; k < NOS; ++k) {
        std::complex<double> acc = 0;
        for (int t = 0; t < ntaps; ++t) acc += taps[t] * std::polar(1.0, -2 * M_PI * (double)((long)k * t % NOS) / NOS);
        H[k] = make_float2((float)(acc.real() / NOS), (float)(acc.imag() / NOS));
    }
    if (cudaMalloc((void**)d_out, sizeof(float2) * NOS) != cudaSuccess) return OFDM_E_CUDA;
    if (cudaMemcpy(*d_out, H.data(), sizeof(float2) * NOS, cudaMemcpyHostToDevice) != cudaSuccess) return OFDM_E_CUDA;
    *nos_out = NOS;
    return OFDM_OK;
}

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

size_t sync_alt_scratch_bytes(int64_t n) {
    if (n < 0) n = 0;
    return 3 * up256((size_t)n * 8) + 2 * up256((size_t)n * 4) + 2 * up256((size_t)n) + up256(((size_t)(n + 1023) / 1024) * 4) + 256;
}

static int compact_events(int mode, const uint8_t* flag, int64_t n, int32_t* blk, int64_t nblk, int cap, const float2* c, int half,
                          int32_t* n_out, int64_t* idx_out, float* ang_out, uint32_t* status, cudaStream_t st) {
    evt_count_kernel<<<(unsigned)nblk, 1024, 0, st>>>(flag, n, blk);
    OFDM_LAUNCH_CHECK();
    evt_scan_kernel<<<1, 1024, 0, st>>>(blk, nblk, n_out, cap, status);
    OFDM_LAUNCH_CHECK();
    if (mode == 1) evt_write_kernel<1><<<(unsigned)nblk, 1024, 0, st>>>(flag, n, blk, cap, c, half, idx_out, ang_out);
    else evt_write_kernel<2><<<(unsigned)nblk, 1024, 0, st>>>(flag, n, blk, cap, c, half, idx_out, ang_out);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// mode 1: "pnac", 2: "ml".  Fills io->trig_idx / trig_ang / n_trig with the TIMING triggers; for ml also the NCO event
// list (every detector peak) in ws->nco_idx / nco_ang / n_nco (n_nco = -1 behind pnac: the NCO follows the triggers).
int launch_sync_alt(ofdm_handle* h, const float2* y, int64_t n, int mode, float snr_db, ofdm_rx_io* io, RxWorkspace* ws,
                    void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (mode != 1 && mode != 2) { ofdm_set_error("sync_alt: mode %d (1 = pnac, 2 = ml)", mode); return OFDM_E_INVAL; }
    if (n <= 0) { OFDM_CUDA_CHECK(cudaMemsetAsync(io->n_trig, 0, sizeof(int32_t), st)); return OFDM_OK; }
    if (scratch_bytes < sync_alt_scratch_bytes(n) || !scratch) { ofdm_set_error("sync_alt: scratch too small"); return OFDM_E_NOMEM; }
    const int N = h->N, cp = h->cp;
    const int ntaps = mode == 1 ? N / 2 : N;
    float2** Hslot = mode == 1 ? &h->d_Hks_half : &h->d_Hks_full;
    int* nos = mode == 1 ? &h->nos_ks_half : &h->nos_ks_full;
    if (!*Hslot) {
        int rc = make_xcorr_response(h, ntaps, Hslot, nos);
        if (rc) return rc;
    }
    OFDM_CUDA_CHECK(cudaMemsetAsync(io->status, 0, sizeof(uint32_t), st));
    OFDM_CUDA_CHECK(cudaMemsetAsync(ws->nco_init, 0, sizeof(double), st));
    { int rc0 = launch_nco_mode(h, ws, 1, mode == 2 ? -1.0 : -2.0, st); if (rc0) return rc0; }    // ofdm_receiver.py~:91,103
    OFDM_CUDA_CHECK(cudaMemsetAsync(ws->first_nan, 0x7F, sizeof(int64_t), st));      // no NaN cut behind these synchronisers
    char* b = (char*)scratch;
    float2* cc = (float2*)b; b += up256((size_t)n * 8);       // cross-correlation with the known symbol
    float2* mix = (float2*)b; b += up256((size_t)n * 8);      // ml: conj(y[n-N]) y[n]
    float2* ms2 = (float2*)b; b += up256((size_t)n * 8);      // ml: its cp-wide sum
    float* a = (float*)b; b += up256((size_t)n * 4);          // pnac: |cc|^2 ; ml: |y|^2 + |y[n-N]|^2
    float* sum = (float*)b; b += up256((size_t)n * 4);        // pnac: power ; ml: energy
    uint8_t* flag = (uint8_t*)b; b += up256((size_t)n);
    uint8_t* timing = (uint8_t*)b; b += up256((size_t)n);
    int32_t* blk = (int32_t*)b;
    const int64_t nblk = (n + 1023) / 1024;
    int rc = launch_xcorr(h, y, n, *Hslot, *nos, ntaps, cc, st);
    if (rc) return rc;
    int64_t pb = (n + 255) / 256;
    if (pb > (int64_t)h->sms * 16) pb = (int64_t)h->sms * 16;
    OFDM_CUDA_CHECK(cudaMemsetAsync(flag, 0, (size_t)n, st));
    if (mode == 1) {
        alt_point_kernel<1><<<(int)pb, 256, 0, st>>>(cc, n, N, a, nullptr);
        OFDM_LAUNCH_CHECK();
        if ((rc = launch_window_sum(a, n, N, 1.0, sum, false, h->device, st))) return rc;
        pnac_flag_kernel<<<(int)pb, 256, 0, st>>>(cc, sum, n, N / 2, flag);
        OFDM_LAUNCH_CHECK();
        if ((rc = compact_events(1, flag, n, blk, nblk, io->max_frames, cc, N / 2, io->n_trig, io->trig_idx, io->trig_ang, io->status, st))) return rc;
        return OFDM_OK;                                     // n_nco stays -1: the NCO follows the triggers
    }
    const double snr = pow(10.0, (double)snr_db / 10.0);
    const float tap = (float)((snr / (snr + 1.0)) / 2.0);                            // rho / 2 as the float32 FIR tap
    alt_point_kernel<2><<<(int)pb, 256, 0, st>>>(y, n, N, a, mix);
    OFDM_LAUNCH_CHECK();
    if ((rc = launch_window_sum(a, n, cp, (double)tap, sum, false, h->device, st))) return rc;
    if ((rc = launch_window_sum((const float*)mix, n, cp, 1.0, (float*)ms2, true, h->device, st))) return rc;
    const int64_t nseg = (n + ML_SEG - 1) / ML_SEG;
    ml_detect_kernel<<<(unsigned)((nseg + 63) / 64), 64, 0, st>>>(ms2, sum, n, flag);
    OFDM_LAUNCH_CHECK();
    // every detector peak steps the NCO (sample_and_hold of the cp-correlation angle) ...
    if ((rc = compact_events(2, flag, n, blk, nblk, io->max_frames, ms2, 0, ws->n_nco, ws->nco_idx, ws->nco_ang, io->status, st))) return rc;
    // ... the timing output keeps the peaks on a known symbol
    ml_gate_kernel<<<(int)pb, 256, 0, st>>>(flag, cc, sum, n, timing);
    OFDM_LAUNCH_CHECK();
    return compact_events(2, timing, n, blk, nblk, io->max_frames, ms2, 0, io->n_trig, io->trig_idx, io->trig_ang, io->status, st);
}
