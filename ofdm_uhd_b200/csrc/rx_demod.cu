// Receive back end: derotation + forward FFT + ofdm_frame_acquisition (one CTA per frame), ofdm_frame_sink (one warp
// per sink session), then the sink's liveness walk, dewhitening and CRC, and the dense hand-over to the host.
// Reference wiring: ofdm_receiver.py~:124-129 (sigmix, fft_demod, ofdm_frame_acq), ofdm.py:238-247
// (ofdm_frame_sink), ofdm.py:300-305 + ofdm_packet_utils.py:169-191 (unmake_packet).
#include "internal.h"
#include "fft.cuh"
#include <limits.h>
#include <stdlib.h>

// The receive back end is two kernels:
//   acq_kernel<N>   multiply_cc (NCO derotation) + fft_vcc (forward, shifted) + ofdm_frame_acquisition for every vector
//   / acq_warp_kernel<N>  the sampler emits, one CTA per frame (acq_warp_kernel, N = 512 / 1024 without taps: half a warp
//                   / a warp per frame on the warp plans): the equalised occ-wide vectors land in the workspace (they are
//                   a function of the vector stream alone: the acquisition block re-estimates at EVERY flagged vector,
//                   whatever state the sink is in);
//   sink_kernel     ofdm_frame_sink, one WARP per speculative session (one per frame): walks the equalised vectors from
//                   its frame's preamble on -- across later frames if the header it finds says so -- with the PLL / DFE
//                   slicer, the byte packing and the header parse done warp-synchronously (no block barriers, the
//                   serial carrier-loop update computed redundantly by all lanes).
// (One fused kernel did both until round 2: 8 block barriers and a one-thread PLL section per vector on 64-thread CTAs,
// 20 % of its stall samples waiting on instruction fetch.)
struct AcqParams {
    const float2* y;
    const int64_t* soff;           // stream offsets (nullptr: one stream of n samples); tables below are per stream
    int64_t n;
    const int64_t* trig_idx;
    const double* phi0;
    const double* step;
    const double* nco_init;
    const int32_t* n_trig;
    const int32_t* first_ok;
    const int32_t* n_frames;
    const int32_t* frame_ndata;
    const int64_t* vbase;
    const int32_t* n_nco;          // the NCO's own event list (ofdm_sync_ml); *n_nco < 0: it follows the triggers
    const int64_t* nco_idx;
    const float2* tw;
    const float* ks;
    const float* kd;
    int occ, cp, zl, L, max_frames;
    float2* eq;                    // [S][eq_stride][occ] equalised vectors (ofdm_frame_acquisition output)
    int64_t eq_stride;             // vectors per stream in eq
    float2* eq_tap;                // optional copy for the parity tests / --log (io->eq_syms), single stream
    float2* fft_tap;               // optional: shifted spectra [max_vectors][N] (ofdm_receiver-fft_out_c.dat)
    float2* samp_tap;              // optional: derotated sampler vectors [max_vectors][N] (sampler_c / sigmix)
    int64_t max_vectors;
};

struct SinkParams {
    const float2* eq;
    int64_t eq_stride;
    int per_stream;                // 1: tables are per stream (blockIdx.y)
    const int32_t* n_frames;
    const int64_t* vbase;
    const float2* cst;
    const int16_t* sinkmap;
    int occ, ncar, nbits, M, max_frames, pkt_stride;
    int grid_L;
    float grid_x0, grid_y0, grid_inv_dx, grid_inv_dy;
    const uint8_t* grid;
    uint8_t* frame_status;
    int32_t* pkt_len;
    int32_t* sess_nvec;
    uint8_t* pkt_bytes;
    uint8_t* sym_idx;              // optional taps (single stream), indexed by vector
    float2* derot_syms;
    int64_t max_vectors;
};

// multiply_cc(chan_filt, frequency_modulator_fc(-2/N)(sample_and_hold(angle))) evaluated at the
// samples of one sampler vector: phi[s] = phi0[k] + step[k]*(s - t_k + 1) for t_k <= s < t_{k+1}  (A.8).
// Inside one trigger segment the phasor of point tid + d is  e^{j phi(st+tid)} * e^{j step d}: one sincos per
// thread and vector plus a per-segment table of the E compile-time offsets d; samples at or after the next
// trigger (the last sample of a preamble vector) take the exact per-sample path.
__device__ __forceinline__ float2 phasor_f64(double ph) {
    double r = ph * 0.15915494309189533577;
    r -= rint(r);
    float sn, cs;
    sincospif(2.0f * (float)r, &sn, &cs);
    return make_float2(cs, sn);
}

// a sample at or after the next trigger (the last sample of a preamble vector): exact per-sample NCO phase.
// Out of line: it is inlined E times otherwise and the per-vector loop must stay small.
__device__ __noinline__ float2 derot_slow(float2 v, int64_t s, const int64_t* trig, const double* phi0, const double* step,
                                          int K, int kk) {
    while (kk + 1 < K && LDG(trig + kk + 1) <= s) ++kk;
    const double ph = LDG(phi0 + kk) + LDG(step + kk) * (double)(s - LDG(trig + kk) + 1);
    return cmul_x(v, phasor_f64(ph));
}

template <bool SMEM>
struct DemodLoad {
    const float2* y;               // the vector's samples: staged in shared memory (SMEM), else the stream at `st`
    int64_t st;
    int lim;                       // points idx < lim lie before the next trigger (32-bit compare per point)
    const int64_t* trig;
    const double* phi0;
    const double* step;
    int K, kk0;
    float2 ph0;
    const float2* Wt;
    __device__ __forceinline__ float2 operator()(int idx, int slot) const {
        const float2 v = SMEM ? y[idx] : LDG(y + idx);
        if (idx < lim) return cmul_x(v, cmul(ph0, Wt[slot]));
        return derot_slow(v, st + idx, trig, phi0, step, K, kk0);
    }
};

template <int N>
struct ShiftStore {                       // fft_vcc(..., shift=True): bin idx lands at (idx + N/2) mod N
    float2* S;
    __device__ __forceinline__ void operator()(int idx, float2 v, int) const { S[(idx + N / 2) & (N - 1)] = v; }
};

__device__ __forceinline__ float2 expj_f32(float ph) {
    double s, c;
    sincos((double)ph, &s, &c);
    return make_float2((float)c, (float)s);
}

// coarse_freq_comp(delta, count) = expj(float(-2*pi*delta*cp) / N * count)   (A.10)
__device__ __forceinline__ float2 coarse_comp(int delta, int cp, int N, int cnt) {
    if (delta == 0) return make_float2(1.0f, -0.0f);     // ph = -0.0: (cos, sin) without the float64 sincos
    const float a = (float)(-2.0 * 3.14159265358979323846 * (double)delta * (double)cp);
    const float ph = fmul_rn(fdiv_rn(a, (float)N), (float)cnt);
    return expj_f32(ph);
}

// LSB-first packing of the slicer decisions of one vector into bytes q in [B0, B1).  One generic routine (nbits is
// 1, 2, 3, 4, 6 or 8: a shift, then an exact multiply-shift division by 3 where needed) instead of eight
// constant-divisor instantiations: the per-vector loop has to stay inside the instruction cache.
struct BitDiv {
    int sh, by3, nbits;
    __device__ __forceinline__ explicit BitDiv(int nb) : nbits(nb) {
        sh = __ffs(nb) - 1;
        by3 = (nb >> sh) == 3;
    }
    // rel / nbits for 0 <= rel < 98304
    __device__ __forceinline__ int div(int rel) const {
        const int pre = rel >> sh;
        return by3 ? (int)(((unsigned)pre * 43691u) >> 17) : pre;
    }
};

// NB: bits per symbol when known at compile time (1 or 2: the per-byte loop unrolls), 0 = bd.nbits at run time
template <int NB = 0>
__device__ __forceinline__ void pack_bytes(const uint8_t* sym, uint8_t* vb, int B0, int B1, int bit_base, unsigned carry,
                                           int tid, int nthreads, const BitDiv bd) {
    if (NB > 0 || !bd.by3) {
        // nbits divides 8 (bpsk, qpsk, qam16, qam256): a symbol never straddles a byte, and bit_base is a multiple of
        // nbits, so a byte is 8/nbits whole symbols (or the previous vector's leftover bits in their place)
        const int nb = NB > 0 ? NB : bd.nbits;
        const int lsh = NB > 0 ? NB - 1 : bd.sh;                   // log2(nb)
        const unsigned smask = (1u << nb) - 1u;
        for (int q = B0 + tid; q < B1; q += nthreads) {
            unsigned byte = 0;
            int rel = 8 * q - bit_base;
#pragma unroll
            for (int sh = 0; sh < 8; sh += nb, rel += nb) {
                const unsigned v = (rel < 0) ? ((carry >> sh) & smask) : (unsigned)sym[rel >> lsh];
                byte |= v << sh;
            }
            vb[q - B0] = (uint8_t)byte;
        }
        return;
    }
    for (int q = B0 + tid; q < B1; q += nthreads) {
        unsigned byte = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rel = 8 * q + i - bit_base;
            unsigned bit;
            if (rel < 0) bit = (carry >> i) & 1u;
            else {
                const int c = bd.div(rel);
                bit = ((unsigned)sym[c] >> (rel - c * bd.nbits)) & 1u;
            }
            byte |= bit << i;
        }
        vb[q - B0] = (uint8_t)byte;
    }
}

// ofdm_frame_sink's slicer: the first minimum of |r - const[k]|^2 over k in index order, every distance in
// individually rounded float32 arithmetic (A.11).
//   Brute force over all M points -- or, for a square-grid constellation (grid_L > 0, |coordinates| <= 1) and a
// received point with |re|, |im| <= 16, over the 3 x 3 cells around the nearest cell: a point outside that
// neighbourhood is at least 1.5 cells away along one axis where some neighbourhood point is at most 0.5, so its
// exact squared distance is larger by >= 2 cell^2 (>= 0.035 for qam256) while the float32 rounding of either
// distance is below 2e-4 in that range; the minimum over the neighbourhood with ties broken towards the smaller
// index is therefore exactly what the scan over all M points returns.
struct Slicer {
    const float2* cst;
    const uint8_t* grid;
    int M, L;
    float x0, y0, inv_dx, inv_dy;
    __device__ __forceinline__ int operator()(const float2 r) const {
        if (L > 0 && fabsf(r.x) <= 16.f && fabsf(r.y) <= 16.f) {
            int ix = __float2int_rn((r.x - x0) * inv_dx), iy = __float2int_rn((r.y - y0) * inv_dy);
            ix = min(max(ix, 0), L - 1);
            iy = min(max(iy, 0), L - 1);
            float best = INFINITY;
            int b = INT_MAX;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = iy + dy;
                if (yy < 0 || yy >= L) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = ix + dx;
                    if (xx < 0 || xx >= L) continue;
                    const int k = grid[yy * L + xx];
                    const float2 ck = cst[k];
                    const float dd = norm_x(csub_x(r, ck));
                    if (dd < best || (dd == best && k < b)) { best = dd; b = k; }
                }
            }
            return b;
        }
        int b = 0;
        float best;
        {
            best = norm_x(csub_x(r, cst[0]));
        }
        for (int k = 1; k < M; ++k) {
            const float dd = norm_x(csub_x(r, cst[k]));
            if (dd < best) { best = dd; b = k; }
        }
        return b;
    }
};
// ---------------------------------------------------------------------------------------------
// acq_kernel: sigmix + fft_demod + ofdm_frame_acq (ofdm_receiver.py~:124-129) for all vectors of one frame per CTA.
// ---------------------------------------------------------------------------------------------
template <int N, bool TAPS>
__global__ void __launch_bounds__((N / FftPlan<N>::E) < 64 ? 64 : (N / FftPlan<N>::E),
                                  TAPS ? 1 : (FftPlan<N>::E == 8 ? 1024 : 512) / ((N / FftPlan<N>::E) < 64 ? 64 : (N / FftPlan<N>::E))) acq_kernel(const AcqParams p) {
    // The tables of this CTA's stream (blockIdx.y): CTA-uniform offsets on top of the kernel parameters, held in a
    // view of their own (v) -- a locally modified COPY of the parameter struct is not safe: nvcc kept reading the
    // unshifted y pointer from the parameter space in the instantiations that load samples straight from global memory.
    struct {
        const float2* y;
        const int64_t* trig_idx;
        const double* phi0;
        const double* step;
        const double* nco_init;
        const int32_t* n_trig;
        const int32_t* first_ok;
        const int32_t* n_frames;
        const int32_t* frame_ndata;
        const int64_t* vbase;
        const int64_t* nidx;       // sample index of every NCO event (the trigger list unless ofdm_sync_ml supplied its own)
        float2* eq;
    } v;
    {
        const int64_t sidx = p.soff ? (int64_t)blockIdx.y : 0;
        const int64_t mf = p.max_frames;
        const int64_t a = p.soff ? p.soff[sidx] : 0;
        v.y = p.y + a;
        v.trig_idx = p.trig_idx + sidx * mf; v.phi0 = p.phi0 + sidx * mf; v.step = p.step + sidx * mf;
        v.nco_init = p.nco_init + sidx; v.n_trig = p.n_trig + sidx; v.first_ok = p.first_ok + sidx;
        v.n_frames = p.n_frames + sidx; v.frame_ndata = p.frame_ndata + sidx * mf;
        v.vbase = p.vbase + sidx * (mf + 1);
        v.eq = p.eq + sidx * p.eq_stride * p.occ;
        v.nidx = (p.n_nco[sidx] >= 0) ? p.nco_idx + sidx * mf : v.trig_idx;
    }
    using P = FftPlan<N>;
    constexpr int E = P::E;
    constexpr int T = N / E;
    constexpr int BT = T < 64 ? 64 : T;
    constexpr int NW = BT / 32;
    constexpr int SB = fft_smem_elems<N>();
    constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2];
    extern __shared__ double smem_d[];
    double* red = smem_d;                                   // [8*NW]
    // N = 4096: a single FFT buffer (in-place passes with a barrier between their loads and stores, no prefetch), so
    // that several CTAs fit an SM
    constexpr bool ONEBUF = (N == 4096);
    float2* bufA = (float2*)(red + 8 * NW);
    float2* bufB = ONEBUF ? bufA : bufA + SB;
    float2* H = bufB + SB;                                  // [occ]
    __shared__ int s_delta;
    __shared__ float2 s_cc[2], s_W[E];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int F = *v.n_frames;
    int K = *v.n_trig;                                      // NCO events: the triggers, or ofdm_sync_ml's own list
    if (K > p.max_frames) K = p.max_frames;
    const bool own_nco = v.nidx != v.trig_idx;
    if (own_nco) { K = p.n_nco[p.soff ? blockIdx.y : 0]; if (K > p.max_frames) K = p.max_frames; }
    const int first_ok = *v.first_ok;
    const int occ = p.occ, zl = p.zl, L = p.L;
    float2* S = (P::NP == 2) ? bufB : bufA;                 // shifted spectrum of the current vector
    // Three-pass plans leave bufB idle while a vector is equalised: the next vector's samples are copied into it
    // asynchronously meanwhile, so the first FFT pass never waits on HBM.
    constexpr bool PF = (P::NP == 3) && !ONEBUF;
    auto prefetch = [&](int64_t st2) {
#pragma unroll
        for (int i = 0; i < N / BT; ++i) {
            const int idx = tid + i * BT;
            cp_async8(bufB + idx, v.y + st2 + idx, 8);
        }
        cp_async_commit();
    };

    for (int f = blockIdx.x; f < F; f += gridDim.x) {
        const int kg = first_ok + f;
        const int64_t t = v.trig_idx[kg];
        const int nd = v.frame_ndata[f];
        const int64_t vb0 = v.vbase[f];
        int cnt = 1, delta = 0;
        int kk_w = INT_MIN;                                 // trigger segment the table s_W belongs to
        if (PF) cp_async_wait_all();                        // a copy left in flight by the previous frame
        __syncthreads();
        if (PF) prefetch(t - N + 1);
        for (int m = 0; m <= nd; ++m) {
            const int64_t st = t - N + 1 + (int64_t)m * L;
            const bool flag = (m == 0);
            const int64_t vglob = vb0 + m;
            const int par = m & 1;
            // ---- sigmix + fft_demod ----
            int kk = kg;
            if (own_nco) {                                  // last NCO event at or before the vector's first sample
                int lo = 0, hi = K;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (v.nidx[mid] <= st) lo = mid + 1; else hi = mid;
                }
                kk = lo - 1;
            } else if (flag) {
                kk = kg - 1;
                while (kk >= 0 && v.nidx[kk] > st) --kk;
            }
            const int64_t t_next = (kk + 1 < K) ? v.nidx[kk + 1] : LLONG_MAX;
            double stp, ph_base;
            if (kk >= 0) {
                stp = v.step[kk];
                ph_base = v.phi0[kk] + stp * (double)(st + tid - v.nidx[kk] + 1);
            } else {                                        // before the first trigger (0 behind ofdm_sync_pn)
                stp = *v.nco_init;
                ph_base = stp * (double)(st + tid + 1);
            }
            if (kk != kk_w) {                               // block-uniform
                __syncthreads();
                if (tid < E) {
                    const int d = (tid / R0) * T + (tid % R0) * (N / R0);     // slot q*R0 + r  <->  offset q*T + r*N/R0
                    s_W[tid] = phasor_f64(stp * (double)d);
                }
                kk_w = kk;
                __syncthreads();
            }
            const int64_t t_lim = (kk < 0) ? ((K > 0) ? v.nidx[0] : LLONG_MAX) : t_next;
            const int lim = (t_lim - st >= (int64_t)N) ? N : (t_lim > st ? (int)(t_lim - st) : 0);
            DemodLoad<PF> ld{PF ? bufB : v.y + st, st, lim, v.nidx, v.phi0, v.step, K, kk < 0 ? 0 : kk, phasor_f64(ph_base), s_W};
            if (PF) cp_async_wait_all();
            if (P::NP == 3) __syncthreads();                // S (= bufA) of the previous vector has been read by everyone
            if (TAPS && p.samp_tap && vglob < p.max_vectors && tid < T) {
                // the sampler's output vector (ofdm_receiver-sampler_c.dat; sigmix sits in front of the sampler, so it
                // is derotated already): the very values the first FFT pass reads, through the same functor
#pragma unroll
                for (int q = 0; q < E / R0; ++q)
#pragma unroll
                    for (int r = 0; r < R0; ++r) {
                        const int idx = tid + q * T + r * (N / R0);
                        p.samp_tap[vglob * N + idx] = ld(idx, q * R0 + r);
                    }
            }
            if (tid < T) fft_pass<N, R0, 1, -1>(tid, p.tw, ld, SmemOut{bufA});
            __syncthreads();
            if constexpr (P::NP == 2) {
                if (tid < T) fft_pass<N, R1, R0, -1>(tid, p.tw, SmemIn{bufA}, ShiftStore<N>{S});
            } else {
                if (tid < T) fft_pass<N, R1, R0, -1, SmemIn, SmemOut, ONEBUF>(tid, p.tw, SmemIn{bufA}, SmemOut{bufB});
                __syncthreads();
                if (tid < T) fft_pass<N, R2, R0 * R1, -1, SmemIn, ShiftStore<N>, ONEBUF>(tid, p.tw, SmemIn{bufB}, ShiftStore<N>{S});
            }
            __syncthreads();
            if (m < nd) {                                   // fetch the vector that follows
                const int64_t st2 = st + L;
                if (PF) {
                    prefetch(st2);                          // bufB is free: asynchronous copy into it
                } else {
                    // two-pass plans and the single-buffer layout have no landing zone in shared memory: at least pull
                    // the vector's lines into L2 while this one is equalised (one 128-byte line per thread)
                    for (int i = tid * 16; i < N; i += BT * 16)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(v.y + st2 + i));
                }
            }
            if (TAPS && p.fft_tap && vglob < p.max_vectors)
                for (int i = tid; i < N; i += BT) p.fft_tap[vglob * N + i] = S[i];
            // ---- ofdm_frame_acquisition: correlate + calculate_equalizer on a flagged vector ----
            if (flag) {
                double acc[2 * OFDM_MAX_SHIFT];
#pragma unroll
                for (int s = 0; s < 2 * OFDM_MAX_SHIFT; ++s) acc[s] = 0.0;
                for (int j = 2 * tid; j < occ - 2; j += 2 * BT) {
                    const float kdj = LDG(p.kd + j);
#pragma unroll
                    for (int s = 0; s < 2 * OFDM_MAX_SHIFT; ++s) {
                        const int pi = zl - OFDM_MAX_SHIFT + s + j;
                        const float2 a = S[pi], b = S[pi + 2];
                        acc[s] += (double)kdj * (double)norm_x(csub_x(a, b));
                    }
                }
#pragma unroll
                for (int s = 0; s < 2 * OFDM_MAX_SHIFT; ++s) {
                    double q = acc[s];
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
                    if (lane == 0) red[w * 8 + s] = q;
                }
                __syncthreads();
                if (tid == 0) {
                    float best = 0.f;
                    int index = 0;
                    for (int s = 0; s < 2 * OFDM_MAX_SHIFT; ++s) {
                        double q = 0.0;
                        for (int ww = 0; ww < NW; ++ww) q += red[ww * 8 + s];
                        const float sf = (float)q;
                        if (sf > best) { best = sf; index = zl - OFDM_MAX_SHIFT + s; }
                    }
                    s_delta = index - zl;
                    s_cc[par] = coarse_comp(index - zl, p.cp, N, 1);
                }
                __syncthreads();
                delta = s_delta;
                cnt = 1;
                const float2 c1 = s_cc[par];
                for (int i = 2 * tid; i < occ; i += 2 * BT) {
                    const float2 b = cmul_x(c1, S[i + zl + delta]);
                    H[i] = cdiv_x(make_float2(LDG(p.ks + i), 0.f), b);
                }
                __syncthreads();
                for (int i = 2 * tid + 1; i + 1 < occ; i += 2 * BT) {
                    const float2 a = H[i + 1], b = H[i - 1];
                    H[i] = cscale_x(cadd_x(a, b), 0.5f);
                }
                if (tid == 0 && (occ & 1) == 0) H[occ - 1] = H[occ - 2];
                __syncthreads();
            }
            // ---- one-tap equaliser with the coarse-offset CP phase compensation: comp(delta, cnt) is in s_cc[par]
            // (written by the acquisition above, or while the previous vector was equalised).  No coarse offset:
            // comp = (1, -0) and H * comp is H itself (up to the sign of an exact zero, which no comparison downstream
            // can see), so that multiply is skipped
            const float2 cc = s_cc[par];
            int cnt_next = cnt + 1;
            if (cnt_next == OFDM_ACQ_MAX_SYMBOLS) cnt_next = 1;
            if (tid == BT - 1 && m < nd) s_cc[par ^ 1] = coarse_comp(delta, p.cp, N, cnt_next);
            float2* dst = v.eq + vglob * occ;
            for (int i = tid; i < occ; i += BT) {
                const float2 Hi = H[i];
                const float2 o = cmul_x(delta == 0 ? Hi : cmul_x(Hi, cc), S[i + zl + delta]);
                dst[i] = o;
                if (TAPS && p.eq_tap && vglob < p.max_vectors) p.eq_tap[vglob * occ + i] = o;
            }
            cnt = cnt_next;
        }
    }
    if (PF) cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------
// sink_kernel: ofdm_frame_sink (ofdm.py:238-247), one warp per speculative session.
// Session f starts at frame f's flagged vector (position vbase[f] of the equalised vector stream) as if the sink were
// in SYNC_SEARCH there, and walks on -- over later frames' vectors too: flags are ignored outside SYNC_SEARCH -- until
// its header turns out bad (2 vectors), its packet is complete, or the stream ends.  The liveness walk (launch_finish)
// then decides which sessions the one real sink would have run.
// ---------------------------------------------------------------------------------------------
// TPS = threads per session: 16 / 32 lanes of a warp (32 / TPS sessions per warp, several warps per CTA) for the
// usual layouts -- the per-vector serial section (float64 atan2 + sincos of the PLL, header check, leftover bits) is
// issued once per warp, so sub-warp sessions share it -- or 128 (the whole CTA works on one session, block barriers
// instead of warp barriers) when a vector has thousands of data carriers.
// MC: the constellation size when it is 2 or 4 and known at compile time (the brute-force slicer's scan unrolls: no loop
// counter, compare and branch per point), 0 = p.M at run time.
template <bool TAPS, int TPS, int MC = 0>
__global__ void __launch_bounds__(128) sink_kernel(const SinkParams p) {
    extern __shared__ __align__(16) unsigned char sink_smem[];
    __shared__ double s_red[2][4][2];
    constexpr bool WARP = (TPS <= 32);                                 // a session lives inside one warp
    const int lane = WARP ? (threadIdx.x & (TPS - 1)) : threadIdx.x;   // thread index inside the session
    const int w = WARP ? (threadIdx.x / TPS) : 0, W = WARP ? (blockDim.x / TPS) : 1;
    // the lanes of this session inside its warp: barriers and shuffles name exactly them, so the sessions of a warp may
    // part ways (one header bad, one packet longer) without waiting for each other
    const unsigned gmask = TPS >= 32 ? 0xffffffffu : (((1u << (TPS & 31)) - 1u) << ((threadIdx.x & 31) & ~(TPS - 1)));
    auto sync = [&] { if (WARP) __syncwarp(gmask); else __syncthreads(); };
    const int ncar = p.ncar, nbits = p.nbits, occ = p.occ;
    // shared: constellation [M], grid table [grid_L^2] (CTA-wide), then per warp: dfe [ncar] float2, sym [ncar], vb
    float2* s_cst = (float2*)sink_smem;
    uint8_t* s_grid = (uint8_t*)(s_cst + p.M);
    const int grid_bytes = (p.grid_L * p.grid_L + 15) & ~15;
    const int sym_bytes = (ncar + 15) & ~15, vb_bytes = (ncar * nbits / 8 + 16 + 15) & ~15;
    const size_t per_warp = (size_t)ncar * sizeof(float2) + sym_bytes + vb_bytes;
    unsigned char* mine = (unsigned char*)(s_grid + grid_bytes) + (size_t)w * per_warp;
    float2* dfe = (float2*)mine;
    uint8_t* sym = (uint8_t*)(dfe + ncar);
    uint8_t* vb = sym + sym_bytes;
    for (int i = threadIdx.x; i < p.M; i += blockDim.x) s_cst[i] = p.cst[i];
    for (int i = threadIdx.x; i < p.grid_L * p.grid_L; i += blockDim.x) s_grid[i] = p.grid[i];
    __syncthreads();
    const Slicer slicer{s_cst, s_grid, p.M, p.grid_L, p.grid_x0, p.grid_y0, p.grid_inv_dx, p.grid_inv_dy};
    const int64_t sidx = p.per_stream ? (int64_t)blockIdx.y : 0;
    const int64_t mf = p.max_frames;
    const int F = p.n_frames[sidx];
    const int64_t* vbase = p.vbase + sidx * (mf + 1);
    const float2* eq = p.eq + sidx * p.eq_stride * occ;
    uint8_t* frame_status = p.frame_status + sidx * mf;
    int32_t* pkt_len = p.pkt_len + sidx * mf;
    int32_t* sess_nvec = p.sess_nvec + sidx * mf;
    uint8_t* pkt_bytes = p.pkt_bytes + sidx * mf * (int64_t)p.pkt_stride;
    const int64_t vtot = F > 0 ? vbase[F] : 0;              // vectors in the stream
    const int bits_this = ncar * nbits;
    const BitDiv bd(nbits);

    for (int f = blockIdx.x * W + w; f < F; f += gridDim.x * W) {
        const int64_t v0 = vbase[f];
        const int64_t v_own_end = vbase[f + 1];             // taps: only the vectors of the session's own frame
        // enter_have_sync: the flagged vector itself is not demapped
        for (int c = lane; c < ncar; c += TPS) dfe[c] = make_float2(1.f, 0.f);
        float phase = 0.f, freq = 0.f;
        float2 car = make_float2(1.f, 0.f);                 // expj(0)
        unsigned carry = 0u;
        int bit_base = 0, len = 0, status = 3, nvec = INT_MAX / 2;
        sync();
        for (int vi = 1; v0 + vi < vtot; ++vi) {
            const int64_t vg = v0 + vi;
            const float2* e = eq + vg * occ;
            const bool tap = TAPS && vg < v_own_end && vg < p.max_vectors;
            double er = 0.0, ei = 0.0;
            // two carriers per pass with every load ahead of the first store: two independent slicer chains
            for (int cbase = 0; cbase < ncar; cbase += 2 * TPS) {
                const int c = cbase + lane;
                if (cbase + TPS >= ncar) {
                    // last pass of a vector whose carriers do not fill a pair of rows (session-uniform branch): one
                    // chain, instead of a second one computed and thrown away (198 carriers = 3 pairs + 6)
                    if (c < ncar) {
                        const float2 d0a = dfe[c];
                        const float2 eqa = LDG(e + LDG(p.sinkmap + c));
                        const float2 ra = cmul_x(cmul_x(eqa, car), d0a);
                        int ba;
                        if (MC == 0 && p.grid_L > 0) {
                            ba = slicer(ra);
                        } else {
                            ba = 0;
                            float besta = norm_x(csub_x(ra, s_cst[0]));
#pragma unroll
                            for (int k = 1; k < (MC ? MC : p.M); ++k) {
                                const float dda = norm_x(csub_x(ra, s_cst[k]));
                                if (dda < besta) { besta = dda; ba = k; }
                            }
                        }
                        const float2 cla = s_cst[ba];
                        const float2 ea = cmulc_x(ra, cla);
                        er += (double)ea.x;
                        ei += (double)ea.y;
                        if (norm_x(ra) > 0.001f) {
                            const float2 dq = cscale_x(csub_x(cdiv_x(cla, ra), d0a), 0.05f);
                            dfe[c] = make_float2(fadd_rn(d0a.x, dq.x), fadd_rn(d0a.y, dq.y));
                        }
                        sym[c] = (uint8_t)ba;
                        if (TAPS && tap) {
                            if (p.sym_idx) p.sym_idx[vg * ncar + c] = (uint8_t)ba;
                            if (p.derot_syms) p.derot_syms[vg * ncar + c] = ra;
                        }
                    }
                    break;
                }
                const int c2 = c + TPS;                     // c < cbase + TPS < ncar
                const bool two = c2 < ncar;
                const int cb = two ? c2 : c;
                const float2 d0a = dfe[c], d0b = dfe[cb];
                const int ia = LDG(p.sinkmap + c), ib = LDG(p.sinkmap + cb);
                const float2 eqa = LDG(e + ia), eqb = LDG(e + ib);
                const float2 ra = cmul_x(cmul_x(eqa, car), d0a), rb = cmul_x(cmul_x(eqb, car), d0b);
                int ba, bb;
                if (MC == 0 && p.grid_L > 0) {
                    ba = slicer(ra);
                    bb = slicer(rb);
                } else {
                    // small constellations: both scans fused so the two chains interleave
                    ba = 0; bb = 0;
                    const float2 c0 = s_cst[0];
                    float besta = norm_x(csub_x(ra, c0)), bestb = norm_x(csub_x(rb, c0));
#pragma unroll
                    for (int k = 1; k < (MC ? MC : p.M); ++k) {
                        const float2 ck = s_cst[k];
                        const float dda = norm_x(csub_x(ra, ck));
                        const float ddb = norm_x(csub_x(rb, ck));
                        if (dda < besta) { besta = dda; ba = k; }
                        if (ddb < bestb) { bestb = ddb; bb = k; }
                    }
                }
                const float2 cla = s_cst[ba], clb = s_cst[bb];
                const float2 ea = cmulc_x(ra, cla), eb = cmulc_x(rb, clb);
                er += (double)ea.x;
                ei += (double)ea.y;
                if (norm_x(ra) > 0.001f) {
                    const float2 dq = cscale_x(csub_x(cdiv_x(cla, ra), d0a), 0.05f);      // eq_gain * (q - dfe)
                    dfe[c] = make_float2(fadd_rn(d0a.x, dq.x), fadd_rn(d0a.y, dq.y));
                }
                sym[c] = (uint8_t)ba;
                if (TAPS && tap) {
                    if (p.sym_idx) p.sym_idx[vg * ncar + c] = (uint8_t)ba;
                    if (p.derot_syms) p.derot_syms[vg * ncar + c] = ra;
                }
                if (two) {
                    er += (double)eb.x;
                    ei += (double)eb.y;
                    if (norm_x(rb) > 0.001f) {
                        const float2 dq = cscale_x(csub_x(cdiv_x(clb, rb), d0b), 0.05f);
                        dfe[c2] = make_float2(fadd_rn(d0b.x, dq.x), fadd_rn(d0b.y, dq.y));
                    }
                    sym[c2] = (uint8_t)bb;
                    if (TAPS && tap) {
                        if (p.sym_idx) p.sym_idx[vg * ncar + c2] = (uint8_t)bb;
                        if (p.derot_syms) p.derot_syms[vg * ncar + c2] = rb;
                    }
                }
            }
#pragma unroll
            for (int d = (TPS < 32 ? TPS / 2 : 16); d > 0; d >>= 1) {
                er += __shfl_xor_sync(gmask, er, d);
                ei += __shfl_xor_sync(gmask, ei, d);
            }
            if (!WARP) {                                    // warp sums -> shared memory (double-buffered by vi) -> everyone
                if ((threadIdx.x & 31) == 0) { s_red[vi & 1][threadIdx.x >> 5][0] = er; s_red[vi & 1][threadIdx.x >> 5][1] = ei; }
                __syncthreads();
                er = 0.0; ei = 0.0;
#pragma unroll
                for (int ww = 0; ww < TPS / 32; ++ww) { er += s_red[vi & 1][ww][0]; ei += s_red[vi & 1][ww][1]; }
            }
            // PLL update and the carrier of the next vector: every lane holds the same sums and computes the same values
            {
                const float angle = (float)atan2((double)(float)ei, (double)(float)er);
                freq = fsub_rn(freq, fmul_rn(0.015625f, angle));                 // freq_gain = 0.25^2/4
                float ph = fsub_rn(fadd_rn(phase, freq), fmul_rn(0.25f, angle)); // phase_gain = 0.25
                if ((double)ph >= 6.283185307179586) ph = (float)((double)ph - 6.283185307179586);
                if ((double)ph < 0.0) ph = (float)((double)ph + 6.283185307179586);
                phase = ph;
                car = expj_f32(ph);
            }
            sync();
            // LSB-first byte packing; bits left over from the previous vector sit in carry
            const int B0 = bit_base >> 3, B1 = (bit_base + bits_this) >> 3;
            pack_bytes<(MC == 4 ? 2 : (MC == 2 ? 1 : 0))>(sym, vb, B0, B1, bit_base, carry, lane, TPS, bd);
            sync();
            if (vi == 1) {
                const unsigned hdr = ((unsigned)vb[0] << 24) | ((unsigned)vb[1] << 16) | ((unsigned)vb[2] << 8) | vb[3];
                const bool hdr_ok = ((hdr >> 16) ^ (hdr & 0xFFFFu)) == 0u;
                len = hdr_ok ? (int)((hdr >> 16) & 0x0FFFu) : 0;
                if (!hdr_ok) { status = 1; nvec = 2; break; }
            }
            {
                const int nrb = bit_base + bits_this - 8 * B1;
                unsigned nc = 0;
                for (int i = 0; i < nrb; ++i) {
                    const int rel = 8 * B1 + i - bit_base;
                    const int c = bd.div(rel);
                    nc |= (((unsigned)sym[c] >> (rel - c * nbits)) & 1u) << i;
                }
                carry = nc;
            }
            for (int q = B0 + lane; q < B1; q += TPS) {
                const int pq = q - 4;
                if (pq >= 0 && pq < len && pq < p.pkt_stride) pkt_bytes[(size_t)f * p.pkt_stride + pq] = vb[q - B0];
            }
            sync();                                         // sym / vb are rewritten by the next vector
            bit_base += bits_this;
            if (B1 >= 4 + len) { status = 2; nvec = vi + 1; break; }
            // A header that announces more than a slot holds (max_pkt_bytes; a bogus length, e.g. of a header read
            // through a constellation turned by 90 degrees: both copies turn alike and still agree) cannot be delivered
            // intact: once every byte the slot keeps is written, nothing the remaining vectors decode is observable --
            // only WHERE the session ends, and that follows from the announced length alone.  (One such 83-vector
            // session walked to its end by 16 lanes was a 0.2 ms tail behind the whole kernel.)
            if (!TAPS && len > p.pkt_stride && B1 >= 4 + p.pkt_stride) {
                const int vi_end = (8 * (4 + len) + bits_this - 1) / bits_this;
                if (v0 + vi_end < vtot) { status = 2; nvec = vi_end + 1; }
                break;
            }
        }
        if (lane == 0) {
            frame_status[f] = (uint8_t)status;
            pkt_len[f] = status == 2 ? len : 0;
            sess_nvec[f] = nvec;
        }
        sync();
    }
}

// ---------------------------------------------------------------------------------------------
// acq_warp_kernel (N = 512 / 1024, no taps): the same stage on the warp plan of fft.cuh.  A frame lives on the
// T = N / 32 lanes of a GROUP (half a warp at 512: two frames per warp, walked in lock step), one warp per CTA:
// samples go from global memory straight into the radix-32 first pass (32 independent loads per lane; the next
// vector's lines are pulled into L2 meanwhile), ONE shared-memory exchange, and for data vectors the second pass
// hands its bins to the one-tap equaliser in registers -- they are stored to the workspace without touching shared
// memory again.  Only the flagged vector of a frame parks its (shifted) spectrum in shared memory for the correlation
// and the LS estimate -- in the exchange buffer itself (fft_pass WPRE), so a warp needs 12.2 KB and 16 warps fit an SM.
// No block barrier anywhere.
// ---------------------------------------------------------------------------------------------
// NARROW: the layout's occupied band +- the coarse search range lies inside the second-pass rows [0, RLO) and
// [RHI, R) (rows of 32 bins; 4 / 12 of 16 at N = 512, 7 / 25 of 32 at N = 1024: occupied_tones / fft_length <= 0.41):
// the rows in between are neither equalised nor stored -- at compile time, so the data-vector path has no branch.
template <int N, bool NARROW>
__global__ void __launch_bounds__(32, 16) acq_warp_kernel(const AcqParams p) {
    struct {
        const float2* y;
        const int64_t* trig_idx;
        const double* phi0;
        const double* step;
        const double* nco_init;
        const int32_t* n_trig;
        const int32_t* first_ok;
        const int32_t* n_frames;
        const int32_t* frame_ndata;
        const int64_t* vbase;
        const int64_t* nidx;
        float2* eq;
    } v;
    {
        const int64_t sidx = p.soff ? (int64_t)blockIdx.y : 0;
        const int64_t mf = p.max_frames;
        const int64_t a = p.soff ? p.soff[sidx] : 0;
        v.y = p.y + a;
        v.trig_idx = p.trig_idx + sidx * mf; v.phi0 = p.phi0 + sidx * mf; v.step = p.step + sidx * mf;
        v.nco_init = p.nco_init + sidx; v.n_trig = p.n_trig + sidx; v.first_ok = p.first_ok + sidx;
        v.n_frames = p.n_frames + sidx; v.frame_ndata = p.frame_ndata + sidx * mf;
        v.vbase = p.vbase + sidx * (mf + 1);
        v.eq = p.eq + sidx * p.eq_stride * p.occ;
        v.nidx = (p.n_nco[sidx] >= 0) ? p.nco_idx + sidx * mf : v.trig_idx;
    }
    using P = typename FftPlanW<N>::type;
    constexpr int T = N / 32, G = 32 / T;
    constexpr int SB = FFT_PAD32(N) + 2;
    constexpr int R1 = P::R[1];
    constexpr int RLO = NARROW ? (N == 512 ? 4 : 7) : R1, RHI = NARROW ? (N == 512 ? 12 : 25) : R1;
    const int occ = p.occ, zl = p.zl, L = p.L;
    const int HW = (occ + 1) & ~1;
    extern __shared__ float2 smem_w[];
    const int g = threadIdx.x / T, tid = threadIdx.x - g * T;
    float2* buf = smem_w + (size_t)g * (SB + HW + 32);
    float2* S = buf;                                        // the parked spectrum overwrites the exchange buffer (WPRE pass)
    float2* H = buf + SB;
    float2* Wt = H + HW;                                    // [32] e^{j step T r}: NCO turn between the rows of the first pass
    const unsigned gmask = (T >= 32) ? 0xffffffffu : (((1u << (T & 31)) - 1u) << (g * T));
    const int F = *v.n_frames;
    int K = *v.n_trig;
    if (K > p.max_frames) K = p.max_frames;
    const bool own_nco = v.nidx != v.trig_idx;
    if (own_nco) { K = p.n_nco[p.soff ? blockIdx.y : 0]; if (K > p.max_frames) K = p.max_frames; }
    const int first_ok = *v.first_ok;
    const double nco_init = *v.nco_init;
    const int64_t first_evt = (K > 0) ? v.nidx[0] : LLONG_MAX;

    for (int fb = blockIdx.x * G; fb < F; fb += gridDim.x * G) {
        const int f = fb + g;
        const bool have = f < F;
        const int kg = first_ok + (have ? f : fb);
        const int64_t t = v.trig_idx[kg];
        const int nd = have ? v.frame_ndata[f] : -1;
        const int64_t vb0 = have ? v.vbase[f] : 0;
        int nd_max = nd;
        if (G > 1) nd_max = max(nd_max, __shfl_xor_sync(0xffffffffu, nd_max, 16));
        int cnt = 1, delta = 0;
        // NCO segment of the current vector (event kk): its step, phase and position, and where the next event cuts it
        int kk = INT_MIN, kk_w = INT_MIN;
        double stp = 0.0, seg_phi = 0.0;
        int64_t seg_at = 0, t_lim = LLONG_MAX;
        auto segment = [&](int k2) {
            kk = k2;
            if (kk >= 0) { stp = v.step[kk]; seg_phi = v.phi0[kk]; seg_at = v.nidx[kk] - 1; t_lim = (kk + 1 < K) ? v.nidx[kk + 1] : LLONG_MAX; }
            else { stp = nco_init; seg_phi = 0.0; seg_at = -1; t_lim = first_evt; }
        };
        for (int m = 0; m <= nd_max; ++m) {
            const bool act = m <= nd;
            const int64_t st = t - N + 1 + (int64_t)m * L;
            const int64_t vglob = vb0 + m;
            // ---- sigmix: NCO phase of this lane's first sample, per-row turn table ----
            if (own_nco) {
                int lo = 0, hi = K;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (v.nidx[mid] <= st) lo = mid + 1; else hi = mid;
                }
                if (lo - 1 != kk) segment(lo - 1);
            } else if (m == 0) {
                int k2 = kg - 1;
                while (k2 >= 0 && v.nidx[k2] > st) --k2;
                segment(k2);
            } else if (m == 1) {
                segment(kg);
            }
            const double ph_base = seg_phi + stp * (double)(st + tid - seg_at);
            const bool retab = act && kk != kk_w;
            if (__any_sync(0xffffffffu, retab)) {
                if (retab) {
                    for (int r = tid; r < 32; r += T) Wt[r] = phasor_f64(stp * (double)(r * T));
                    kk_w = kk;
                }
                __syncwarp();
            }
            const int lim = (t_lim - st >= (int64_t)N) ? N : (t_lim > st ? (int)(t_lim - st) : 0);
            if (act && m < nd) {                            // the vector that follows: one 128-byte line per lane and step
                const float2* nx = v.y + st + L;
                for (int i = tid * 16; i < N; i += T * 16) asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + i));
            }
            // Samples at or behind the next NCO event take the exact per-sample phase (derot_slow).  In a flagged
            // vector that is the last sample (the trigger itself), i.e. the last row of the first pass: it is
            // settled first, then all 32 samples of the lane are requested before any is used.  An event deeper
            // inside a vector (a second trigger within a symbol) takes the plain per-point path.
            const int kk0 = kk < 0 ? 0 : kk;
            const bool deep = act && lim < N - T;
            float2 v31 = make_float2(0.f, 0.f);
            bool have31 = false;
            if (act && !deep && tid + 31 * T >= lim) {
                v31 = derot_slow(LDG(v.y + st + tid + 31 * T), st + tid + 31 * T, v.nidx, v.phi0, v.step, K, kk0);
                have31 = true;
            }
            if (__any_sync(0xffffffffu, deep)) {
                if (act) {
                    DemodLoad<false> ld{v.y + st, st, lim, v.nidx, v.phi0, v.step, K, kk0, phasor_f64(ph_base), Wt};
                    fft_pass<N, P::R[0], 1, -1, decltype(ld), SmemOut32, false, P>(tid, p.tw, ld, SmemOut32{buf});
                }
            } else if (act) {
                float2 raw[32];
                const float2* src = v.y + st + tid;
#pragma unroll
                for (int r = 0; r < 32; ++r) raw[r] = LDG(src + r * T);
                const float2 ph0 = phasor_f64(ph_base);
                // (plain packed multiply: what follows is a float32 FFT, the last-bit rounding order of its inputs is moot)
                auto ld = [&](int, int slot) -> float2 {
                    const float2 d = cmul(raw[slot], cmul(ph0, Wt[slot]));
                    return (slot == 31 && have31) ? v31 : d;
                };
                fft_pass<N, P::R[0], 1, -1, decltype(ld), SmemOut32, false, P>(tid, p.tw, ld, SmemOut32{buf});
            }
            __syncwarp();
            // A frame whose coarse search found nothing (delta = -zl, see below) keeps the parked path for its data vectors
            if (m == 0 || __any_sync(0xffffffffu, act && delta == -zl)) {
                // ---- flagged vector: park the (shifted) spectrum, correlate, estimate, equalise (ofdm_frame_acquisition).
                // The parked spectrum overwrites the exchange buffer: the pass reads all its inputs first (WPRE).
                // No shift with a positive correlation -- an all-zero or NaN spectrum -- leaves the upstream search at
                // index 0, i.e. delta = -zl ----
                {
                    // (run by every lane -- the pass has a warp barrier inside; a group without a vector stores nothing)
                    auto park = [&](int idx, float2 val, int) { if (act) S[(idx + N / 2) & (N - 1)] = val; };
                    fft_pass<N, R1, P::R[0], -1, SmemIn32, decltype(park), false, P, true>(tid, p.tw, SmemIn32{buf}, park);
                    __syncwarp();
                    if (m == 0 && act) {
                        const float2* Sc = S + (zl - OFDM_MAX_SHIFT);
                        double acc[2 * OFDM_MAX_SHIFT];
#pragma unroll
                        for (int s = 0; s < 2 * OFDM_MAX_SHIFT; ++s) acc[s] = 0.0;
                        for (int j = 2 * tid; j < occ - 2; j += 2 * T) {
                            const float kdj = LDG(p.kd + j);
#pragma unroll
                            for (int s = 0; s < 2 * OFDM_MAX_SHIFT; ++s) {
                                const float2 a = Sc[s + j], b = Sc[s + j + 2];
                                acc[s] += (double)kdj * (double)norm_x(csub_x(a, b));
                            }
                        }
                        float best = 0.f;
                        int index = -zl;
#pragma unroll
                        for (int s = 0; s < 2 * OFDM_MAX_SHIFT; ++s) {
                            double q = acc[s];
#pragma unroll
                            for (int d = T / 2; d > 0; d >>= 1) q += __shfl_xor_sync(gmask, q, d);
                            const float sf = (float)q;
                            if (sf > best) { best = sf; index = s - OFDM_MAX_SHIFT; }
                        }
                        delta = index;
                        cnt = 1;
                    }
                }
                const int sb = zl + delta;                  // parked position of equalised bin 0
                if (m == 0) {
                    if (act) {
                        const float2 c1 = coarse_comp(delta, p.cp, N, 1);
                        for (int i = 2 * tid; i < occ; i += 2 * T) {
                            const float2 b = cmul_x(c1, S[i + sb]);
                            H[i] = cdiv_x(make_float2(LDG(p.ks + i), 0.f), b);
                        }
                    }
                    __syncwarp();
                    if (act) {
                        for (int i = 2 * tid + 1; i + 1 < occ; i += 2 * T) {
                            const float2 a = H[i + 1], b = H[i - 1];
                            H[i] = cscale_x(cadd_x(a, b), 0.5f);
                        }
                        if (tid == 0 && (occ & 1) == 0) H[occ - 1] = H[occ - 2];
                    }
                    __syncwarp();
                }
                if (act) {
                    const float2 cc = coarse_comp(delta, p.cp, N, cnt);
                    float2* dst = v.eq + vglob * occ;
                    for (int i = tid; i < occ; i += T) {
                        const float2 Hi = H[i];
                        dst[i] = cmul_x(delta == 0 ? Hi : cmul_x(Hi, cc), S[i + sb]);
                    }
                }
            } else {
                // ---- data vector: one-tap equaliser applied to the second pass's registers.  comp(0, cnt) = (1, -0):
                // H * comp is then H itself up to the sign of an exact zero, which nothing downstream can see ----
                const float2 cc = coarse_comp(delta, p.cp, N, cnt);
                float2* dst = v.eq + vglob * occ;
                const int lo = zl + delta;
                auto equalise = [&](int idx, float2 val, int slot) {
                    const int row = slot & (R1 - 1);        // radix-R1 output row: bins [32 row, 32 row + 32)
                    if (row >= RLO && row < RHI) return;
                    const int i = ((idx + N / 2) & (N - 1)) - lo;
                    const bool ok = (unsigned)i < (unsigned)occ;
                    const float2 Hi = H[ok ? i : 0];
                    const float2 o = cmul(cmul(Hi, cc), val);   // (val is a float32 FFT output: last-bit rounding order is moot)
                    if (ok) dst[i] = o;
                };
                if (act) fft_pass<N, R1, P::R[0], -1, SmemIn32, decltype(equalise), false, P>(tid, p.tw, SmemIn32{buf}, equalise);
            }
            __syncwarp();
            ++cnt;
            if (cnt == OFDM_ACQ_MAX_SYMBOLS) cnt = 1;
        }
    }
}

template <int N>
static int launch_acq_warp(ofdm_handle* h, const AcqParams& p, int max_frames, int S, cudaStream_t st) {
    constexpr int G = 32 / (N / 32);
    const int HW = (p.occ + 1) & ~1;
    const size_t smem = sizeof(float2) * (size_t)G * (FFT_PAD32(N) + 2 + HW + 32);
    int grid = (h->sms * 64 + S - 1) / S;                  // 16 one-warp CTAs per SM, four waves
    const int want = (max_frames + G - 1) / G;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    // bins an equalised vector can come from: [zl - MAX_SHIFT, zl + occ + MAX_SHIFT) of the shifted spectrum
    const int hi_idx = p.zl + p.occ + OFDM_MAX_SHIFT - N / 2, lo_idx = p.zl - OFDM_MAX_SHIFT + N / 2;
    const int rlo = N == 512 ? 4 : 7, rhi = N == 512 ? 12 : 25;
    const bool narrow = hi_idx >= 0 && lo_idx < N && (hi_idx + 31) / 32 <= rlo && lo_idx / 32 >= rhi;
    if (narrow) {
        OFDM_SET_MAX_SMEM((acq_warp_kernel<N, true>), smem, h->device);
        acq_warp_kernel<N, true><<<dim3(grid, S), 32, smem, st>>>(p);
    } else {
        OFDM_SET_MAX_SMEM((acq_warp_kernel<N, false>), smem, h->device);
        acq_warp_kernel<N, false><<<dim3(grid, S), 32, smem, st>>>(p);
    }
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

template <int N, bool TAPS>
static int launch_acq_nt(ofdm_handle* h, const AcqParams& p, int max_frames, int S, cudaStream_t st) {
    constexpr int T = N / FftPlan<N>::E;
    constexpr int BT = T < 64 ? 64 : T;
    constexpr int NW = BT / 32;
    size_t smem = sizeof(double) * 8 * NW + sizeof(float2) * ((N == 4096 ? 1 : 2) * (size_t)fft_smem_elems<N>() + (size_t)p.occ);
    OFDM_SET_MAX_SMEM((acq_kernel<N, TAPS>), smem, h->device);
    int grid = (h->sms * 32 + S - 1) / S;                  // resident CTAs are shared by the streams
    if (grid > max_frames) grid = max_frames;
    if (grid < 1) grid = 1;
    acq_kernel<N, TAPS><<<dim3(grid, S), BT, smem, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

template <int N>
static int launch_acq_n(ofdm_handle* h, const AcqParams& p, int max_frames, int S, bool taps, cudaStream_t st) {
    return taps ? launch_acq_nt<N, true>(h, p, max_frames, S, st) : launch_acq_nt<N, false>(h, p, max_frames, S, st);
}

int launch_demod(ofdm_handle* h, const float2* y, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st,
                 int parts) {
    const bool taps = io->eq_syms || io->sym_idx || io->derot_syms || io->fft_out || io->sampler_out;
    if (taps && ss.S > 1) { ofdm_set_error("demod: the parity taps (eq_syms / sym_idx / derot_syms / ...) are single-stream only"); return OFDM_E_INVAL; }
    AcqParams a;
    a.y = y; a.soff = ss.off; a.n = ss.n_max; a.trig_idx = io->trig_idx; a.phi0 = ws->phi0; a.step = ws->step; a.nco_init = ws->nco_init;
    a.n_trig = io->n_trig; a.first_ok = ws->first_ok; a.n_frames = io->n_frames; a.frame_ndata = io->frame_ndata; a.vbase = ws->vbase;
    a.n_nco = ws->n_nco; a.nco_idx = ws->nco_idx;
    a.tw = h->d_tw; a.ks = h->d_ks; a.kd = h->d_kd;
    a.occ = h->occ; a.cp = h->cp; a.zl = h->zl; a.L = h->L; a.max_frames = io->max_frames;
    a.eq = ws->eq; a.eq_stride = ws->eq_stride;
    a.eq_tap = (float2*)io->eq_syms; a.fft_tap = (float2*)io->fft_out; a.samp_tap = (float2*)io->sampler_out;
    a.max_vectors = io->max_vectors;
    int rc = OFDM_OK;
    static const bool old_acq = getenv("OFDM_ACQ_OLD") != nullptr;
    if ((parts & 1) && !taps && !old_acq && h->d_tw_w && (h->N == 512 || h->N == 1024)) {
        a.tw = h->d_tw_w;
        rc = h->N == 512 ? launch_acq_warp<512>(h, a, io->max_frames, ss.S, st) : launch_acq_warp<1024>(h, a, io->max_frames, ss.S, st);
    } else
    if (parts & 1) switch (h->N) {
        case 64:   rc = launch_acq_n<64>(h, a, io->max_frames, ss.S, taps, st); break;
        case 128:  rc = launch_acq_n<128>(h, a, io->max_frames, ss.S, taps, st); break;
        case 256:  rc = launch_acq_n<256>(h, a, io->max_frames, ss.S, taps, st); break;
        case 512:  rc = launch_acq_n<512>(h, a, io->max_frames, ss.S, taps, st); break;
        case 1024: rc = launch_acq_n<1024>(h, a, io->max_frames, ss.S, taps, st); break;
        case 2048: rc = launch_acq_n<2048>(h, a, io->max_frames, ss.S, taps, st); break;
        case 4096: rc = launch_acq_n<4096>(h, a, io->max_frames, ss.S, taps, st); break;
        default:
            ofdm_set_error("demod: unsupported fft_length %d", h->N);
            return OFDM_E_INVAL;
    }
    if (rc) return rc;
    if (!(parts & 2)) return OFDM_OK;

    SinkParams p;
    p.eq = ws->eq; p.eq_stride = ws->eq_stride; p.per_stream = ss.off ? 1 : 0;
    p.n_frames = io->n_frames; p.vbase = ws->vbase; p.cst = h->d_const; p.sinkmap = h->d_sinkmap;
    p.occ = h->occ; p.ncar = h->ncar; p.nbits = h->nbits; p.M = h->M; p.max_frames = io->max_frames; p.pkt_stride = io->pkt_stride;
    p.grid_L = h->grid_L; p.grid_x0 = h->grid_x0; p.grid_y0 = h->grid_y0; p.grid_inv_dx = h->grid_inv_dx;
    p.grid_inv_dy = h->grid_inv_dy; p.grid = h->d_grid;
    p.frame_status = io->frame_status; p.pkt_len = io->pkt_len; p.sess_nvec = ws->sess_nvec; p.pkt_bytes = io->pkt_bytes;
    p.sym_idx = io->sym_idx; p.derot_syms = (float2*)io->derot_syms; p.max_vectors = io->max_vectors;
    // threads per session: a warp (several sessions per CTA, as many as keep the per-session tables -- DFE taps, symbols,
    // bytes -- within ~48 KB) or, for vectors with more than 1024 data carriers, a whole 128-thread CTA
    const size_t per_sess = (size_t)h->ncar * sizeof(float2) + ((h->ncar + 15) & ~15) + ((h->ncar * h->nbits / 8 + 16 + 15) & ~15);
    const size_t fixed = (size_t)h->M * sizeof(float2) + (((size_t)h->grid_L * h->grid_L + 15) & ~(size_t)15);
    const bool wide = h->ncar > 1024;
    // half-warp sessions while a session's tables are small (measured, sink kernel alone: 512/200 QPSK 1.50 -> 1.26 ms,
    // QAM16 1.17 -> 1.05; at 398 carriers two sessions' tables per warp cost more occupancy than the shared serial
    // section saves: 0.55 -> 0.60)
    const int tps = wide ? 128 : (h->ncar <= 256 ? 16 : 32);
    int W = wide ? 1 : 128 / tps;
    while (W > 1 && fixed + W * per_sess > 48 * 1024) W >>= 1;
    const size_t smem = fixed + W * per_sess;
    const bool staps = io->sym_idx || io->derot_syms;
    int grid = (io->max_frames + W - 1) / W;
    const int cap = (h->sms * 16 + ss.S - 1) / ss.S;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const dim3 g(grid, ss.S);
#define OFDM_SINK_LAUNCH_M(TAPS_, TPS_, MC_)                                                \
    do {                                                                                    \
        OFDM_SET_MAX_SMEM((sink_kernel<TAPS_, TPS_, MC_>), smem, h->device);                \
        sink_kernel<TAPS_, TPS_, MC_><<<g, (TPS_) <= 32 ? (TPS_) * W : 128, smem, st>>>(p); \
    } while (0)
#define OFDM_SINK_LAUNCH(TAPS_, TPS_) OFDM_SINK_LAUNCH_M(TAPS_, TPS_, 0)
    if (wide) { if (staps) OFDM_SINK_LAUNCH(true, 128); else OFDM_SINK_LAUNCH(false, 128); }
    else if (tps == 16 && !staps && h->M == 4) OFDM_SINK_LAUNCH_M(false, 16, 4);          // qpsk, the bench layout
    else if (tps == 16 && !staps && h->M == 2) OFDM_SINK_LAUNCH_M(false, 16, 2);
    else if (tps == 16) { if (staps) OFDM_SINK_LAUNCH(true, 16); else OFDM_SINK_LAUNCH(false, 16); }
    else { if (staps) OFDM_SINK_LAUNCH(true, 32); else OFDM_SINK_LAUNCH(false, 32); }
#undef OFDM_SINK_LAUNCH
#undef OFDM_SINK_LAUNCH_M
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// ---------------------------------------------------------------------------------------------
// options.log taps of the receiver (ofdm_receiver.py~:150-151): the NCO output e^{j phi[n]} (frequency_modulator_fc
// driven by the held angle, A.8) and sigmix = chan_filt * nco, per sample.  Debugging aid, not on the hot path.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nco_tap_kernel(const float2* __restrict__ y, int64_t n, const int32_t* __restrict__ n_trig,
                                                      const int32_t* __restrict__ n_nco, int max_frames,
                                                      const int64_t* __restrict__ trig_in, const int64_t* __restrict__ nco_idx,
                                                      const double* __restrict__ phi0, const double* __restrict__ step,
                                                      const double* __restrict__ nco_init, float2* __restrict__ nco_out,
                                                      float2* __restrict__ sigmix_out) {
    const bool own = *n_nco >= 0;                      // ofdm_sync_ml: the NCO has its own event list
    const int64_t* trig = own ? nco_idx : trig_in;
    int K = own ? *n_nco : *n_trig;
    if (K > max_frames) K = max_frames;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = K;                              // triggers at indices <= i
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (trig[mid] <= i) lo = mid + 1; else hi = mid;
        }
        const int k = lo - 1;
        const double ph = (k >= 0) ? phi0[k] + step[k] * (double)(i - trig[k] + 1) : *nco_init * (double)(i + 1);
        const float2 w = phasor_f64(ph);
        if (nco_out) nco_out[i] = w;
        if (sigmix_out) sigmix_out[i] = cmul_x(y[i], w);
    }
}

int launch_nco_taps(ofdm_handle* h, const float2* y, int64_t n, ofdm_rx_io* io, RxWorkspace* ws, float2* nco_out,
                    float2* sigmix_out, cudaStream_t st) {
    if (n <= 0 || (!nco_out && !sigmix_out)) return OFDM_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)h->sms * 16) blocks = (int64_t)h->sms * 16;
    nco_tap_kernel<<<(int)blocks, 256, 0, st>>>(y, n, io->n_trig, ws->n_nco, io->max_frames, io->trig_idx, ws->nco_idx, ws->phi0,
                                               ws->step, ws->nco_init, nco_out, sigmix_out);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// ---------------------------------------------------------------------------------------------
// Frame-sink liveness: the sink only starts on a preamble it sees in SYNC_SEARCH; a session started at
// frame f swallows sess_nvec[f] vectors of the sampler's stream, flagged ones included (A.11).  The live
// frames are the orbit of frame 0 under next(f) = first frame whose preamble lies at or after the end of
// f's session.
//   next_kernel      (grid-wide)  next[f]; every frame presumed live; frames with next(f) != f+1 are flagged.
//   liveness_fast    (one CTA)    ordered list of the flagged frames (normally a handful: a session only
//                                 swallows a preamble after a bogus header or an early re-sync), serial hop
//                                 over that list in shared memory, dead ranges cleared in parallel.
//   liveness_kernel  (one CTA)    the general chunked pointer walk, run only if the list overflows.
// ---------------------------------------------------------------------------------------------
constexpr int LIVE_CAP = 3072;

__global__ void __launch_bounds__(256) next_kernel(const int32_t* __restrict__ n_frames, const int64_t* __restrict__ vbase,
                                                   int64_t vbase_stride, const int32_t* __restrict__ sess_nvec,
                                                   int32_t* __restrict__ next, uint8_t* __restrict__ exc,
                                                   uint8_t* __restrict__ live, int max_frames) {
    {
        const int64_t s = blockIdx.y;                   // per-stream tables
        n_frames += s; vbase += s * vbase_stride; sess_nvec += s * max_frames; next += s * max_frames;
        exc += s * max_frames * 4; live += s * max_frames;   // exc aliases the int32 exit table: 4 bytes per frame
    }
    const int F = *n_frames;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < max_frames; f += gridDim.x * blockDim.x) {
        if (f >= F) { live[f] = 0; continue; }
        const int64_t target = vbase[f] + (int64_t)sess_nvec[f];
        int nx;
        if (f + 1 >= F) nx = F;
        else if (vbase[f + 1] >= target) nx = f + 1;
        else {
            int lo = f + 1, hi = F;                 // first g in (f, F) with vbase[g] >= target, else F
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (vbase[mid] >= target) hi = mid; else lo = mid + 1;
            }
            nx = lo;
        }
        next[f] = nx;
        exc[f] = (nx != f + 1) ? 1 : 0;
        live[f] = 1;
    }
}

__global__ void __launch_bounds__(1024) liveness_fast_kernel(const int32_t* __restrict__ n_frames,
                                                             const int32_t* __restrict__ next,
                                                             const uint8_t* __restrict__ exc, uint8_t* __restrict__ live,
                                                             int32_t* __restrict__ overflow, int force_general, int max_frames) {
    {
        const int64_t s = blockIdx.x;                   // one CTA per stream
        n_frames += s; next += s * max_frames; exc += s * max_frames * 4; live += s * max_frames; overflow += s;
    }
    __shared__ int s_e[LIVE_CAP], s_nx[LIVE_CAP];   // flagged frames in order, and their next(); reused as dead ranges
    __shared__ int s_w[32];
    __shared__ int s_total, s_ndead;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int F = *n_frames;
    if (force_general) {
        if (tid == 0) *overflow = 1;
        return;
    }
    if (tid == 0) { s_total = 0; s_ndead = 0; *overflow = 0; }
    __syncthreads();
    // ordered compaction of the flags, 8 frames per thread and pass
    for (int base = 0; base < F; base += 1024 * 8) {
        const int f0 = base + tid * 8;
        unsigned bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (f0 + i < F && exc[f0 + i]) bits |= 1u << i;
        const int cnt = __popc(bits);
        int inc = cnt;
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
        int pos = s_total + s_w[w] + inc - cnt;
        for (unsigned b = bits; b; b &= b - 1) {
            if (pos < LIVE_CAP) { const int f = f0 + __ffs(b) - 1; s_e[pos] = f; s_nx[pos] = next[f]; }
            ++pos;
        }
        __syncthreads();
        if (tid == 1023) s_total = pos;             // the last thread's end = running total
        __syncthreads();
    }
    const int total = s_total;
    if (total > LIVE_CAP) {                         // the general walk takes over
        if (tid == 0) *overflow = 1;
        return;
    }
    if (tid == 0) {
        int cur = 0, nd = 0;
        for (int i = 0; i < total; ++i) {
            const int e = s_e[i];
            if (e < cur) continue;                  // swallowed by an earlier session: its own next() is moot
            const int nx = s_nx[i];
            if (nx > e + 1) { s_e[nd] = e + 1; s_nx[nd] = nx; ++nd; }   // nd <= i: never overtakes the read position
            cur = nx;
        }
        s_ndead = nd;
    }
    __syncthreads();
    for (int r = tid; r < s_ndead; r += 1024)
        for (int f = s_e[r]; f < s_nx[r]; ++f) live[f] = 0;
}

__global__ void __launch_bounds__(1024) liveness_kernel(const int32_t* __restrict__ n_frames,
                                                        const int32_t* __restrict__ run_flag,
                                                        const int32_t* __restrict__ next, int32_t* __restrict__ exitf,
                                                        uint8_t* __restrict__ live, int max_frames) {
    {
        const int64_t s = blockIdx.x;                   // one CTA per stream
        n_frames += s; run_flag += s; next += s * max_frames; exitf += s * max_frames; live += s * max_frames;
    }
    __shared__ int s_entry[1024];
    __shared__ int s_first_exit[1024];
    const int tid = threadIdx.x;
    const int F = *n_frames;
    if (*run_flag == 0 || F <= 0) return;
    s_entry[tid] = -1;
    for (int f = tid; f < F; f += 1024) live[f] = 0;
    __syncthreads();
    const int Kc = (F + 1023) / 1024;
    const int cs = tid * Kc;
    const int ce = (cs + Kc < F) ? cs + Kc : F;
    for (int f = ce - 1; f >= cs; --f) {
        const int nx = next[f];
        exitf[f] = (nx < ce) ? exitf[nx] : nx;
    }
    // exit of each chunk when entered at its first frame (the common case) -> shared memory, so that the
    // serial hop over the chunks below is a chain of shared-memory reads, not of global-memory round trips
    s_first_exit[tid] = (cs < F) ? exitf[cs] : F;
    __syncthreads();
    if (tid == 0) {
        int e = 0;
        while (e < F) {
            const int c = e / Kc;
            s_entry[c] = e;
            e = (e == c * Kc) ? s_first_exit[c] : exitf[e];
        }
    }
    __syncthreads();
    int f = s_entry[tid];
    if (f >= 0)
        while (f < ce) {
            live[f] = 1;
            f = next[f];
        }
}

// unmake_packet: dewhiten (offset 0) + check_crc32 for every delivered message; counters for the stats.
// One warp per 32 consecutive packet slots: when 32 slots fit, they are staged through shared memory (coalesced
// word copies, rows padded by one word against bank conflicts) and each lane works on its own row there.
constexpr int CRC_SMEM_MAX = 160 * 1024;  // staging is used while 32 padded rows fit (any stride up to 4096 + 16 does)

// b: the packet's row (4-byte aligned when WORDS); s_crc: [T0 | T1 | T2 | T3].  WORDS: dewhitening and CRC run a
// 32-bit word at a time over the part of the CRC body made of whole words (four independent table lookups per step
// instead of a dependent chain of four), the rest -- at most 3 body bytes and the 4 CRC bytes -- byte by byte.
template <bool WORDS>
__device__ __forceinline__ uint8_t dewhiten_crc_row(uint8_t* b, int len, int stride, const uint8_t* __restrict__ mask,
                                                    const uint32_t* s_crc) {
    const int nst = len < stride ? len : stride;
    uint32_t crc = 0xFFFFFFFFu, tail = 0;
    int i = 0;
    if (WORDS) {
        const int body = len - 4 < nst ? len - 4 : nst;
        const int nw = body > 0 ? body >> 2 : 0;
        uint32_t* w = (uint32_t*)b;
        const uint32_t* m4 = (const uint32_t*)mask;
        for (int k = 0; k < nw; ++k) {
            const uint32_t v = w[k] ^ LDG(m4 + (k & 1023));
            w[k] = v;
            crc = crc32_step4(crc, __byte_perm(v, 0, 0x0123), s_crc);
        }
        i = 4 * nw;
    }
    for (; i < nst; ++i) {
        const uint8_t v = (uint8_t)(b[i] ^ mask[i & 4095]);
        b[i] = v;
        if (i < len - 4) crc = s_crc[(v ^ (crc >> 24)) & 0xFF] ^ (crc << 8);
        else tail = (tail << 8) | v;
    }
    return (len >= 4 && len <= stride && (~crc) == tail) ? 1 : 0;
}

__global__ void __launch_bounds__(32) crc_kernel(const int32_t* __restrict__ n_frames, const uint8_t* __restrict__ live,
                                                 const uint8_t* __restrict__ status, const int32_t* __restrict__ pkt_len,
                                                 uint8_t* __restrict__ pkt_bytes, int stride, uint8_t* __restrict__ pkt_ok,
                                                 const uint8_t* __restrict__ mask, const uint32_t* __restrict__ crctab,
                                                 int64_t* __restrict__ counters, int staged_in, int max_frames) {
    {
        const int64_t s = blockIdx.y;                   // per-stream tables
        n_frames += s; live += s * max_frames; status += s * max_frames; pkt_len += s * max_frames;
        pkt_bytes += s * max_frames * (int64_t)stride; pkt_ok += s * max_frames; counters += s * 8;
    }
    __shared__ uint32_t s_crc[1024];                              // T0 .. T3
    extern __shared__ uint32_t s_rows[];                          // [32][stride/4 + 1] when staged, else empty
    const int lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) s_crc[i] = crctab[i];
    for (int i = lane; i < 768; i += 32) s_crc[256 + i] = crctab[OFDM_CRC_SLICE + i];
    __syncwarp();
    const int F = *n_frames;
    const int wpr = stride >> 2, spr = wpr + 1;                   // words per row in global / shared memory
    const bool staged = staged_in != 0;
    for (int f0 = blockIdx.x * 32; f0 < F; f0 += gridDim.x * 32) {
        const int f = f0 + lane;
        const bool mine = f < F && live[f] && status[f] == 2;
        const unsigned any = __ballot_sync(0xffffffffu, mine);
        uint8_t ok = 0;
        int len = 0;
        if (any) {
            if (staged) {
                const int nf = (F - f0 < 32) ? F - f0 : 32;
                uint32_t* g = (uint32_t*)(pkt_bytes + (size_t)f0 * stride);
                const int nw = nf * wpr;                          // the 32 slots are contiguous in global memory
#pragma unroll 4
                for (int i = lane; i < nw; i += 32) {
                    const int r = i / wpr;
                    s_rows[i + r] = g[i];                         // row r starts at r*spr = r*wpr + r
                }
                __syncwarp();
                if (mine) { len = pkt_len[f]; ok = dewhiten_crc_row<true>((uint8_t*)(s_rows + lane * spr), len, stride, mask, s_crc); }
                __syncwarp();
#pragma unroll 4
                for (int i = lane; i < nw; i += 32) {
                    const int r = i / wpr;
                    if ((any >> r) & 1u) g[i] = s_rows[i + r];
                }
                __syncwarp();
            } else if (mine) {
                len = pkt_len[f];
                ok = dewhiten_crc_row<false>(pkt_bytes + (size_t)f * stride, len, stride, mask, s_crc);
            }
        }
        if (f < F) pkt_ok[f] = ok;
        // warp-aggregated counters: messages, CRC ok, payload bytes of the good ones
        const unsigned okm = __ballot_sync(0xffffffffu, ok != 0);
        int pay = ok ? len - 4 : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) pay += __shfl_xor_sync(0xffffffffu, pay, d);
        if (lane == 0 && any) {
            atomicAdd((unsigned long long*)&counters[1], (unsigned long long)__popc(any));
            if (okm) {
                atomicAdd((unsigned long long*)&counters[2], (unsigned long long)__popc(okm));
                atomicAdd((unsigned long long*)&counters[3], (unsigned long long)pay);
            }
        }
    }
}

// unmake_packet for long packets: one warp per packet.  CRC-32 is linear over GF(2): lane l runs the table CRC over
// its 128-byte slice (lane 0 from the all-ones preset, the others from 0) and the slice CRCs are combined after
// multiplying each by x^(8 * bytes after the slice) mod P, the factor taken from two host tables (x^(1024 j), x^(8 r)).
__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b) {
    uint32_t r = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; --i) {
        r = (r << 1) ^ ((r & 0x80000000u) ? 0x04C11DB7u : 0u);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

__global__ void __launch_bounds__(128) crc_warp_kernel(const int32_t* __restrict__ n_frames, const uint8_t* __restrict__ live,
                                                       const uint8_t* __restrict__ status, const int32_t* __restrict__ pkt_len,
                                                       uint8_t* __restrict__ pkt_bytes, int stride, uint8_t* __restrict__ pkt_ok,
                                                       const uint8_t* __restrict__ mask, const uint32_t* __restrict__ crctab,
                                                       int64_t* __restrict__ counters, int max_frames) {
    {
        const int64_t s = blockIdx.y;                   // per-stream tables
        n_frames += s; live += s * max_frames; status += s * max_frames; pkt_len += s * max_frames;
        pkt_bytes += s * max_frames * (int64_t)stride; pkt_ok += s * max_frames; counters += s * 8;
    }
    __shared__ uint32_t s_crc[256 + 32 + 128];
    __shared__ uint32_t s_mask[1024];
    for (int i = threadIdx.x; i < 256 + 32 + 128; i += blockDim.x) s_crc[i] = crctab[i];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_mask[i] = ((const uint32_t*)mask)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int F = *n_frames;
    const int wpb = blockDim.x >> 5;
    for (int f = blockIdx.x * wpb + (threadIdx.x >> 5); f < F; f += gridDim.x * wpb) {
        const bool mine = live[f] && status[f] == 2;
        uint8_t ok = 0;
        int len = 0;
        if (mine) {
            len = pkt_len[f];
            const int nst = len < stride ? len : stride;           // bytes present in the slot
            const int Lc = len - 4;                                 // bytes under the CRC
            uint32_t* row = (uint32_t*)(pkt_bytes + (size_t)f * stride);
            const int b0 = lane * 128;
            uint32_t crc = (lane == 0) ? 0xFFFFFFFFu : 0u;
            uint32_t tail = 0;                                      // this lane's share of the 4 CRC bytes
#pragma unroll 4
            for (int w = 0; w < 32; ++w) {
                const int b = b0 + 4 * w;
                if (b >= nst) break;
                const uint32_t keep = (b + 4 <= nst) ? 0xFFFFFFFFu : ((1u << (8 * (nst - b))) - 1u);
                const uint32_t v = row[b >> 2] ^ (s_mask[(b >> 2) & 1023] & keep);     // bytes past nst are left alone
                row[b >> 2] = v;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t byte = (v >> (8 * k)) & 0xFFu;
                    const int i = b + k;
                    if (i < Lc) crc = s_crc[(byte ^ (crc >> 24)) & 0xFFu] ^ (crc << 8);
                    else if (Lc >= 0 && i < len && i < nst) tail |= byte << (8 * (3 - (i - Lc)));
                }
            }
            // bytes of the CRC'd region that follow this lane's slice
            int after = Lc - (b0 + 128);
            if (after < 0) after = 0;
            if (b0 >= Lc) crc = 0;                                  // no slice (lane 0 always has one when Lc > 0)
            else crc = crc_mulmod(crc_mulmod(crc, s_crc[256 + (after >> 7)]), s_crc[256 + 32 + (after & 127)]);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                crc ^= __shfl_xor_sync(0xffffffffu, crc, d);
                tail |= __shfl_xor_sync(0xffffffffu, tail, d);
            }
            if (Lc <= 0) crc = 0xFFFFFFFFu;                         // empty payload: the preset itself
            ok = (len >= 4 && len <= stride && (~crc) == tail) ? 1 : 0;
        }
        if (lane == 0) {
            pkt_ok[f] = ok;
            if (mine) {
                atomicAdd((unsigned long long*)&counters[1], 1ull);
                if (ok) {
                    atomicAdd((unsigned long long*)&counters[2], 1ull);
                    atomicAdd((unsigned long long*)&counters[3], (unsigned long long)(len - 4));
                }
            }
        }
    }
}

int launch_liveness(int sms, int S, const int32_t* n_frames, const int64_t* vbase, int64_t vbase_stride, const int32_t* sess_nvec,
                    int32_t max_frames, int32_t* next, int32_t* exitf, int32_t* overflow, uint8_t* live, int force_general,
                    cudaStream_t st) {
    // exitf doubles as the flag bytes of next_kernel (the general walk overwrites it only after the fast path is done)
    uint8_t* exc = (uint8_t*)exitf;
    int ngrid = (max_frames + 255) / 256;
    const int cap = (sms * 8 + S - 1) / S;
    if (ngrid > cap) ngrid = cap;
    if (ngrid < 1) ngrid = 1;
    next_kernel<<<dim3(ngrid, S), 256, 0, st>>>(n_frames, vbase, vbase_stride, sess_nvec, next, exc, live, max_frames);
    OFDM_LAUNCH_CHECK();
    liveness_fast_kernel<<<S, 1024, 0, st>>>(n_frames, next, exc, live, overflow, force_general, max_frames);
    OFDM_LAUNCH_CHECK();
    liveness_kernel<<<S, 1024, 0, st>>>(n_frames, overflow, next, exitf, live, max_frames);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

int launch_finish(ofdm_handle* h, int S, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st) {
    int rc = launch_liveness(h->sms, S, io->n_frames, ws->vbase, (int64_t)io->max_frames + 1, ws->sess_nvec, io->max_frames,
                             ws->next_frame, ws->exit_frame, ws->live_overflow, io->frame_live, 0, st);
    if (rc) return rc;
    if (io->pkt_stride >= 1024 && (io->pkt_stride & 3) == 0 && ((((uintptr_t)io->pkt_bytes) & 3) == 0)) {
        int wgrid = (io->max_frames + 3) / 4;                       // long packets: a warp each
        const int cap = (h->sms * 16 + S - 1) / S;
        if (wgrid > cap) wgrid = cap;
        crc_warp_kernel<<<dim3(wgrid, S), 128, 0, st>>>(io->n_frames, io->frame_live, io->frame_status, io->pkt_len, io->pkt_bytes,
                                                        io->pkt_stride, io->pkt_ok, h->d_mask, h->d_crctab, io->counters,
                                                        io->max_frames);
        OFDM_LAUNCH_CHECK();
        return OFDM_OK;
    }
    int grid = (io->max_frames + 31) / 32;
    const int cap = (h->sms * 32 + S - 1) / S;
    if (grid > cap) grid = cap;
    const size_t row_smem = (size_t)32 * (io->pkt_stride / 4 + 1) * sizeof(uint32_t);
    const int staged = (io->pkt_stride & 3) == 0 && row_smem <= (size_t)CRC_SMEM_MAX && ((((uintptr_t)io->pkt_bytes) & 3) == 0);
    if (staged) OFDM_SET_MAX_SMEM(crc_kernel, row_smem, h->device);
    crc_kernel<<<dim3(grid, S), 32, staged ? row_smem : 0, st>>>(io->n_frames, io->frame_live, io->frame_status, io->pkt_len,
                                                                 io->pkt_bytes, io->pkt_stride, io->pkt_ok, h->d_mask,
                                                                 h->d_crctab, io->counters, staged, io->max_frames);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// ---------------------------------------------------------------------------------------------
// Hand-over of the delivered messages to the host (the reference pushes each message into a gr.msg_queue that
// _queue_watcher_thread pops, ofdm.py:290-305): the messages of all streams, in stream / arrival order, packed into
// one dense byte array with an offset table and a bit-packed CRC verdict -- so the device -> host copy is sized by
// what was delivered, not by max_frames * pkt_stride.
//   compact_count_kernel   per 1024-frame block: messages and bytes
//   compact_scan_kernel    (one CTA) exclusive scan of the block totals, grand totals
//   compact_copy_kernel    per block: in-block prefix, offsets / frame ids / verdict bits, one warp per message row
// ---------------------------------------------------------------------------------------------
struct CompactParams {
    const int32_t* n_frames;     // [S]
    const uint8_t* live;         // [S][max_frames]
    const uint8_t* status;
    const int32_t* pkt_len;
    const uint8_t* pkt_ok;
    const uint8_t* pkt_bytes;    // [S][max_frames][stride]
    int S, max_frames, stride;
    int64_t capacity;            // bytes in out_bytes
    int64_t* blk_cnt;            // [nblk] then, after the scan, exclusive prefixes
    int64_t* blk_bytes;          // [nblk]
    uint8_t* out_bytes;
    int64_t* msg_off;            // [n_msgs + 1]
    int32_t* msg_frame;          // [n_msgs] global frame slot s * max_frames + f
    uint32_t* ok_bits;           // [(S * max_frames + 31) / 32], bit m = CRC verdict of message m
    int64_t* totals;             // [0] messages, [1] bytes (uncapped), [2] messages copied
};

__device__ __forceinline__ bool compact_flag(const CompactParams& p, int64_t g, int& len) {
    const int s = (int)(g / p.max_frames), f = (int)(g - (int64_t)s * p.max_frames);
    len = 0;
    if (s >= p.S || f >= p.n_frames[s]) return false;
    if (!(p.live[g] && p.status[g] == 2)) return false;
    const int l = p.pkt_len[g];
    len = l < p.stride ? l : p.stride;
    if (len < 0) len = 0;
    return true;
}

__global__ void __launch_bounds__(1024) compact_count_kernel(const CompactParams p) {
    __shared__ long long s_c[32], s_b[32];
    const int64_t g = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int len;
    const bool m = compact_flag(p, g, len);
    long long c = m ? 1 : 0, b = len;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = c; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        c = s_c[threadIdx.x]; b = s_b[threadIdx.x];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
        if (threadIdx.x == 0) { p.blk_cnt[blockIdx.x] = c; p.blk_bytes[blockIdx.x] = b; }
    }
}

__global__ void __launch_bounds__(1024) compact_scan_kernel(const CompactParams p, int nblk) {
    __shared__ long long s_c[33], s_b[33];
    __shared__ long long s_cc, s_cb;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_cc = 0; s_cb = 0; }
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const long long vc = i < nblk ? p.blk_cnt[i] : 0, vb = i < nblk ? p.blk_bytes[i] : 0;
        long long ic = vc, ib = vb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long oc = __shfl_up_sync(0xffffffffu, ic, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ic += oc; ib += ob; }
        }
        if (lane == 31) { s_c[w] = ic; s_b[w] = ib; }
        __syncthreads();
        if (w == 0) {
            long long tc = s_c[lane], tb = s_b[lane];
            const long long oc0 = tc, ob0 = tb;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long oc = __shfl_up_sync(0xffffffffu, tc, d), ob = __shfl_up_sync(0xffffffffu, tb, d);
                if (lane >= d) { tc += oc; tb += ob; }
            }
            s_c[lane] = tc - oc0; s_b[lane] = tb - ob0;
            if (lane == 31) { s_c[32] = tc; s_b[32] = tb; }
        }
        __syncthreads();
        if (i < nblk) { p.blk_cnt[i] = s_cc + s_c[w] + ic - vc; p.blk_bytes[i] = s_cb + s_b[w] + ib - vb; }
        __syncthreads();
        if (threadIdx.x == 0) { s_cc += s_c[32]; s_cb += s_b[32]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { p.totals[0] = s_cc; p.totals[1] = s_cb; p.msg_off[s_cc] = s_cb; }
}

__global__ void __launch_bounds__(1024) compact_copy_kernel(const CompactParams p) {
    __shared__ long long s_c[33], s_b[33];
    __shared__ int s_len[1024];
    __shared__ long long s_dst[1024];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int len;
    const bool m = compact_flag(p, g, len);
    const long long vc = m ? 1 : 0, vb = len;
    long long ic = vc, ib = vb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long oc = __shfl_up_sync(0xffffffffu, ic, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ic += oc; ib += ob; }
    }
    if (lane == 31) { s_c[w] = ic; s_b[w] = ib; }
    __syncthreads();
    if (w == 0) {
        long long tc = s_c[lane], tb = s_b[lane];
        const long long oc0 = tc, ob0 = tb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long oc = __shfl_up_sync(0xffffffffu, tc, d), ob = __shfl_up_sync(0xffffffffu, tb, d);
            if (lane >= d) { tc += oc; tb += ob; }
        }
        s_c[lane] = tc - oc0; s_b[lane] = tb - ob0;
    }
    __syncthreads();
    const long long mi = p.blk_cnt[blockIdx.x] + s_c[w] + ic - vc;        // message number
    const long long bo = p.blk_bytes[blockIdx.x] + s_b[w] + ib - vb;      // byte offset
    s_len[threadIdx.x] = -1;
    if (m) {
        p.msg_off[mi] = bo;
        p.msg_frame[mi] = (int32_t)g;
        if (p.pkt_ok[g]) atomicOr(p.ok_bits + (mi >> 5), 1u << (mi & 31));
        if (bo + len <= p.capacity) { s_len[threadIdx.x] = len; s_dst[threadIdx.x] = bo; atomicAdd((unsigned long long*)&p.totals[2], 1ull); }
    }
    __syncthreads();
    // one warp per message row of this block
    for (int r = w; r < 1024; r += 32) {
        const int l = s_len[r];
        if (l <= 0) continue;
        const uint8_t* src = p.pkt_bytes + ((int64_t)blockIdx.x * 1024 + r) * p.stride;
        uint8_t* dst = p.out_bytes + s_dst[r];
        if (((((uintptr_t)src) | ((uintptr_t)dst)) & 3) == 0) {
            const int nw = l >> 2;
            for (int i = lane; i < nw; i += 32) ((uint32_t*)dst)[i] = ((const uint32_t*)src)[i];
            for (int i = 4 * nw + lane; i < l; i += 32) dst[i] = src[i];
        } else {
            for (int i = lane; i < l; i += 32) dst[i] = src[i];
        }
    }
}

int launch_compact(ofdm_handle* h, const ofdm_rx_io* io, int32_t n_streams, uint8_t* out_bytes, int64_t capacity,
                   int64_t* msg_off, int32_t* msg_frame, uint32_t* ok_bits, int64_t* totals, int64_t* scratch,
                   cudaStream_t st) {
    (void)h;
    CompactParams p;
    p.n_frames = io->n_frames; p.live = io->frame_live; p.status = io->frame_status; p.pkt_len = io->pkt_len;
    p.pkt_ok = io->pkt_ok; p.pkt_bytes = io->pkt_bytes; p.S = n_streams; p.max_frames = io->max_frames; p.stride = io->pkt_stride;
    p.capacity = capacity; p.out_bytes = out_bytes; p.msg_off = msg_off; p.msg_frame = msg_frame; p.ok_bits = ok_bits;
    p.totals = totals;
    const int64_t slots = (int64_t)n_streams * io->max_frames;
    const int nblk = (int)((slots + 1023) / 1024);
    p.blk_cnt = scratch; p.blk_bytes = scratch + nblk;
    OFDM_CUDA_CHECK(cudaMemsetAsync(ok_bits, 0, sizeof(uint32_t) * (size_t)((slots + 31) / 32), st));
    OFDM_CUDA_CHECK(cudaMemsetAsync(totals, 0, 3 * sizeof(int64_t), st));
    compact_count_kernel<<<nblk, 1024, 0, st>>>(p);
    OFDM_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, st>>>(p, nblk);
    OFDM_LAUNCH_CHECK();
    compact_copy_kernel<<<nblk, 1024, 0, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}
