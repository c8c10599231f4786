// Receive front end: channel filter, ofdm_sync_pn (Schmidl-Cox metric + peak detector + angle latch),
// and the sampler / NCO plan.  Reference wiring: ofdm_receiver.py~:69-76,97-101,123-125,131-136.
#include "internal.h"
#include <stdlib.h>
#include "fft.cuh"
#include <limits.h>
#include <vector>

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int rx_workspace_layout(const ofdm_handle* h, const StreamSet& ss, int32_t max_frames, void* base, size_t bytes,
                        RxWorkspace* ws, size_t* need) {
    const int64_t n_total = ss.n_total < 0 ? 0 : ss.n_total, n_max = ss.n_max < 0 ? 0 : ss.n_max;
    const size_t S = ss.S < 1 ? 1 : (size_t)ss.S;
    if (max_frames < 1) max_frames = 1;
    const int sms = (h && h->sms > 0) ? h->sms : 148;
    // one full wave of the detector kernel (24 one-warp CTAs per SM) over all streams together
    int64_t seg_len = (n_total + (int64_t)sms * 24 - 1) / ((int64_t)sms * 24);
    // (a segment re-runs OFDM_PEAK_WARM samples of warm-up: 16 384 keeps a short capture -- a feed_stream pass, one
    // stream of a batch -- on a few thousand warps at 2.5x the detector work, instead of a few hundred at 1.4x)
    if (seg_len < 16384) seg_len = 16384;
    seg_len = (seg_len + 31) / 32 * 32;
    ws->seg_len = seg_len;
    ws->n_seg = n_max > 0 ? (n_max + seg_len - 1) / seg_len : 0;
    ws->seg_cap = seg_len >> OFDM_SEG_CAP_SHIFT;
    ws->max_frames = max_frames;
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t sz) { size_t o = off; off = align_up(off + sz, 256); return b ? (void*)(b + o) : nullptr; };
    // fixed-size part first, so that the pointers do not depend on the stream lengths
    ws->first_nan = (int64_t*)take(sizeof(int64_t) * S);
    ws->first_ok = (int32_t*)take(sizeof(int32_t) * S);
    ws->live_overflow = (int32_t*)take(sizeof(int32_t) * S);
    ws->nco_init = (double*)take(sizeof(double) * S);
    ws->n_nco = (int32_t*)take(sizeof(int32_t) * S);
    ws->nco_sens = (double*)take(sizeof(double) * S);
    ws->plan_hdr = (int32_t*)take(4 * sizeof(int32_t) * S);
    ws->plan_blk_d = (double*)take(sizeof(double) * 1024 * S);
    ws->plan_blk_i = (int64_t*)take(sizeof(int64_t) * 1024 * S);
    ws->phi0 = (double*)take(sizeof(double) * max_frames * S);
    ws->step = (double*)take(sizeof(double) * max_frames * S);
    ws->vbase = (int64_t*)take(sizeof(int64_t) * ((size_t)max_frames + 1) * S);
    ws->sess_nvec = (int32_t*)take(sizeof(int32_t) * max_frames * S);
    ws->next_frame = (int32_t*)take(sizeof(int32_t) * max_frames * S);
    ws->exit_frame = (int32_t*)take(sizeof(int32_t) * max_frames * S);
    ws->seg_count = (int32_t*)take(sizeof(int32_t) * (size_t)(ws->n_seg + 1) * S);
    ws->seg_off = (int32_t*)take(sizeof(int32_t) * (size_t)(ws->n_seg + 1) * S);
    ws->seg_trig = (int64_t*)take(sizeof(int64_t) * (size_t)(ws->n_seg * ws->seg_cap) * S);
    ws->nco_idx = (int64_t*)take(sizeof(int64_t) * max_frames * S);
    ws->nco_ang = (float*)take(sizeof(float) * max_frames * S);
    // every frame is a flagged vector plus at most (distance to the next trigger) / L data vectors
    ws->eq_stride = (h ? n_max / h->L : 0) + max_frames + 2;
    ws->eq = (float2*)take(sizeof(float2) * (size_t)ws->eq_stride * (size_t)(h ? h->occ : 0) * S);
    ws->mf = (float*)take(sizeof(float) * (size_t)n_total);
    ws->y = (float2*)take(sizeof(float2) * (size_t)n_total);
    *need = off;
    if (b && bytes < off) return OFDM_E_NOMEM;
    return OFDM_OK;
}

// ---------------------------------------------------------------------------------------------
// K_RX1: gr.fft_filter_ccc(1, firdes.low_pass(...)) as overlap-save on the shared FFT core.
// Block b: forward FFT of x[b*V-(ntaps-1) .. +NOS), multiply by H/NOS in registers, inverse FFT with the
// reversed radix order (so the forward pass's register slots are exactly the inverse pass's inputs),
// keep the last V = NOS-(ntaps-1) points.  x is read once (+8% halo), y written once.
// ---------------------------------------------------------------------------------------------
struct FiltParams {
    const float2* x;
    float2* y;
    const int64_t* soff;       // stream offsets (nullptr: one stream of n samples)
    int64_t n;
    int V, hist;
    const float2* tw;
    const float2* H;
};

struct SmemRawIn {
    const float2* p;
    HDM float2 operator()(int i, int) const { return p[i]; }
};

// The two shared buffers of a block alternate roles: while the last inverse pass of block b streams its results
// out of one of them, the input of block b+1 is already landing in the other (cp.async, zero-filled outside the
// stream), so no pass waits on a global load.
template <int NOS, int G>
__global__ void __launch_bounds__(G * (NOS / FftPlan<NOS>::E), 512 / (G * (NOS / FftPlan<NOS>::E))) chan_filter_kernel(const FiltParams p) {
    using P = FftPlan<NOS>;
    constexpr int E = P::E;
    constexpr int T = NOS / E;
    constexpr int SB = (fft_smem_elems<NOS>() + 1) & ~1;      // even: both buffers stay 16-byte aligned
    constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2];
    static_assert(P::NP == 3 && R2 == E && R1 == E, "overlap-save plan must end with full-width radices");
    extern __shared__ __align__(16) float2 smem[];
    const int g = threadIdx.x / T;
    const int tid = threadIdx.x - g * T;
    float2* raw = smem + (size_t)g * 2 * SB;                  // holds the raw input of the current block
    float2* oth = raw + SB;
    int64_t s_a, s_n;
    stream_span(p.soff, blockIdx.y, p.n, s_a, s_n);           // this CTA's stream: zero history in front of it
    const float2* __restrict__ px = p.x + s_a;
    float2* __restrict__ py = p.y + s_a;
    const int64_t nblk = (s_n + p.V - 1) / p.V;
    const bool al16 = (((uintptr_t)px) & 15) == 0 && ((p.V | p.hist) & 1) == 0;
    auto prefetch = [&](float2* dst, int64_t blk) {
        if (blk < nblk) {
            const int64_t in0 = blk * p.V - p.hist;
            if (al16 && in0 >= 0 && in0 + NOS <= s_n) {
                // interior block (all but the first and the last few of a stream): no bounds to check
                const float2* src = px + in0;
#pragma unroll
                for (int i = 0; i < E / 2; ++i) {
                    const int idx = 2 * (tid + i * T);
                    cp_async16(dst + idx, src + idx, 16);
                }
            } else if (al16) {
#pragma unroll
                for (int i = 0; i < E / 2; ++i) {
                    const int idx = 2 * (tid + i * T);
                    const int64_t gi = in0 + idx;              // even, so a pair never straddles sample 0
                    int bytes = 0;
                    if (gi >= 0 && gi < s_n) bytes = (gi + 1 < s_n) ? 16 : 8;
                    cp_async16(dst + idx, px + (bytes ? gi : 0), bytes);
                }
            } else {
#pragma unroll
                for (int i = 0; i < E; ++i) {
                    const int idx = tid + i * T;
                    const int64_t gi = in0 + idx;
                    const bool in = gi >= 0 && gi < s_n;
                    cp_async8(dst + idx, px + (in ? gi : 0), in ? 8 : 0);
                }
            }
        }
        cp_async_commit();
    };
    prefetch(raw, (int64_t)blockIdx.x * G + g);
    for (int64_t base = (int64_t)blockIdx.x * G; base < nblk; base += (int64_t)gridDim.x * G) {
        const int64_t blk = base + g;
        const bool active = blk < nblk;
        float2 regs[E];
        auto mulH = [&](int idx, float2 v, int slot) { regs[slot] = cmul(v, LDG(p.H + idx)); };
        auto fromRegs = [&](int, int slot) -> float2 { return regs[slot]; };
        auto st = [&](int idx, float2 v, int) {
            if (idx >= p.hist) {
                int64_t o = blk * p.V + (idx - p.hist);
                if (o < s_n) py[o] = v;
            }
        };
        // interior block: every output lands inside the stream
        float2* const yb = py + (blk * p.V - p.hist);
        auto st_in = [&](int idx, float2 v, int) { if (idx >= p.hist) yb[idx] = v; };
        const bool inner = active && (blk + 1) * p.V <= s_n;
        cp_async_wait_all();
        __syncthreads();
        if (active) fft_pass<NOS, R0, 1, -1>(tid, p.tw, SmemRawIn{raw}, SmemOut{oth});
        __syncthreads();
        if (active) fft_pass<NOS, R1, R0, -1>(tid, p.tw, SmemIn{oth}, SmemOut{raw});
        __syncthreads();
        if (active) fft_pass<NOS, R2, R0 * R1, -1>(tid, p.tw, SmemIn{raw}, mulH);
        // inverse, radix order reversed: R2, R1, R0
        if (active) fft_pass<NOS, R2, 1, 1>(tid, p.tw, fromRegs, SmemOut{oth});
        __syncthreads();
        if (active) fft_pass<NOS, R1, R2, 1>(tid, p.tw, SmemIn{oth}, SmemOut{raw});
        __syncthreads();
        prefetch(oth, blk + (int64_t)gridDim.x * G);
        if (inner) fft_pass<NOS, R0, R2 * R1, 1>(tid, p.tw, SmemIn{raw}, st_in);
        else if (active) fft_pass<NOS, R0, R2 * R1, 1>(tid, p.tw, SmemIn{raw}, st);
        float2* t = raw; raw = oth; oth = t;
    }
    cp_async_wait_all();
}

// Warp-per-block overlap-save at NOS = 1024 = 32 x 32: one warp owns a block, 32 points per lane.  The first forward
// pass takes its points straight from global memory (stride-1 across the lanes), the last inverse pass stores y the
// same way; two radix-32 passes each way leave ONE shared-memory exchange per transform (2 per block of V outputs
// against 5 per block in the 3-pass kernel above) and no block barrier at all.
// HR: rows of 32 outputs discarded in front of a block (32 * HR >= ntaps - 1, a compile-time row predicate in the last
// pass); with 128-byte aligned streams every global row a warp loads or stores is two full lines.
template <int HR>
__global__ void __launch_bounds__(32, 16) chan_filter_warp_kernel(const FiltParams p) {
    constexpr int NOS = 1024, HIST = 32 * HR, V = NOS - HIST;
    using P = FftPlanW1024;
    __shared__ __align__(16) float2 buf[FFT_PAD32(NOS) + 2];
    const int lane = threadIdx.x;
    int64_t s_a, s_n;
    stream_span(p.soff, blockIdx.y, p.n, s_a, s_n);
    const float2* __restrict__ px = p.x + s_a;
    float2* __restrict__ py = p.y + s_a;
    const int64_t nblk = (s_n + V - 1) / V;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t in0 = blk * V - HIST;
        const bool inner = in0 >= 0 && in0 + NOS <= s_n;
        float2 regs[32];
        auto mulH = [&](int idx, float2 v, int slot) { regs[slot] = cmul(v, LDG(p.H + idx)); };
        auto fromRegs = [&](int, int slot) -> float2 { return regs[slot]; };
        const float2* src = px + in0;
        auto ld_in = [&](int idx, int) -> float2 { return src[idx]; };
        auto ld = [&](int idx, int) -> float2 {
            const int64_t gi = in0 + idx;
            return (gi >= 0 && gi < s_n) ? px[gi] : make_float2(0.f, 0.f);
        };
        float2* const yb = py + in0;
        auto st_in = [&](int idx, float2 v, int slot) { if (slot >= HR) yb[idx] = v; };
        auto st = [&](int idx, float2 v, int slot) { if (slot >= HR && in0 + idx < s_n) yb[idx] = v; };
        if (inner) fft_pass<NOS, 32, 1, -1, decltype(ld_in), SmemOut32, false, P>(lane, p.tw, ld_in, SmemOut32{buf});
        else fft_pass<NOS, 32, 1, -1, decltype(ld), SmemOut32, false, P>(lane, p.tw, ld, SmemOut32{buf});
        __syncwarp();
        fft_pass<NOS, 32, 32, -1, SmemIn32, decltype(mulH), false, P>(lane, p.tw, SmemIn32{buf}, mulH);
        __syncwarp();
        fft_pass<NOS, 32, 1, 1, decltype(fromRegs), SmemOut32, false, P>(lane, p.tw, fromRegs, SmemOut32{buf});
        __syncwarp();
        if (inner) fft_pass<NOS, 32, 32, 1, SmemIn32, decltype(st_in), false, P>(lane, p.tw, SmemIn32{buf}, st_in);
        else fft_pass<NOS, 32, 32, 1, SmemIn32, decltype(st), false, P>(lane, p.tw, SmemIn32{buf}, st);
        __syncwarp();
    }
}

template <int HR>
static int launch_filter_warp(ofdm_handle* h, const FiltParams& p, int64_t n_max, int S, cudaStream_t st) {
    const int64_t want = (n_max + (1024 - 32 * HR) - 1) / (1024 - 32 * HR);
    const int64_t cap = ((int64_t)h->sms * 64 + S - 1) / S;      // 16 one-warp CTAs per SM, four waves
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    chan_filter_warp_kernel<HR><<<dim3(grid, S), 32, 0, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

template <int NOS, int G>
static int launch_filter_n(ofdm_handle* h, const FiltParams& p, int64_t nblk_max, int S, cudaStream_t st) {
    constexpr int T = NOS / FftPlan<NOS>::E;
    constexpr int SB = (fft_smem_elems<NOS>() + 1) & ~1;
    size_t smem = ((size_t)G * 2 * SB) * sizeof(float2);
    OFDM_SET_MAX_SMEM((chan_filter_kernel<NOS, G>), smem, h->device);
    const int sms = h->sms;
    int64_t want = (nblk_max + G - 1) / G;
    int64_t cap = ((int64_t)sms * 16 + S - 1) / S;            // resident CTAs are shared by the streams
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    chan_filter_kernel<NOS, G><<<dim3(grid, S), G * T, smem, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

int launch_chan_filter(ofdm_handle* h, const float2* x, const StreamSet& ss, float2* y, cudaStream_t st) {
    if (ss.n_max <= 0) return OFDM_OK;
    FiltParams p;
    p.x = x; p.y = y; p.soff = ss.off; p.n = ss.n_max; p.hist = h->ntaps - 1; p.V = h->NOS - p.hist;
    p.tw = h->d_tw_os; p.H = h->d_Hos;
    const int64_t nblk_max = (ss.n_max + p.V - 1) / p.V;
    if (h->NOS == 1024) {                                                             // one block per warp
        if (p.hist <= 96) return launch_filter_warp<3>(h, p, ss.n_max, ss.S, st);
        if (p.hist <= 160) return launch_filter_warp<5>(h, p, ss.n_max, ss.S, st);
        return launch_filter_warp<8>(h, p, ss.n_max, ss.S, st);
    }
    if (h->NOS == 2048) return launch_filter_n<2048, 1>(h, p, nblk_max, ss.S, st);   // one block per CTA: 4 independent CTAs per SM
    if (h->NOS == 4096) return launch_filter_n<4096, 1>(h, p, nblk_max, ss.S, st);
    ofdm_set_error("chan_filter: unsupported overlap-save size %d", h->NOS);
    return OFDM_E_INVAL;
}

// gr.fir_filter_ccc with `ntaps` complex taps whose overlap-save response is H (rx_sync_alt.cu): the same kernel
int launch_xcorr(ofdm_handle* h, const float2* x, int64_t n, const float2* H, int nos, int ntaps, float2* out, cudaStream_t st) {
    if (n <= 0) return OFDM_OK;
    if (nos != 2048 && nos != 4096) { ofdm_set_error("xcorr: unsupported overlap-save size %d", nos); return OFDM_E_INVAL; }
    float2*& slot = h->d_tw_os_alt[nos == 4096];               // twiddles of the 3-pass kernel at this size, made on first use
    const float2* tw = (nos == h->NOS) ? h->d_tw_os : slot;
    if (!tw) {
        std::vector<float2> t((size_t)fft_twiddle_elems(nos), make_float2(0.f, 0.f));
        fft_fill_twiddles_n(nos, t.data());
        OFDM_CUDA_CHECK(cudaMalloc((void**)&slot, sizeof(float2) * t.size()));
        OFDM_CUDA_CHECK(cudaMemcpy(slot, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice));
        tw = slot;
    }
    FiltParams p;
    p.x = x; p.y = out; p.soff = nullptr; p.n = n; p.hist = ntaps - 1; p.V = nos - p.hist; p.tw = tw; p.H = H;
    const int64_t nblk = (n + p.V - 1) / p.V;
    if (nos == 2048) return launch_filter_n<2048, 1>(h, p, nblk, 1, st);
    if (nos == 4096) return launch_filter_n<4096, 1>(h, p, nblk, 1, st);
    ofdm_set_error("xcorr: unsupported overlap-save size %d", nos);
    return OFDM_E_INVAL;
}

// ---------------------------------------------------------------------------------------------
// K_RX2a: Schmidl-Cox timing metric (upstream ofdm_sync_pn up to add_const_ff(-1); A.6).
// The three N/2-wide moving sums (the reference runs them as brute-force FIRs) accumulate the float32
// products in float64 and round once to float32 -- the oracle's precision policy; the cp-wide average of
// the (non-negative) metric is a difference of tile-local float64 prefix sums.
// ---------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 512;

// Exclusive block scan.  Exclusive values are taken by shuffling the inclusive ones (never as
// "inclusive - own"), so a NaN/inf element only affects the elements after it.
__device__ __forceinline__ double block_excl_scan_f64(double v, double* s_warp, double* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    double exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 0.0;
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        double ti = lane < (SM_THREADS / 32) ? s_warp[lane] : 0.0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double o = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += o;
        }
        double te = __shfl_up_sync(0xffffffffu, ti, 1);
        if (lane == 0) te = 0.0;
        if (lane < SM_THREADS / 32) s_warp[lane] = te;         // exclusive warp offsets
        if (lane == SM_THREADS / 32 - 1) s_warp[32] = ti;      // block total
    }
    __syncthreads();
    double r = s_warp[w] + exc;
    if (total) *total = s_warp[32];
    __syncthreads();
    return r;
}

// Exclusive scan of one value per thread inside segments of G consecutive threads (G a power of two),
// forwards (sum of the segment's earlier threads) or backwards (later threads).  Only additions.
template <bool FWD, int G>
__device__ __forceinline__ double seg_excl_scan(double v, double* s_wt) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int width = G < 32 ? G : 32;
    const int pos = lane & (width - 1);
    double inc = v;
#pragma unroll
    for (int d = 1; d < width; d <<= 1) {
        const double o = FWD ? __shfl_up_sync(0xffffffffu, inc, d, width) : __shfl_down_sync(0xffffffffu, inc, d, width);
        if (FWD ? (pos >= d) : (pos + d < width)) inc += o;
    }
    double exc = FWD ? __shfl_up_sync(0xffffffffu, inc, 1, width) : __shfl_down_sync(0xffffffffu, inc, 1, width);
    if (FWD ? (pos == 0) : (pos == width - 1)) exc = 0.0;
    if (G > 32) {
        if (lane == (FWD ? 31 : 0)) s_wt[w] = inc;                        // warp total
        __syncthreads();
        constexpr int wps = G >> 5;
        const int wseg = w & (wps - 1), w0 = w - wseg;
        if (FWD) { for (int j = 0; j < wseg; ++j) exc += s_wt[w0 + j]; }
        else     { for (int j = wseg + 1; j < wps; ++j) exc += s_wt[w0 + j]; }
        __syncthreads();
    }
    return exc;
}

template <int K, int G>
__global__ void __launch_bounds__(SM_THREADS, (K <= 8 ? 2 : 1)) sync_metric_kernel(const float2* __restrict__ y, int64_t n, int cp,
                                                                 float tapf, float* __restrict__ mf,
                                                                 int64_t* __restrict__ first_nan) {
    constexpr int C = SM_THREADS * K;
#define PADK(e) ((e) + (e) / K)
    extern __shared__ double S[];                  // PADK(C) doubles, later reused as float staging
    __shared__ double s_warp[33];
    constexpr int h = G * K;                       // N/2
    const int T = C - cp - h;                      // outputs per tile
    const int64_t t0 = (int64_t)blockIdx.x * T;
    const int64_t a = t0 - cp - h;                 // global index of tile element 0
    const int tid = threadIdx.x;
    const int e0 = tid * K;

    float cre[K], cim[K], en[K];
    {
        float2 yv[K], yd[K];
        const int64_t u0 = a + e0;
        // interior tiles: 16-byte vector loads, no bounds checks (block-uniform branch)
        const bool fast = (a - h >= 0) && (a + C <= n) && ((((uintptr_t)(y + a)) & 15) == 0);
        if (fast) {
            const float4* pv = (const float4*)(y + u0);
            const float4* pd = (const float4*)(y + u0 - h);
#pragma unroll
            for (int i = 0; i < K / 2; ++i) {
                const float4 q = __ldg(pv + i), r = __ldg(pd + i);
                yv[2 * i] = make_float2(q.x, q.y); yv[2 * i + 1] = make_float2(q.z, q.w);
                yd[2 * i] = make_float2(r.x, r.y); yd[2 * i + 1] = make_float2(r.z, r.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const int64_t u = u0 + i, ud = u - h;
                yv[i] = (u >= 0 && u < n) ? LDG(y + u) : make_float2(0.f, 0.f);
                yd[i] = (ud >= 0 && ud < n) ? LDG(y + ud) : make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float2 c = cmulc_x(yv[i], yd[i]);   // y[n] * conj(y[n-N/2])
            cre[i] = c.x; cim[i] = c.y;
            en[i] = norm_x(yv[i]);
        }
    }
    float Pr[K], Pi[K], R[K];
    // Three moving sums of width h WITHOUT subtraction (van Herk / Gil-Werman): the tile is cut into blocks of
    // h samples (G = h/K threads); a window ending at e is  suffix(e-h+1 .. end of the previous block) +
    // prefix(start of e's block .. e).  Only terms inside the window are ever added, so an all-zero window
    // gives exactly 0 (-> 0/0 = NaN like the reference's FIR sums) and there is no cancellation noise when
    // the signal level drops.
    // A window ending at e0+i starts at (e0-h)+(i+1): elements i+1.. of thread tid-G, or (i = K-1) element 0 of
    // thread tid-G+1; that last window is a whole block (no tail) exactly when e0+K is a multiple of h.
    const bool has_prev = tid >= G;
    const bool whole_last = ((e0 + K) & (h - 1)) == 0;
    const int pb = e0 + tid;                                       // PADK(e0)
    const int qb = pb - (h + G);                                   // PADK(e0 - h)
#pragma unroll
    for (int arr = 0; arr < 3; ++arr) {
        const float* src = arr == 0 ? cre : (arr == 1 ? cim : en);
        float* dst = arr == 0 ? Pr : (arr == 1 ? Pi : R);
        double x[K], pre[K];
        double run = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) { x[i] = (double)src[i]; run += x[i]; pre[i] = run; }
        const double tot = run;
        const double fwd = seg_excl_scan<true, G>(tot, s_warp);
        const double bwd = seg_excl_scan<false, G>(tot, s_warp);
        run = bwd;
#pragma unroll
        for (int i = K - 1; i >= 0; --i) { run += x[i]; S[pb + i] = run; }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < K; ++i) {
            double tail = 0.0;
            if (has_prev) {
                if (i < K - 1) tail = S[qb + i + 1];
                else if (!whole_last) tail = S[qb + K + 1];
            }
            dst[i] = (float)(tail + (pre[i] + fwd));
        }
        __syncthreads();
    }
    // normalised metric, then the cp-wide average
    float Mt[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int e = e0 + i;
        const int64_t m = a + e;
        float num = fadd_rn(fmul_rn(Pr[i], Pr[i]), fmul_rn(Pi[i], Pi[i]));
        float den = fmul_rn(R[i], R[i]);
        float q = fdiv_rn(num, den);                // 0/0 -> NaN exactly like divide_ff
        Mt[i] = (e > h && m >= 0) ? q : 0.f;        // below: window not inside the tile / before the stream
    }
    {
        double loc[K];
        double run = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) { run += (double)Mt[i]; loc[i] = run; }
        const double base = block_excl_scan_f64(run, s_warp, nullptr);
#pragma unroll
        for (int i = 0; i < K; ++i) { loc[i] += base; S[PADK(e0 + i)] = loc[i]; }
        __syncthreads();
        float out[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int e = e0 + i - cp;
            const double prev = e >= 0 ? S[PADK(e)] : 0.0;
            float s = (float)((loc[i] - prev) * (double)tapf);
            out[i] = fadd_rn(s, -1.0f);
        }
        __syncthreads();
        float* stage = (float*)S;                   // coalesced write-out through shared memory
#pragma unroll
        for (int i = 0; i < K; ++i) stage[e0 + i] = out[i];
        __syncthreads();
        for (int o = tid; o < T; o += SM_THREADS) {
            const int64_t gi = t0 + o;
            if (gi < n) {
                float v = stage[o + cp + h];
                mf[gi] = v;
                if (v != v) atomicMin((unsigned long long*)first_nan, (unsigned long long)gi);
            }
        }
    }
#undef PADK
}

__global__ void init_i64_kernel(int64_t* p, int64_t v) { *p = v; }

template <int K, int G>
static int launch_metric_kg(ofdm_handle* h, const float2* y, int64_t n, float* mf, int64_t* first_nan, cudaStream_t st) {
    const int need = h->cp + h->N / 2;
    const float tapf = (float)(1.0 / (double)h->cp);
    const int T = SM_THREADS * K - need;
    if (T < 256) { ofdm_set_error("sync_metric: cp_length too large for the tile"); return OFDM_E_INVAL; }
    size_t smem = sizeof(double) * (size_t)(SM_THREADS * K + SM_THREADS + 8);
    OFDM_SET_MAX_SMEM((sync_metric_kernel<K, G>), smem, h->device);
    sync_metric_kernel<K, G><<<(unsigned)((n + T - 1) / T), SM_THREADS, smem, st>>>(y, n, h->cp, tapf, mf, first_nan);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

int launch_sync_metric(ofdm_handle* h, const float2* y, int64_t n, float* mf, int64_t* first_nan, cudaStream_t st) {
    init_i64_kernel<<<1, 1, 0, st>>>(first_nan, LLONG_MAX);
    OFDM_LAUNCH_CHECK();
    if (n <= 0) return OFDM_OK;
    switch (h->N) {                                   // G = (N/2) / K threads per N/2-sample block
        case 64:   return launch_metric_kg<8, 4>(h, y, n, mf, first_nan, st);
        case 128:  return launch_metric_kg<8, 8>(h, y, n, mf, first_nan, st);
        case 256:  return launch_metric_kg<8, 16>(h, y, n, mf, first_nan, st);
        case 512:  return launch_metric_kg<8, 32>(h, y, n, mf, first_nan, st);
        case 1024: return launch_metric_kg<8, 64>(h, y, n, mf, first_nan, st);
        case 2048: return launch_metric_kg<16, 64>(h, y, n, mf, first_nan, st);
        case 4096: return launch_metric_kg<16, 128>(h, y, n, mf, first_nan, st);
    }
    ofdm_set_error("sync_metric: unsupported fft_length %d", h->N);
    return OFDM_E_INVAL;
}

// ---------------------------------------------------------------------------------------------
// K_RX2b: gr.peak_detector_fb(0.20, 0.20, 30, 0.001) (A.7).  Every sample updates the IIR average
// exactly once, so avg is a pure linear recurrence of mf; one warp walks one contiguous segment,
// 32 samples per step: warp scan for the recurrence (float64 state like the oracle), a ballot for
// `mf > 0.2*avg_prev`, and a warp-uniform state machine over the ballot for the run / arg-max logic.
// A segment starts OFDM_PEAK_WARM samples early so that the average has converged (0.999^24576 ~ 2e-11)
// and keeps going past its end until an open run closes; a run belongs to the segment it started in.
// ---------------------------------------------------------------------------------------------
struct PeakParams {
    const float* mf;
    int64_t n;
    const int64_t* first_nan;
    int64_t seg_len, n_seg;
    int seg_cap;
    int32_t* seg_count;
    int64_t* seg_trig;
    uint32_t* status;
};

constexpr int PK = 8;                        // consecutive samples per lane and step (256 per warp step)

__global__ void __launch_bounds__(128) peak_detect_kernel(const PeakParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= p.n_seg) return;
    const int64_t s0 = seg * p.seg_len;
    const int64_t s1 = (s0 + p.seg_len < p.n) ? s0 + p.seg_len : p.n;
    int64_t w0 = s0 - OFDM_PEAK_WARM;
    if (w0 < 0) w0 = 0;
    w0 &= ~(int64_t)(32 * PK - 1);
    const int64_t fnan = *p.first_nan;
    const bool vec_ok = (((uintptr_t)p.mf) & 15) == 0;
    int count = 0;
    if (fnan >= w0) {                               // else: average already poisoned, nothing can fire (C.1)
        const double a1 = (double)0.001f;
        const double a2 = 1.0 - a1;
        double a2k = 1.0;                           // a2^PK: multiplier of one lane's block
#pragma unroll
        for (int k = 0; k < PK; ++k) a2k *= a2;
        double pw[5];                               // (a2^PK)^(2^k) for the scan steps
        pw[0] = a2k;
#pragma unroll
        for (int k = 1; k < 5; ++k) pw[k] = pw[k - 1] * pw[k - 1];
        const double p32 = pw[4] * pw[4];           // (a2^PK)^32
        double plane = 1.0;                         // (a2^PK)^lane: weight of the carry at the start of this lane
        for (int k = 0; k < lane; ++k) plane *= a2k;
        double carry = 0.0;                         // avg after the last consumed sample
        int state = 0;
        float peak = -INFINITY;
        int64_t ind = 0, run_start = 0;
        for (int64_t i0 = w0; i0 < p.n; i0 += 32 * PK) {
            if (i0 >= s1 && state == 0) break;
            const int64_t b0 = i0 + (int64_t)lane * PK;
            float v[PK];
            if (vec_ok && i0 + 32 * PK <= p.n) {
                const float4 q0 = __ldg((const float4*)(p.mf + b0));
                const float4 q1 = __ldg((const float4*)(p.mf + b0) + 1);
                v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
                v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
            } else {
#pragma unroll
                for (int i = 0; i < PK; ++i) v[i] = (b0 + i < p.n) ? p.mf[b0 + i] : 0.f;
            }
            // lane-local recurrence from a zero state, then the affine scan across lanes
            double loc = 0.0;
#pragma unroll
            for (int i = 0; i < PK; ++i) loc = a2 * loc + a1 * (double)v[i];
            double b = loc;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const double o = __shfl_up_sync(0xffffffffu, b, 1 << k);
                if (lane >= (1 << k)) b = b + pw[k] * o;
            }
            double prev = __shfl_up_sync(0xffffffffu, b, 1);      // state after the previous lane's block (zero carry)
            if (lane == 0) prev = 0.0;
            prev = prev + plane * carry;                          // avg just before this lane's first sample
            carry = __shfl_sync(0xffffffffu, b, 31) + p32 * carry;      // state after the 256th sample
            unsigned mk = 0;
#pragma unroll
            for (int i = 0; i < PK; ++i) {
                const float thr = fmul_rn((float)prev, 0.2f);
                if (b0 + i < p.n && v[i] > thr) mk |= 1u << i;
                prev = a2 * prev + a1 * (double)v[i];
            }
            const unsigned any = __ballot_sync(0xffffffffu, mk != 0);
            if (state == 0 && any == 0) continue;
            // slow path (a run is open or starts in this step): walk the 256 samples in order
            for (int l = (state == 0 ? __ffs(any) - 1 : 0); l < 32; ++l) {
                const unsigned m = __shfl_sync(0xffffffffu, mk, l);
                if (state == 0 && m == 0) continue;
                const int64_t base = i0 + (int64_t)l * PK;
#pragma unroll
                for (int i = 0; i < PK; ++i) {
                    const float vi = __shfl_sync(0xffffffffu, v[i], l);
                    const int64_t idx = base + i;
                    if (idx >= p.n) break;                        // end of stream: an open run emits nothing
                    const bool bit = (m >> i) & 1u;
                    if (state == 0) {
                        if (bit) { state = 1; peak = vi; ind = idx; run_start = idx; }
                    } else if (vi > peak) {
                        peak = vi; ind = idx;                     // also keeps the run alive when !bit
                    } else if (!bit) {
                        if (run_start >= s0 && run_start < s1) {
                            if (count < p.seg_cap) {
                                if (lane == 0) p.seg_trig[seg * p.seg_cap + count] = ind;
                            } else if (lane == 0) {
                                atomicOr(p.status, OFDM_ST_SEG_OVERFLOW);
                            }
                            ++count;
                        }
                        state = 0;                                // the closing sample is consumed in state 0
                    }
                }
            }
        }
    }
    if (lane == 0) p.seg_count[seg] = count < p.seg_cap ? count : p.seg_cap;
}

// exclusive scan of the segment counts (one CTA), then gather + angle
__global__ void __launch_bounds__(1024) seg_scan_kernel(const int32_t* __restrict__ seg_count, int64_t seg_stride,
                                                        const int64_t* __restrict__ soff, int64_t n_single, int64_t seg_len,
                                                        int32_t* __restrict__ seg_off, int32_t* __restrict__ n_trig,
                                                        int max_frames, uint32_t* status) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t s_a, s_n;
    stream_span(soff, blockIdx.x, n_single, s_a, s_n);
    const int64_t n_seg = s_n > 0 ? (s_n + seg_len - 1) / seg_len : 0;      // segments of THIS stream
    seg_count += (int64_t)blockIdx.x * (seg_stride + 1);
    seg_off += (int64_t)blockIdx.x * (seg_stride + 1);
    n_trig += blockIdx.x;
    status += blockIdx.x;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_seg; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int v = i < n_seg ? seg_count[i] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        if (w == 0) {
            int t = s_w[lane];
            int ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int o = __shfl_up_sync(0xffffffffu, ti, d);
                if (lane >= d) ti += o;
            }
            s_w[lane] = ti - t;
        }
        __syncthreads();
        const int excl = s_carry + s_w[w] + inc - v;
        if (i < n_seg) seg_off[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int total = s_carry;
        if (total > max_frames) { total = max_frames; atomicOr(status, OFDM_ST_TRIG_OVERFLOW); }
        *n_trig = total;
    }
}

// one warp per trigger slot: copy the index, recompute P = sum_{k<N/2} y[t-k] conj(y[t-k-N/2]) (float64
// accumulation of float32 products, rounded to float32) and latch angle = atan2(Im P, Re P)
__global__ void __launch_bounds__(256) trig_gather_kernel(const float2* __restrict__ y, const int64_t* __restrict__ soff,
                                                          int64_t n_single, int64_t seg_len, int N,
                                                          const int32_t* __restrict__ seg_count,
                                                          const int32_t* __restrict__ seg_off,
                                                          const int64_t* __restrict__ seg_trig, int64_t seg_stride, int seg_cap,
                                                          int max_frames, int64_t* __restrict__ trig_idx,
                                                          float* __restrict__ trig_ang) {
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t s_a, n;
    stream_span(soff, blockIdx.y, n_single, s_a, n);
    const int64_t n_seg = n > 0 ? (n + seg_len - 1) / seg_len : 0;
    if (seg >= n_seg) return;
    y += s_a;
    seg_count += (int64_t)blockIdx.y * (seg_stride + 1);
    seg_off += (int64_t)blockIdx.y * (seg_stride + 1);
    seg_trig += (int64_t)blockIdx.y * seg_stride * seg_cap;
    trig_idx += (int64_t)blockIdx.y * max_frames;
    trig_ang += (int64_t)blockIdx.y * max_frames;
    const int cnt = seg_count[seg];
    const int off = seg_off[seg];
    const int h = N / 2;
    for (int c = 0; c < cnt; ++c) {
        const int dst = off + c;
        if (dst >= max_frames) break;
        const int64_t t = seg_trig[seg * seg_cap + c];
        double pr = 0.0, pi = 0.0;
        for (int k = lane; k < h; k += 32) {
            const int64_t u = t - k, ud = u - h;
            float2 v = (u >= 0 && u < n) ? LDG(y + u) : make_float2(0.f, 0.f);
            float2 d = (ud >= 0 && ud < n) ? LDG(y + ud) : make_float2(0.f, 0.f);
            float2 cc = cmulc_x(v, d);
            pr += (double)cc.x;
            pi += (double)cc.y;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            pr += __shfl_xor_sync(0xffffffffu, pr, d);
            pi += __shfl_xor_sync(0xffffffffu, pi, d);
        }
        if (lane == 0) {
            trig_idx[dst] = t;
            trig_ang[dst] = (float)atan2((double)(float)pi, (double)(float)pr);
        }
    }
}

int launch_trig_compact(ofdm_handle* h, const float2* y, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st) {
    seg_scan_kernel<<<ss.S, 1024, 0, st>>>(ws->seg_count, ws->n_seg, ss.off, ss.n_max, ws->seg_len, ws->seg_off, io->n_trig,
                                           io->max_frames, io->status);
    OFDM_LAUNCH_CHECK();
    trig_gather_kernel<<<dim3((unsigned)((ws->n_seg + 7) / 8), ss.S), 256, 0, st>>>(
        y, ss.off, ss.n_max, ws->seg_len, h->N, ws->seg_count, ws->seg_off, ws->seg_trig, ws->n_seg, (int)ws->seg_cap,
        io->max_frames, io->trig_idx, io->trig_ang);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

int launch_peak_detect(ofdm_handle* h, const float2* y, const float* mf, int64_t n, const int64_t* first_nan,
                       ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st) {
    OFDM_CUDA_CHECK(cudaMemsetAsync(io->status, 0, sizeof(uint32_t), st));
    if (first_nan != ws->first_nan)                  // the plan reads it from the workspace
        OFDM_CUDA_CHECK(cudaMemcpyAsync(ws->first_nan, first_nan, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    if (ws->n_seg == 0) {
        OFDM_CUDA_CHECK(cudaMemsetAsync(io->n_trig, 0, sizeof(int32_t), st));
        return OFDM_OK;
    }
    PeakParams p;
    p.mf = mf; p.n = n; p.first_nan = first_nan; p.seg_len = ws->seg_len; p.n_seg = ws->n_seg; p.seg_cap = (int)ws->seg_cap;
    p.seg_count = ws->seg_count; p.seg_trig = ws->seg_trig; p.status = io->status;
    const int wpb = 4;
    peak_detect_kernel<<<(unsigned)((ws->n_seg + wpb - 1) / wpb), wpb * 32, 0, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return launch_trig_compact(h, y, single_stream(n), io, ws, st);
}

// ---------------------------------------------------------------------------------------------
// K_RX3: gr.frequency_modulator_fc(-2/N) phase bookkeeping (A.8) and digital.ofdm_sampler (A.9) in closed
// form per trigger.  Two grid-wide launches: CTA-local prefixes + totals, then offsets (the divisions of the
// closed form were 0.3 ms on a single SM).
// ---------------------------------------------------------------------------------------------
struct PlanParams {
    int64_t n;
    const int64_t* soff;   // stream offsets (nullptr: one stream of n samples); every table below is per stream
    int N, L, max_frames;
    int32_t* n_trig;
    const int64_t* first_nan;
    const int64_t* trig_idx;
    const float* trig_ang;
    double* phi0;
    double* step;
    int32_t* first_ok;
    int32_t* n_frames;
    int64_t* frame_start;
    int32_t* frame_ndata;
    int64_t* vbase;
    int64_t* counters;
    const double* nco_init;// [1] phase step per sample before the first trigger
    int32_t* hdr;          // [4] scratch: K, first_ok, frames before the first call that cannot run
    double* blk_d;         // [nblk] CTA totals of the phase increments
    long long* blk_i;      // [nblk] CTA totals of the vector counts
    // the NCO's own event list (ofdm_sync_ml: the held angle changes at every detector peak, only some of which are timing
    // triggers); *n_nco < 0: the NCO follows the trigger list.  sens = the NCO sensitivity (-2/N; ofdm_sync_ml: -1/N)
    const int32_t* n_nco;
    const int64_t* nco_idx;
    const float* nco_ang;
    const double* sens;    // [1] written by the synchroniser stage
};

__device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// the tables of stream s (blockIdx.y), built field by field (not as a modified copy of the kernel parameter: see
// demod_kernel in rx_demod.cu)
__device__ __forceinline__ PlanParams plan_stream_view(const PlanParams& q, int s) {
    PlanParams p;
    int64_t a;
    stream_span(q.soff, s, q.n, a, p.n);
    const int64_t mf = q.max_frames;
    p.soff = q.soff; p.N = q.N; p.L = q.L; p.max_frames = q.max_frames;
    p.n_trig = q.n_trig + s; p.first_nan = q.first_nan + s; p.trig_idx = q.trig_idx + s * mf; p.trig_ang = q.trig_ang + s * mf;
    p.phi0 = q.phi0 + s * mf; p.step = q.step + s * mf; p.first_ok = q.first_ok + s; p.n_frames = q.n_frames + s;
    p.frame_start = q.frame_start + s * mf; p.frame_ndata = q.frame_ndata + s * mf; p.vbase = q.vbase + s * (mf + 1);
    p.counters = q.counters + s * 8; p.nco_init = q.nco_init + s; p.hdr = q.hdr + s * 4; p.blk_d = q.blk_d + s * 1024;
    p.blk_i = q.blk_i + s * 1024;
    p.n_nco = q.n_nco + s; p.nco_idx = q.nco_idx + s * mf; p.nco_ang = q.nco_ang + s * mf; p.sens = q.sens + s;
    return p;
}

// exclusive block scan of one value per thread (1024 threads); total = sum over the block.  s_w: [33] scratch.
template <typename T>
__device__ __forceinline__ T plan_block_scan(T v, T* s_w, T& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const T o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += o;
    }
    T ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = (T)0;
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        const T t = s_w[lane];
        T ti = t;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const T o = __shfl_up_sync(0xffffffffu, ti, s);
            if (lane >= s) ti += o;
        }
        T te = __shfl_up_sync(0xffffffffu, ti, 1);
        if (lane == 0) te = (T)0;
        s_w[lane] = te;
        if (lane == 31) s_w[32] = ti;
    }
    __syncthreads();
    ex += s_w[w];
    total = s_w[32];
    __syncthreads();
    return ex;
}

// P1 (grid-wide, one trigger per thread): per-trigger NCO step and phase increment, per-frame vector count, their
// CTA-local exclusive prefixes and the CTA totals; hdr = {K after the NaN cut, first_ok, first trigger whose call cannot run}.
__global__ void __launch_bounds__(1024) plan_local_kernel(const PlanParams p_all) {
    const PlanParams p = plan_stream_view(p_all, blockIdx.y);
    __shared__ double s_wd[33];
    __shared__ long long s_wi[33];
    __shared__ int s_K, s_first_ok, s_Kn;
    const int tid = threadIdx.x;
    const int64_t N = p.N, L = p.L, n = p.n;
    if (tid == 0) {
        int K = *p.n_trig;
        if (K > p.max_frames) K = p.max_frames;
        // nothing fires after the first NaN of the timing metric (C.1): a run that was open closes there with
        // its arg-max before it, so exactly the triggers at indices < first_nan survive
        const int64_t fn = *p.first_nan;
        int lo = 0, hi = K;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (p.trig_idx[mid] < fn) lo = mid + 1; else hi = mid;
        }
        K = lo;
        // (2) the sampler never looks at indices < N: first visible trigger (the indices ascend)
        lo = 0; hi = K;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (p.trig_idx[mid] < N) lo = mid + 1; else hi = mid;
        }
        int Kn = *p.n_nco;
        if (Kn < 0) Kn = K; else if (Kn > p.max_frames) Kn = p.max_frames;
        s_K = K; s_first_ok = lo; s_Kn = Kn;
        if (blockIdx.x == 0) { p.hdr[0] = K; p.hdr[1] = lo; p.hdr[3] = Kn; }
    }
    __syncthreads();
    const int K = s_K, first_ok = s_first_ok, Kn = s_Kn;
    const bool own = *p.n_nco >= 0;                        // the NCO has its own event list
    const int64_t* nidx = own ? p.nco_idx : p.trig_idx;
    const float* nang = own ? p.nco_ang : p.trig_ang;
    const int k = blockIdx.x * 1024 + tid;
    double d = 0.0;
    long long v = 0;
    if (k < Kn) {
        // (1) NCO: step_k = sens * ang_k ; phi0_k = sum_{j<k} step_j * (t_{j+1} - t_j) over the NCO's events
        const double stp = *p.sens * (double)nang[k];
        p.step[k] = stp;
        if (k + 1 < Kn) d = stp * (double)(nidx[k + 1] - nidx[k]);
    }
    if (k < K) {
        const int64_t t = p.trig_idx[k];
        const int64_t tn = (k + 1 < K) ? p.trig_idx[k + 1] : 0;
        if (k >= first_ok) {
            // (3) can the call that finds trigger k run?  (a call at read pointer pos needs pos+L+N < n)
            bool ok;
            if (k == first_ok) {
                const int64_t c = (t - N) / (L + 1);
                ok = c * (L + 1) + L + N < n;
            } else {
                const int64_t tp = p.trig_idx[k - 1];
                int64_t mm = ceil_div64(t - tp - 1, L);
                if (mm < 1) mm = 1;
                if (mm <= OFDM_SAMPLER_MAXDATA) {
                    ok = tp + 1 + mm * L < n;
                } else {
                    const int64_t q0 = tp + 1 + (int64_t)OFDM_SAMPLER_MAXDATA * L;   // pos + N of the first NO_SIG call
                    const int64_t c = (t - q0) / (L + 1);
                    ok = q0 + c * (L + 1) + L < n;
                }
            }
            if (!ok) atomicMin(&p.hdr[2], k - first_ok);     // later ones fail too (pos only grows)
            // (4) data vectors of the frame
            int64_t J = OFDM_SAMPLER_MAXDATA;
            if (k + 1 < K) {
                int64_t mm = ceil_div64(tn - t - 1, L);
                if (mm < 1) mm = 1;
                if (mm - 1 < J) J = mm - 1;
            }
            const int64_t room = (n - 2 - t >= 0) ? (n - 2 - t) / L : 0;
            if (room < J) J = room;
            if (J < 0) J = 0;
            p.frame_start[k - first_ok] = t - N + 1;
            p.frame_ndata[k - first_ok] = (int)J;
            v = 1 + J;
        }
    }
    double tot_d;
    long long tot_i;
    const double ex_d = plan_block_scan<double>(d, s_wd, tot_d);
    const long long ex_i = plan_block_scan<long long>(v, s_wi, tot_i);
    if (k < Kn) p.phi0[k] = ex_d;
    if (k < K && k >= first_ok) p.vbase[k - first_ok] = ex_i;
    if (tid == 0) { p.blk_d[blockIdx.x] = tot_d; p.blk_i[blockIdx.x] = tot_i; }
}

// P2 (same grid): every CTA scans the CTA totals (at most 1024 of them), adds its own offset; CTA 0 publishes the scalars.
__global__ void __launch_bounds__(1024) plan_offset_kernel(const PlanParams p_all, int nblk) {
    const PlanParams p = plan_stream_view(p_all, blockIdx.y);
    __shared__ double s_wd[33];
    __shared__ long long s_wi[33];
    __shared__ double s_off_d;
    __shared__ long long s_off_i;
    const int tid = threadIdx.x;
    const int K = p.hdr[0], first_ok = p.hdr[1], Kn = p.hdr[3];
    const int64_t* nidx = (*p.n_nco >= 0) ? p.nco_idx : p.trig_idx;
    int F = K - first_ok;
    if (p.hdr[2] < F) F = p.hdr[2];
    if (F < 0) F = 0;
    double tot_d;
    long long tot_i;
    const double ex_d = plan_block_scan<double>(tid < nblk ? p.blk_d[tid] : 0.0, s_wd, tot_d);
    const long long ex_i = plan_block_scan<long long>(tid < nblk ? p.blk_i[tid] : 0, s_wi, tot_i);
    if (tid == blockIdx.x) {
        // phase accumulated over samples 0 .. t0-1 (non-zero only behind ofdm_sync_fixed)
        s_off_d = ex_d + (Kn > 0 ? *p.nco_init * (double)nidx[0] : 0.0);
        s_off_i = ex_i;
    }
    __syncthreads();
    const int k = blockIdx.x * 1024 + tid;
    if (k < Kn) p.phi0[k] = s_off_d + p.phi0[k];
    if (k < K) {
        const int f = k - first_ok;
        if (f >= 0 && f <= F) {
            const long long vb = s_off_i + p.vbase[f];
            p.vbase[f] = vb;
            if (f == F) p.counters[6] = vb;         // F < K - first_ok: vectors of the F frames
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        if (F == K - first_ok) { p.vbase[F] = tot_i; p.counters[6] = tot_i; }
        *p.n_trig = K;
        *p.first_ok = first_ok;
        *p.n_frames = F;
        p.counters[0] = F;
        p.counters[4] = p.n;
        p.counters[5] = K;
    }
}

__global__ void plan_init_kernel(int32_t* hdr, int S) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) { hdr[4 * s] = 0; hdr[4 * s + 1] = 0; hdr[4 * s + 2] = INT_MAX; hdr[4 * s + 3] = 0; }
}

__global__ void nco_mode_kernel(int32_t* n_nco, double* sens, int S, int32_t n_value, double sens_value) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) { n_nco[s] = n_value; sens[s] = sens_value; }
}

// what the synchroniser stage leaves for the NCO: its sensitivity (x fft_length: -2 behind ofdm_sync_pn / pnac / fixed, -1
// behind ofdm_sync_ml, ofdm_receiver.py~:91,97,103,115) and whether it brings its own event list (n_nco >= 0 is written later)
int launch_nco_mode(ofdm_handle* h, RxWorkspace* ws, int S, double sens_times_n, cudaStream_t st) {
    nco_mode_kernel<<<(S + 255) / 256, 256, 0, st>>>(ws->n_nco, ws->nco_sens, S, -1, sens_times_n / (double)h->N);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

int launch_plan(ofdm_handle* h, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st) {
    PlanParams p;
    p.n_nco = ws->n_nco; p.nco_idx = ws->nco_idx; p.nco_ang = ws->nco_ang; p.sens = ws->nco_sens;
    p.n = ss.n_max; p.soff = ss.off; p.N = h->N; p.L = h->L; p.max_frames = io->max_frames; p.n_trig = io->n_trig; p.first_nan = ws->first_nan;
    p.trig_idx = io->trig_idx;
    p.trig_ang = io->trig_ang; p.phi0 = ws->phi0; p.step = ws->step; p.first_ok = ws->first_ok; p.n_frames = io->n_frames;
    p.frame_start = io->frame_start; p.frame_ndata = io->frame_ndata; p.vbase = ws->vbase; p.counters = io->counters;
    p.nco_init = ws->nco_init; p.hdr = ws->plan_hdr; p.blk_d = ws->plan_blk_d; p.blk_i = (long long*)ws->plan_blk_i;
    const int nblk = (io->max_frames + 1023) / 1024;
    if (nblk > 1024) {
        ofdm_set_error("plan: max_frames %d exceeds 1048576", io->max_frames);
        return OFDM_E_INVAL;
    }
    OFDM_CUDA_CHECK(cudaMemsetAsync(io->counters, 0, 8 * sizeof(int64_t) * ss.S, st));
    plan_init_kernel<<<(ss.S + 255) / 256, 256, 0, st>>>(p.hdr, ss.S);
    OFDM_LAUNCH_CHECK();
    plan_local_kernel<<<dim3(nblk, ss.S), 1024, 0, st>>>(p);
    OFDM_LAUNCH_CHECK();
    plan_offset_kernel<<<dim3(nblk, ss.S), 1024, 0, st>>>(p, nblk);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// ---------------------------------------------------------------------------------------------
// upstream ofdm_sync_fixed (ofdm_receiver.py~:108-119, "for testing only"): a trigger at the last sample of the
// first symbol of every nsymbols-symbol packet and a constant frequency-offset stream pi*freq_offset.
// ---------------------------------------------------------------------------------------------
__global__ void sync_fixed_kernel(int64_t n, int L, int64_t period, float ang, double init_step, int max_frames,
                                  int32_t* __restrict__ n_trig, int64_t* __restrict__ trig_idx,
                                  float* __restrict__ trig_ang, double* __restrict__ nco_init,
                                  int64_t* __restrict__ first_nan, uint32_t* __restrict__ status) {
    const int64_t count = (n > L - 1) ? (n - L) / period + 1 : 0;          // indices L-1 + k*period < n
    const int64_t kept = count < max_frames ? count : max_frames;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < kept; k += (int64_t)gridDim.x * blockDim.x) {
        trig_idx[k] = (int64_t)(L - 1) + k * period;
        trig_ang[k] = ang;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *n_trig = (int32_t)kept;
        *nco_init = init_step;
        *first_nan = LLONG_MAX;
        *status = count > max_frames ? OFDM_ST_TRIG_OVERFLOW : 0u;
    }
}

int launch_sync_fixed(ofdm_handle* h, int64_t n, int32_t nsymbols, float freq_offset, ofdm_rx_io* io, RxWorkspace* ws,
                      cudaStream_t st) {
    if (nsymbols < 1) {
        ofdm_set_error("sync_fixed: nsymbols %d < 1", nsymbols);
        return OFDM_E_INVAL;
    }
    const float ang = (float)(3.14159265358979323846 * (double)freq_offset);
    const double init_step = (-2.0 / (double)h->N) * (double)ang;
    int grid = (io->max_frames + 255) / 256;
    if (grid > h->sms * 4) grid = h->sms * 4;
    if (grid < 1) grid = 1;
    sync_fixed_kernel<<<grid, 256, 0, st>>>(n, h->L, (int64_t)nsymbols * h->L, ang, init_step, io->max_frames, io->n_trig,
                                            io->trig_idx, io->trig_ang, ws->nco_init, ws->first_nan, io->status);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}
