// Internal declarations shared by the translation units of libofdm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ofdm_b200.h"

#define OFDM_MAX_TAPS 512
#define OFDM_SAMPLER_TIMEOUT 1000       // digital_swig.py:4719-4720
#define OFDM_SAMPLER_MAXDATA (OFDM_SAMPLER_TIMEOUT + 1)   // `if (d_timeout-- == 0)`: the FRAME state emits timeout + 1 data vectors
#define OFDM_ACQ_MAX_SYMBOLS 1000       // upstream MAX_NUM_SYMBOLS
#define OFDM_MAX_SHIFT 4                // digital_swig.py:4320 (max_fft_shift_len)
#define OFDM_PEAK_WARM 24576            // samples of IIR warm-up before a detector segment
#define OFDM_CRC_SLICE (256 + 32 + 128)   // offset of the slicing tables T1, T2, T3 in d_crctab
#define OFDM_SEG_CAP_SHIFT 6            // detector segment slot list: seg_len >> 6 triggers

// One receive call serves S independent streams laid back to back in one sample buffer (S = 1: the classic
// single-stream call).  Stream s occupies samples [off[s], off[s+1]) of x / y / mf and starts from zero history
// (filter, window sums, detector average, NCO phase) exactly like a call of its own; every per-stream table sits at
// base + s * stride.  Kernels take the stream index from blockIdx.y.
struct StreamSet {
    int S;                   // number of streams
    const int64_t* off;      // device [S+1]; nullptr for a single stream [0, n_max)
    int64_t n_max;           // longest stream (host knowledge: grids and segment tables are sized by it)
    int64_t n_total;         // samples in all (off[S])
};
static inline StreamSet single_stream(int64_t n) { StreamSet ss; ss.S = 1; ss.off = nullptr; ss.n_max = n; ss.n_total = n; return ss; }
#ifdef __CUDACC__
// span of stream s: first sample a, length n
__device__ __forceinline__ void stream_span(const int64_t* off, int s, int64_t n_single, int64_t& a, int64_t& n) {
    if (off) { a = off[s]; n = off[s + 1] - a; } else { a = 0; n = n_single; }
}
#endif

struct ofdm_handle {
    int device;
    int sms;                 // multiprocessors of the device (grid sizing)
    int N, occ, cp, L, zl, M, nbits, ncar, ntaps, NOS, pkt_stride;
    float amp;
    uint64_t pad_seed;
    int tx_row_lo, tx_row_hi;   // IFFT input rows (of N/32 bins, ifftshift applied) [lo, hi) carry no data carrier
    // device tables
    float2* d_const;       // [M]
    int16_t* d_bin2car;    // [N]  FFT-vector index -> data-carrier ordinal, -1 if unused (A.3 mapper map)
    int16_t* d_sinkmap;    // [ncar] index into the occ-wide equalised vector (A.3 sink map)
    float* d_ks;           // [occ] known symbol with odd bins zeroed
    float* d_kd;           // [occ] |ks[i]-ks[i+2]|^2 on even i
    float2* d_tw;          // [N]   exp(-2*pi*j*i/N)
    float2* d_tw_w;        // N == 512 / 1024: twiddles of the warp plan (fft.cuh FftPlanW), else null
    float2* d_tw_os;       // [NOS]
    float2* d_Hos;         // [NOS] FFT of the channel taps / NOS
    float2* d_pre_time;    // [N+cp] time-domain preamble incl. CP, scaled by 1/sqrt(N)
    // known-symbol correlators of the "pnac" / "ml" synchronisers (created on first use, rx_sync_alt.cu)
    float2* d_Hks_half;    // [nos_ks_half] response of conj(ks0time[:N/2]) reversed
    float2* d_Hks_full;    // [nos_ks_full] response of conj(ks0time) reversed
    int nos_ks_half, nos_ks_full;
    float2* d_tw_os_alt[2];// twiddles of the 2048 / 4096-point overlap-save sizes the channel filter itself does not use
    float2* d_pre_freq;    // [N] the known symbol as the mapper-order (unshifted) vector   (options.log taps)
    float2* d_pre_ifft;    // [N] its unscaled IFFT
    uint8_t* d_mask;       // [4096] whitening mask
    uint32_t* d_crctab;    // [256] byte table, [32] + [128] shift factors (crc_warp_kernel), then 3 x [256] slicing tables T1..T3
    float h_taps[OFDM_MAX_TAPS];
    // square-grid constellations (qam64 / qam256): levels per axis (0 = brute-force slicer only), cell -> index table
    int grid_L;
    float grid_x0, grid_y0, grid_inv_dx, grid_inv_dy;
    uint8_t* d_grid;       // [grid_L * grid_L], row = y level
};

struct ofdm_sense_handle {
    int device;
    int sms;
    int N;
    float* d_win;          // [N] Blackman-Harris
    float2* d_tw;          // [N]
};

void ofdm_set_error(const char* fmt, ...);
#define OFDM_CUDA_CHECK(call)                                                               \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            ofdm_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return OFDM_E_CUDA;                                                             \
        }                                                                                   \
    } while (0)
#define OFDM_LAUNCH_CHECK() OFDM_CUDA_CHECK(cudaGetLastError())
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember the largest size set on each device
// (one process may hold handles on several GPUs).  Pass a template-id in parentheses.
#define OFDM_MAX_DEVICES 64
#define OFDM_SET_MAX_SMEM(func, bytes, device)                                                                   \
    do {                                                                                                          \
        static size_t done__[OFDM_MAX_DEVICES] = {0};                                                             \
        const int d__ = ((device) >= 0 && (device) < OFDM_MAX_DEVICES) ? (device) : 0;                            \
        if ((size_t)(bytes) > done__[d__]) {                                                                      \
            OFDM_CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            done__[d__] = (size_t)(bytes);                                                                        \
        }                                                                                                         \
    } while (0)

// workspace carve-out (rx_front.cu); "per stream" arrays hold S consecutive copies (stride = the size given)
struct RxWorkspace {
    float2* y;             // [n_total]
    float* mf;             // [n_total]
    int64_t* first_nan;    // per stream [1]
    int32_t* seg_count;    // per stream [n_seg + 1]
    int64_t* seg_trig;     // per stream [n_seg * seg_cap]
    double* phi0;          // per stream [max_frames] NCO phase just before each trigger takes effect
    double* step;          // per stream [max_frames] NCO phase step per sample after each trigger
    int32_t* first_ok;     // per stream [1] index of the first trigger the sampler can see
    double* nco_init;      // per stream [1] NCO phase step per sample before the first trigger (0 behind ofdm_sync_pn)
    int32_t* plan_hdr;     // per stream [4] scratch of the two-launch plan
    double* plan_blk_d;    // per stream [1024]
    int64_t* plan_blk_i;   // per stream [1024]
    int32_t* live_overflow;// per stream [1] set when the liveness fast path hands over to the general walk
    int64_t* vbase;        // per stream [max_frames + 1] position of each frame's preamble vector in the vector stream
    int32_t* sess_nvec;    // per stream [max_frames] vectors consumed by a sink session started at this frame
    int32_t* next_frame;   // per stream [max_frames] scratch of the liveness walk
    int32_t* exit_frame;   // per stream [max_frames] scratch of the liveness walk
    int32_t* seg_off;      // per stream [n_seg + 1] exclusive scan of seg_count
    int64_t* nco_idx;      // per stream [max_frames] NCO events of ofdm_sync_ml (every detector peak)
    float* nco_ang;        // per stream [max_frames]
    int32_t* n_nco;        // per stream [1]; -1: the NCO follows the trigger list (pn, pnac, fixed)
    double* nco_sens;      // per stream [1] NCO sensitivity, set by the synchroniser stage
    float2* eq;            // per stream [eq_stride * occ] equalised vectors (acq_kernel -> sink_kernel)
    int64_t eq_stride;     // vectors per stream: n_max / L + max_frames + 2 bounds what the sampler can emit
    int64_t n_seg, seg_len, seg_cap;      // n_seg: detector segments of the LONGEST stream (table stride)
    int32_t max_frames;
};
int rx_workspace_layout(const ofdm_handle* h, const StreamSet& ss, int32_t max_frames, void* base, size_t bytes,
                        RxWorkspace* ws, size_t* need);

// launchers
int launch_make_packets(ofdm_handle* h, const uint8_t* payload, const int64_t* payload_off, int32_t n_pkts,
                        int whitening, uint8_t* pkts, const int64_t* pkt_off, cudaStream_t st);
int launch_tx(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames, int64_t first_frame,
              const int64_t* sym_off, int64_t total_syms, int32_t uniform_syms, const int64_t* stream_frame0,
              const int64_t* stream_out_off, int32_t n_streams, float2* out, cudaStream_t st, float2* map_tap = nullptr,
              float2* pre_tap = nullptr, float2* ifft_tap = nullptr);
int launch_xcorr(ofdm_handle* h, const float2* x, int64_t n, const float2* H, int nos, int ntaps, float2* out, cudaStream_t st);
size_t sync_alt_scratch_bytes(int64_t n);
int launch_sync_alt(ofdm_handle* h, const float2* y, int64_t n, int mode, float snr_db, ofdm_rx_io* io, RxWorkspace* ws,
                    void* scratch, size_t scratch_bytes, cudaStream_t st);
int launch_nco_taps(ofdm_handle* h, const float2* y, int64_t n, ofdm_rx_io* io, RxWorkspace* ws, float2* nco_out,
                    float2* sigmix_out, cudaStream_t st);
int launch_chan_filter(ofdm_handle* h, const float2* x, const StreamSet& ss, float2* y, cudaStream_t st);
int launch_sync_metric(ofdm_handle* h, const float2* y, int64_t n, float* mf, int64_t* first_nan, cudaStream_t st);
int launch_peak_detect(ofdm_handle* h, const float2* y, const float* mf, int64_t n, const int64_t* first_nan,
                       ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st);
int launch_trig_compact(ofdm_handle* h, const float2* y, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st);
int launch_sync_stream(ofdm_handle* h, const float2* y, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, int parts,
                       cudaStream_t st);
int launch_plan(ofdm_handle* h, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st);
int launch_nco_mode(ofdm_handle* h, RxWorkspace* ws, int S, double sens_times_n, cudaStream_t st);
int launch_sync_fixed(ofdm_handle* h, int64_t n, int32_t nsymbols, float freq_offset, ofdm_rx_io* io, RxWorkspace* ws,
                      cudaStream_t st);
// parts: bit 0 acq_kernel (derotation + FFT + frame acquisition), bit 1 sink_kernel (ofdm_frame_sink); 3 = both
int launch_demod(ofdm_handle* h, const float2* y, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st,
                 int parts = 3);
int launch_finish(ofdm_handle* h, int S, ofdm_rx_io* io, RxWorkspace* ws, cudaStream_t st);
// per-stream strides: n_frames / overflow 1, vbase max_frames + 1, everything else max_frames
int launch_liveness(int sms, int S, const int32_t* n_frames, const int64_t* vbase, int64_t vbase_stride, const int32_t* sess_nvec,
                    int32_t max_frames, int32_t* next, int32_t* exitf, int32_t* overflow, uint8_t* live, int force_general,
                    cudaStream_t st);
int launch_compact(ofdm_handle* h, const ofdm_rx_io* io, int32_t n_streams, uint8_t* out_bytes, int64_t capacity,
                   int64_t* msg_off, int32_t* msg_frame, uint32_t* ok_bits, int64_t* totals, int64_t* scratch,
                   cudaStream_t st);
int launch_channel(ofdm_handle* h, const float2* x, int64_t n, float cfo, double phase0, float sigma,
                   uint64_t seed, float2* y, cudaStream_t st);
int launch_sense(ofdm_sense_handle* s, const float2* x, int64_t n_frames, int shift, int32_t tune_delay,
                 int32_t dwell_delay, float* maxhold, float2* spectra, cudaStream_t st);
int launch_sense_decide(ofdm_sense_handle* s, const float* maxhold, int32_t n_avg, double threshold,
                        double* avg_inorder, uint8_t* free_bits, char* hex, cudaStream_t st);
int launch_sense_hop(ofdm_sense_handle* s, const double* avg_inorder, const uint8_t* free_bits, int32_t required_index,
                     int32_t* out, cudaStream_t st);
