/* ofdm_b200.h — C ABI of libofdm_b200.so, the B200 (sm_100a) drop-in for the OFDM
 * baseband hot path of rubiruchi/ofdm_uhd.
 *
 * Every entry point replaces a GNU Radio 3.6 block (or chain of blocks) that the
 * reference reaches through SWIG; the reference call site is cited next to it
 * (paths relative to /root/reference).  All data pointers are DEVICE pointers
 * unless the name starts with `host_`; every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*), returns 0 or a negative OFDM_E_* code and
 * never throws; ofdm_last_error() gives the text.  The library allocates only
 * its own constant tables (in ofdm_create / ofdm_sense_create); all stream,
 * packet and scratch buffers are owned by the caller.
 */
#ifndef OFDM_B200_H
#define OFDM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFDM_OK             0
#define OFDM_E_INVAL       -1   /* bad argument / unsupported layout            */
#define OFDM_E_CUDA        -2   /* a CUDA runtime call failed                   */
#define OFDM_E_NOMEM       -3   /* workspace / capacity too small               */

/* bits of ofdm_rx_io.status[0] (device side, sticky) */
#define OFDM_ST_TRIG_OVERFLOW   1u   /* more triggers than max_frames              */
#define OFDM_ST_SEG_OVERFLOW    2u   /* a detector segment overflowed its slot list */

typedef struct ofdm_handle ofdm_handle;
typedef struct ofdm_sense_handle ofdm_sense_handle;

typedef struct ofdm_cfg {
    int32_t fft_length;        /* options.fft_length      (ofdm.py:64)  */
    int32_t occupied_tones;    /* options.occupied_tones  (ofdm.py:65)  */
    int32_t cp_length;         /* options.cp_length       (ofdm.py:66)  */
    int32_t constellation_size;/* arity, ofdm.py:88-89                   */
    const float* host_constellation; /* 2*M floats (re,im): rotated_const of ofdm.py:94-101 */
    float   tx_amplitude;      /* transmit_path.py:44,56-62 (clamped to [0,1]) */
    int32_t device;            /* CUDA device ordinal                    */
    uint64_t pad_seed;         /* seed of the pad-symbol generator (upstream: libc rand()) */
    int32_t max_pkt_bytes;     /* bytes kept per received packet slot (<= 4096) */
    const char* host_carrier_map; /* hex data-carrier mask, NULL = "FE7F" (ofdm_mapper_bcv / ofdm_frame_sink default;
                                   the ctor argument of the commented call at ofdm.py:103-104) */
} ofdm_cfg;

const char* ofdm_last_error(void);
int ofdm_version(void);
/* Device self-test of the packed-fp32 (FFMA2 / FADD2 / FMUL2) arithmetic helpers the kernels are built on: n random
 * operand sets (zeros, denormals, huge values included) through every helper of the decision paths -- the ones that
 * replace the individually rounded float32 operations of ofdm_frame_acquisition / ofdm_frame_sink (ofdm.py:238-247,
 * ofdm_receiver.py~:127-129) -- against their scalar definitions.  host_out2[0] = operand sets with any mismatch
 * (must be 0), host_out2[1] = OR of the failing helpers' bits.  Synchronous; not part of the hot path. */
int ofdm_selftest_packed_math(int32_t device, int64_t n, uint64_t seed, int64_t* host_out2);

ofdm_handle* ofdm_create(const ofdm_cfg* cfg);
void ofdm_destroy(ofdm_handle* h);
/* transmit_path.set_tx_amplitude (transmit_path.py:56-62) */
int ofdm_set_tx_amplitude(ofdm_handle* h, float ampl);
/* derived layout: out[0]=zeros_on_left, [1]=data carriers, [2]=bits/carrier, [3]=symbol length,
 * [4]=channel filter taps, [5]=overlap-save FFT size, [6]=packet slot stride, [7]=reserved */
int ofdm_get_layout(const ofdm_handle* h, int32_t* out8);
/* copy of the channel-filter taps (gr.firdes.low_pass, ofdm_receiver.py~:69-76) to host memory */
int ofdm_get_chan_taps(const ofdm_handle* h, float* host_taps, int32_t max_taps);

/* ---- framing: ofdm_packet_utils.make_packet / unmake_packet (ofdm_packet_utils.py:99-143,169-191)
 *      + upstream crc.gen_and_append_crc32 / check_crc32 ------------------------------------- */
/* pkt f = header(4) || whiten(payload_f || crc32 || 0x55 [|| 0x55 pad]).  pkt_off[f+1]-pkt_off[f]
 * must equal the packet length computed by ofdm_packet_len(). */
int32_t ofdm_packet_len(int32_t payload_len, int pad_for_usrp);
int ofdm_make_packets(ofdm_handle* h, const uint8_t* payload, const int64_t* payload_off, int32_t n_pkts,
                      int whitening, uint8_t* pkts, const int64_t* pkt_off, void* stream);

/* ---- transmit: ofdm_mapper_bcv -> ofdm_insert_preamble -> fft_vcc(inverse) -> ofdm_cyclic_prefixer
 *      -> multiply_const(1/sqrt(N)) -> multiply_const(amp)   (ofdm.py:106-117, transmit_path.py:48) ----
 * sym_off[f] = number of OFDM symbols (preamble included) before frame f; frame f is written at
 * out[sym_off[f]*(N+cp)]; total_syms = sym_off[n_frames].  If every frame has the same number of
 * symbols pass it as uniform_syms (sym_off may then be NULL), else 0.  Pad carriers use
 * pad_index(seed, first_frame+f, symbol, carrier). */
int32_t ofdm_frame_symbols(const ofdm_handle* h, int32_t pkt_len);
int ofdm_tx_modulate_batch(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames,
                           int64_t first_frame, const int64_t* sym_off, int64_t total_syms, int32_t uniform_syms,
                           float* out_iq, void* stream);

/* ofdm_tx_modulate_batch plus the reference's options.log taps (ofdm.py:123-129), fft_length complex values per
 * OFDM symbol each, any of them NULL: mapper_out = ofdm_mapper_bcv's output vectors (data symbols only; frame f's
 * data symbol d is row sym_off[f] - f + d), preambles_out = the stream behind ofdm_insert_preamble (every symbol),
 * ifft_out = the unscaled fft_vcc output (every symbol).  The samples written to out_iq are those of
 * ofdm_tx_modulate_batch (the reference's fourth tap, ofdm_cp_adder_c.dat, up to the two scale stages). */
int ofdm_tx_modulate_taps(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames,
                          int64_t first_frame, const int64_t* sym_off, int64_t total_syms, int32_t uniform_syms,
                          float* out_iq, float* mapper_out, float* preambles_out, float* ifft_out, void* stream);

/* The same transmit chain for the frames of several independent streams in one launch (the transmit side of
 * ofdm_rx_demodulate_batch; the reference runs one benchmark_ofdm_tx.py flowgraph per stream): frames
 * [stream_frame0[s], stream_frame0[s+1]) belong to stream s, whose first OFDM symbol is written at
 * out[stream_out_off[s]] (complex samples) and whose later symbols follow it back to back; the frames of a stream are
 * numbered from first_frame again and its pad carriers use pad_seed + s.  stream_frame0: DEVICE int64[S+1]
 * (stream_frame0[0] = 0, stream_frame0[S] = n_frames), stream_out_off: DEVICE int64[S]. */
int ofdm_tx_modulate_streams(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames,
                             int64_t first_frame, const int64_t* sym_off, int64_t total_syms, int32_t uniform_syms,
                             const int64_t* stream_frame0, const int64_t* stream_out_off, int32_t n_streams,
                             float* out_iq, void* stream);

/* ---- receive -------------------------------------------------------------------------------- */
typedef struct ofdm_rx_io {
    /* capacity */
    int32_t max_frames;        /* capacity of every per-trigger / per-frame array below */
    int32_t pkt_stride;        /* bytes per packet slot in pkt_bytes                    */
    /* scratch (sizes from ofdm_rx_workspace_bytes) */
    void*   workspace;
    size_t  workspace_bytes;
    /* outputs (device) */
    uint32_t* status;          /* [1]  OFDM_ST_* bits                                   */
    int32_t* n_trig;           /* [1]  triggers found by ofdm_sync_pn's peak detector   */
    int64_t* trig_idx;         /* [max_frames] sample index of each trigger             */
    float*   trig_ang;         /* [max_frames] angle(P) latched at the trigger          */
    int32_t* n_frames;         /* [1]  frames the sampler emits                         */
    int64_t* frame_start;      /* [max_frames] first sample of the preamble vector      */
    int32_t* frame_ndata;      /* [max_frames] data vectors emitted after the preamble  */
    uint8_t* frame_live;       /* [max_frames] 1: the frame sink started on this preamble */
    uint8_t* frame_status;     /* [max_frames] 0 none, 1 bad header, 2 message, 3 stream ended */
    int32_t* pkt_len;          /* [max_frames] header length field (payload + 4)        */
    uint8_t* pkt_ok;           /* [max_frames] CRC-32 verdict of the message            */
    uint8_t* pkt_bytes;        /* [max_frames*pkt_stride] dewhitened payload || crc     */
    int64_t* counters;         /* [8] 0:frames 1:messages 2:crc_ok 3:payload bytes ok 4:samples 5:triggers */
    /* optional taps for parity debugging (may be NULL); indexed by vector = position in the
     * sampler's output stream */
    float*   eq_syms;          /* [max_vectors*occ*2] ofdm_frame_acquisition output      */
    uint8_t* sym_idx;          /* [max_vectors*ncar]  slicer decisions                   */
    float*   derot_syms;       /* [max_vectors*ncar*2] frame-sink derotated symbols      */
    int64_t  max_vectors;
    float*   fft_out;          /* [max_vectors*N*2] fft_demod output, shifted (ofdm_receiver-fft_out_c.dat) */
    float*   sampler_out;      /* [max_vectors*N*2] ofdm_sampler output vectors (ofdm_receiver-sampler_c.dat) */
} ofdm_rx_io;

size_t ofdm_rx_workspace_bytes(const ofdm_handle* h, int64_t n_samples, int32_t max_frames);

/* gr.fft_filter_ccc(1, firdes.low_pass(...))  (ofdm_receiver.py~:69-76,131) */
int ofdm_rx_chan_filter(ofdm_handle* h, const float* x_iq, int64_t n, float* y_iq, void* stream);
/* upstream ofdm_sync_pn up to add_const_ff(-1) (ofdm_receiver.py~:97-101): mf[n]; first_nan[0] gets
 * the first index whose metric is NaN (INT64_MAX if none) */
int ofdm_rx_sync_metric(ofdm_handle* h, const float* y_iq, int64_t n, float* mf, int64_t* first_nan,
                        void* stream);
/* gr.peak_detector_fb(0.20,0.20,30,0.001) + complex_to_arg + sample_and_hold (ofdm_sync_pn) */
int ofdm_rx_peak_detect(ofdm_handle* h, const float* y_iq, const float* mf, int64_t n, const int64_t* first_nan,
                        ofdm_rx_io* io, void* stream);
/* the whole ofdm_sync_pn hier-block (ofdm_receiver.py~:97-101) in one pass: y -> io->n_trig / trig_idx / trig_ang.
 * Uses the fused streaming kernel when 32*K = N/2 (N = 128, 256, 512) and cp <= N/2, else the two stages above
 * (ofdm_rx_demodulate also falls back to them for streams too short to fill the GPU with one warp per segment). */
int ofdm_rx_sync(ofdm_handle* h, const float* y_iq, int64_t n, ofdm_rx_io* io, void* stream);
/* upstream ofdm_sync_fixed(fft_length, cp_length, nsymbols, freq_offset) -- the reference's SYNC == "fixed" test mode
 * (ofdm_receiver.py~:108-119): a trigger at the last sample of the first symbol of every nsymbols-symbol packet
 * (index N+cp-1 + k*nsymbols*(N+cp)), every angle = pi*freq_offset, and the NCO already turning at that rate before the
 * first trigger.  Replaces ofdm_rx_sync in the stage sequence; fills io->n_trig / trig_idx / trig_ang. */
int ofdm_rx_sync_fixed(ofdm_handle* h, int64_t n, int32_t nsymbols, float freq_offset, ofdm_rx_io* io, void* stream);
/* The two other synchronisers ofdm_receiver.py~:89-107 names: host_sync = "pnac" (upstream ofdm_sync_pnac(fft_length,
 * cp_length, ks0time): cross-correlation with the known symbol, then its N/2-delayed auto-correlation against the N-sample
 * energy, threshold_ff(0,0,0)) or "ml" (upstream ofdm_sync_ml(fft_length, cp_length, snr, ks0time): van de Beek's
 * cyclic-prefix correlator, peak_detector_fb(0.2, 0.25, 30, 0.0005), timing gated by the known-symbol correlation; the
 * NCO's held angle changes at EVERY detector peak, nco_sensitivity = -1/N).  The reference hard-codes SYNC = "pn", so
 * these are functional restatements (plain kernels, not tuned; oracle: sync_pnac / sync_ml in oracle/ofdm_oracle.py).
 * ofdm_rx_sync_alt replaces ofdm_rx_sync in the stage sequence (timing triggers -> io->n_trig / trig_idx / trig_ang; for
 * "ml" the NCO events go to the workspace, ofdm_rx_workspace_ptr which = 7 count, 8 indices, 9 angles);
 * ofdm_rx_demodulate_alt is the whole chain.  scratch: DEVICE memory of ofdm_rx_sync_alt_scratch_bytes(n) bytes.
 * "ml" needs fft_length <= 2048 (its known-symbol correlator has fft_length taps).  Single stream. */
size_t ofdm_rx_sync_alt_scratch_bytes(const ofdm_handle* h, int64_t n_samples);
int ofdm_rx_sync_alt(ofdm_handle* h, const float* y_iq, int64_t n, const char* host_sync, float snr_db, ofdm_rx_io* io,
                     void* scratch, size_t scratch_bytes, void* stream);
int ofdm_rx_demodulate_alt(ofdm_handle* h, const float* x_iq, int64_t n, const char* host_sync, float snr_db, ofdm_rx_io* io,
                           void* scratch, size_t scratch_bytes, void* stream);
/* gr.frequency_modulator_fc + digital.ofdm_sampler (ofdm_receiver.py~:123-125,133-136) as a frame table */
int ofdm_rx_plan(ofdm_handle* h, int64_t n, ofdm_rx_io* io, void* stream);
/* multiply_cc (derotation) + fft_vcc(forward) + ofdm_frame_acquisition + ofdm_frame_sink
 * (ofdm_receiver.py~:124-129, ofdm.py:240-243) */
int ofdm_rx_demod(ofdm_handle* h, const float* y_iq, int64_t n, ofdm_rx_io* io, void* stream);
/* One kernel of the two multi-kernel stages above, for stage-level timing and profiling (ofdm_rx_sync and
 * ofdm_rx_demod run them back to back): stage 0 = the Schmidl-Cox metric kernel of ofdm_sync_pn (y -> workspace mf),
 * 1 = its peak_detector_fb kernel (mf -> per-segment triggers), 2 = trigger compaction + angle latch,
 * 3 = acq_kernel (multiply_cc + fft_vcc + ofdm_frame_acquisition, ofdm_receiver.py~:124-129),
 * 4 = sink_kernel (ofdm_frame_sink, ofdm.py:240-243).  Each needs the outputs of the stages before it. */
int ofdm_rx_stage(ofdm_handle* h, const float* y_iq, int64_t n, ofdm_rx_io* io, int32_t stage, void* stream);
/* frame-sink liveness (which preambles the sink accepted), unmake_packet (dewhiten + CRC,
 * ofdm.py:300-305) and the counters */
int ofdm_rx_finish(ofdm_handle* h, ofdm_rx_io* io, void* stream);
/* the liveness walk of ofdm_rx_finish on caller-supplied tables (stage-level entry, like ofdm_rx_sync_metric):
 * which frames does an ofdm_frame_sink (ofdm.py:240-243) start a session on, when a session started at frame f
 * consumes sess_nvec[f] vectors of a stream in which frame f's preamble vector sits at position vbase[f]?
 * n_frames: device int32[1]; vbase: device int64[>= F]; sess_nvec: device int32[F]; scratch: device int32[2*max_frames+2];
 * live: device uint8[max_frames].  force_general != 0 skips the short-exception-list fast path. */
/* the whole receive chain in the "fixed" test mode: chan_filt is gr.multiply_const_cc(1.0) there, so the capture is
 * sampled directly: ofdm_rx_sync_fixed -> ofdm_rx_plan -> ofdm_rx_demod -> ofdm_rx_finish */
int ofdm_rx_demodulate_fixed(ofdm_handle* h, const float* x_iq, int64_t n, int32_t nsymbols, float freq_offset,
                             ofdm_rx_io* io, void* stream);
int ofdm_rx_liveness(const int32_t* n_frames, const int64_t* vbase, const int32_t* sess_nvec, int32_t max_frames,
                     int32_t* scratch, uint8_t* live, int force_general, void* stream);
/* all of the above in order, no host synchronisation */
int ofdm_rx_demodulate(ofdm_handle* h, const float* x_iq, int64_t n, ofdm_rx_io* io, void* stream);
/* ---- many independent streams in one call ---------------------------------------------------------------
 * The reference's unit of work is one flowgraph per stream (benchmark_ofdm_rx.py:42-87 builds one
 * ofdm_receiver chain, ofdm_receiver.py~:131-142, per capture); BASELINE configs[2] runs 64 of them.  Here the S
 * streams lie back to back in ONE sample buffer, stream s = samples [stream_off[s], stream_off[s+1]), and every
 * kernel of the chain takes the stream index from its grid: each stream starts from zero history (filter, window
 * sums, detector average, NCO phase, sink state) exactly as if it had been passed to ofdm_rx_demodulate alone, and
 * no launch is repeated per stream.
 *   stream_off : DEVICE int64[S+1], ascending, stream_off[0] >= 0 (offsets that are multiples of 4 samples keep
 *                the 32-byte vector accesses)
 *   io         : one ofdm_rx_io whose arrays hold S consecutive per-stream tables: every [max_frames] array becomes
 *                [S][max_frames] (io->max_frames = capacity PER STREAM), pkt_bytes [S][max_frames][pkt_stride],
 *                status / n_trig / n_frames [S], counters [S][8]; workspace >= ofdm_rx_workspace_bytes_batch().
 *                The parity taps (eq_syms / sym_idx / derot_syms) must be NULL.
 *   total_samples, max_stream_samples : stream_off[S] and the longest stream (host knowledge for grid sizing).
 * Layouts without the streaming synchroniser (fft_length 64, cp_length > fft_length/2) return OFDM_E_INVAL. */
size_t ofdm_rx_workspace_bytes_batch(const ofdm_handle* h, int32_t n_streams, int64_t total_samples,
                                     int64_t max_stream_samples, int32_t max_frames_per_stream);
int ofdm_rx_demodulate_batch(ofdm_handle* h, const float* x_iq, const int64_t* stream_off, int32_t n_streams,
                             int64_t total_samples, int64_t max_stream_samples, ofdm_rx_io* io, void* stream);
/* options.log taps of the receiver that are per-sample streams (ofdm_receiver.py~:150-151), to be called after
 * ofdm_rx_plan / ofdm_rx_demodulate on the same io: nco_out[n] = e^{j phi[n]} (gr.frequency_modulator_fc driven by the
 * held sync angle), sigmix_out[n] = y[n] * nco_out[n]; y is the filtered stream (ofdm_rx_workspace_ptr(.., 0)).  Either
 * output may be NULL.  Single stream. */
int ofdm_rx_nco_taps(ofdm_handle* h, const float* y_iq, int64_t n, ofdm_rx_io* io, float* nco_out, float* sigmix_out,
                     void* stream);
/* The hand-over of the delivered messages to the host (the reference pushes every message the frame sink completes
 * into a gr.msg_queue popped by _queue_watcher_thread, ofdm.py:290-305): after ofdm_rx_finish / ofdm_rx_demodulate
 * [_batch], pack the messages of all n_streams streams (frames with frame_live && frame_status == 2, in stream and
 * arrival order) into one dense byte array, so that the device -> host copy is sized by what was delivered instead of
 * max_frames * pkt_stride.  All pointers are DEVICE pointers:
 *   out_bytes [out_capacity]               dewhitened payload || crc of message 0, 1, ... back to back
 *   msg_off   [n_streams*max_frames + 1]   byte offset of each message in out_bytes (msg_off[n_msgs] = total bytes)
 *   msg_frame [n_streams*max_frames]       frame slot s*max_frames + f the message came from
 *   ok_bits   [(n_streams*max_frames+31)/32] uint32, bit m = CRC-32 verdict of message m
 *   totals    [3]  n_msgs, bytes in all, messages whose bytes fitted out_capacity
 *   scratch   [2 * ((n_streams*max_frames + 1023) / 1024)] int64 */
int ofdm_rx_compact(ofdm_handle* h, const ofdm_rx_io* io, int32_t n_streams, uint8_t* out_bytes, int64_t out_capacity,
                    int64_t* msg_off, int32_t* msg_frame, uint32_t* ok_bits, int64_t* totals, int64_t* scratch,
                    void* stream);
/* pointers into the workspace for parity tests: which = 0 filtered stream y (2n floats), 1 metric mf (n floats) */
void* ofdm_rx_workspace_ptr(const ofdm_handle* h, const ofdm_rx_io* io, int64_t n, int which);

/* ---- synthetic channel used by the loopback drivers and the bench (not a reference block) ----
 * y[n] = x[n]*exp(j*(phase0 + 2*pi*cfo/N*n)) + sigma*(gauss+j*gauss), counter-based noise */
int ofdm_channel(ofdm_handle* h, const float* x_iq, int64_t n, float cfo_subcarriers, double phase0,
                 float sigma, uint64_t seed, float* y_iq, void* stream);

/* ---- spectrum sensing: stream_to_vector -> fft_vcc(N,True,blackmanharris[,shift]) ->
 *      complex_to_mag_squared -> bin_statistics_f  (secondary_tx.py:163-202, usrp_fft_save.py:58-62) ---- */
ofdm_sense_handle* ofdm_sense_create(int32_t fft_size, int32_t device);
void ofdm_sense_destroy(ofdm_sense_handle* s);
/* max-hold power per dwell: maxhold[d][N], d < n_frames / (tune_delay+dwell_delay) */
int ofdm_sense(ofdm_sense_handle* s, const float* x_iq, int64_t n_frames, int shift, int32_t tune_delay,
               int32_t dwell_delay, float* maxhold, void* stream);
/* raw windowed spectra (usrp_fft_save.py:61): out[frame][N] complex */
int ofdm_sense_fft(ofdm_sense_handle* s, const float* x_iq, int64_t n_frames, int shift, float* out_iq,
                   void* stream);
/* sense_loop decision (secondary_tx.py:237-266,306-331): mean of n_avg dwell vectors, free = !(avg > thr),
 * halves swapped into frequency order, nibble-packed hex (first bit = LSB).  avg_inorder: double[N],
 * free_bits: uint8[N], hex: char[N/4] (no terminator) */
int ofdm_sense_decide(ofdm_sense_handle* s, const float* maxhold, int32_t n_avg, double threshold,
                      double* avg_inorder, uint8_t* free_bits, char* hex, void* stream);
/* sense_loop hop decision (secondary_tx.py:268-295) on the outputs of ofdm_sense_decide: out[0] = occupied bins in
 * thrshold_inorder[required_index-16 : required_index+16] (Python slice semantics), out[1] = centre bin (+8) of the
 * quietest 17-bin window over bins 200 .. N-218 (-1 if none sums below 50), out[2] = window length.  out: int32[3] */
int ofdm_sense_hop(ofdm_sense_handle* s, const double* avg_inorder, const uint8_t* free_bits, int32_t required_index,
                   int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFDM_B200_H */
