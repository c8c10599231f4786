// Transmit side: packet framing, the fused mapper + preamble + IFFT + cyclic-prefix kernel,
// and the synthetic channel used by the loopback drivers.
#include "internal.h"
#include <stdlib.h>
#include "fft.cuh"

// ---------------------------------------------------------------------------------------------
// make_packet (ofdm_packet_utils.py:99-143): one thread per packet.  CRC-32 is the upstream
// digital.crc32 (MSB-first 0x04C11DB7, init/final all ones; digital_swig.py:3151-3168).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void frame_one_packet(const uint8_t* src, int plen, uint8_t* dst, int total, int whitening,
                                                 const uint8_t* __restrict__ mask, const uint32_t* s_crc) {
    const uint32_t v = (uint32_t)((plen + 4) & 0x0FFF);            // whitener offset 0 (ofdm.py:143)
    dst[0] = (uint8_t)(v >> 8); dst[1] = (uint8_t)v; dst[2] = (uint8_t)(v >> 8); dst[3] = (uint8_t)v;
    uint32_t crc = 0xFFFFFFFFu;
    int o = 0;                                                      // offset into the whitened body
    // four bytes per step: one mask word, four independent CRC table lookups (s_crc = [T0 | T1 | T2 | T3])
    const uint32_t* m4 = (const uint32_t*)mask;
    for (; o + 4 <= plen; o += 4) {
        const uint32_t b0 = src[o], b1 = src[o + 1], b2 = src[o + 2], b3 = src[o + 3];
        crc = crc32_step4(crc, (b0 << 24) | (b1 << 16) | (b2 << 8) | b3, s_crc);
        const uint32_t m = whitening ? LDG(m4 + (o >> 2)) : 0u;
        dst[4 + o] = (uint8_t)(b0 ^ m); dst[5 + o] = (uint8_t)(b1 ^ (m >> 8));
        dst[6 + o] = (uint8_t)(b2 ^ (m >> 16)); dst[7 + o] = (uint8_t)(b3 ^ (m >> 24));
    }
    for (int i = o; i < plen; ++i, ++o) {
        const uint8_t b = src[i];
        crc = s_crc[(b ^ (crc >> 24)) & 0xFF] ^ (crc << 8);
        dst[4 + o] = whitening ? (uint8_t)(b ^ mask[o]) : b;
    }
    crc = ~crc;
    for (int i = 0; i < 4; ++i, ++o) {
        const uint8_t b = (uint8_t)(crc >> (24 - 8 * i));
        dst[4 + o] = whitening ? (uint8_t)(b ^ mask[o]) : b;
    }
    // the body never exceeds the 4096-byte mask: the host side rejects such payloads where the reference's
    // make_packet raises (ofdm_packet_utils.py:117-135); nothing wraps around the table here
    for (; 4 + o < total && o < 4096; ++o) dst[4 + o] = whitening ? (uint8_t)(0x55 ^ mask[o]) : (uint8_t)0x55;
}

__device__ __forceinline__ void warp_copy_bytes(uint8_t* dst, const uint8_t* src, int nbytes, int lane) {
    // src and dst have the same address phase modulo 4: byte head, word body, byte tail
    int head = (int)((4 - ((uintptr_t)src & 3)) & 3);
    if (head > nbytes) head = nbytes;
    if (lane < head) dst[lane] = src[lane];
    const uint32_t* s4 = (const uint32_t*)(src + head);
    uint32_t* d4 = (uint32_t*)(dst + head);
    const int nw = (nbytes - head) >> 2;
    for (int i = lane; i < nw; i += 32) d4[i] = s4[i];
    const int done = head + 4 * nw;
    if (lane < nbytes - done) dst[done + lane] = src[done + lane];
}

// One warp per 32 consecutive packets: their payloads (contiguous in the input) are staged through shared
// memory with coalesced word copies, each lane frames one packet there, and the 32 framed packets (contiguous
// in the output) are written back coalesced.  Groups that do not fit the staging buffers are framed straight
// from / to global memory by the same lanes.
constexpr int MP_IN = 13312;      // 32 payloads (32 * 402 = 12 864 for the benchmark packet)
constexpr int MP_OUT = 13824;     // 32 framed packets

__global__ void __launch_bounds__(32) make_packets_kernel(const uint8_t* __restrict__ payload,
                                                           const int64_t* __restrict__ payload_off, int n_pkts,
                                                           int whitening, uint8_t* __restrict__ pkts,
                                                           const int64_t* __restrict__ pkt_off,
                                                           const uint8_t* __restrict__ mask,
                                                           const uint32_t* __restrict__ crctab) {
    __shared__ uint32_t s_crc[1024];                                // T0 .. T3
    __shared__ __align__(16) uint8_t s_in[MP_IN];
    __shared__ __align__(16) uint8_t s_out[MP_OUT];
    const int lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) s_crc[i] = crctab[i];
    for (int i = lane; i < 768; i += 32) s_crc[256 + i] = crctab[OFDM_CRC_SLICE + i];
    __syncwarp();
    const int f0 = blockIdx.x * 32;
    if (f0 >= n_pkts) return;
    const int nf = (n_pkts - f0 < 32) ? n_pkts - f0 : 32;
    const int64_t in0 = payload_off[f0], out0 = pkt_off[f0];
    const int64_t in_bytes = payload_off[f0 + nf] - in0, out_bytes = pkt_off[f0 + nf] - out0;
    const int p = f0 + lane;
    if (in_bytes + 4 > MP_IN || out_bytes + 4 > MP_OUT) {
        if (lane < nf)
            frame_one_packet(payload + payload_off[p], (int)(payload_off[p + 1] - payload_off[p]), pkts + pkt_off[p],
                             (int)(pkt_off[p + 1] - pkt_off[p]), whitening, mask, s_crc);
        return;
    }
    uint8_t* bi = s_in + ((uintptr_t)(payload + in0) & 3);       // same phase as the source / destination
    uint8_t* bo = s_out + ((uintptr_t)(pkts + out0) & 3);
    warp_copy_bytes(bi, payload + in0, (int)in_bytes, lane);
    __syncwarp();
    if (lane < nf)
        frame_one_packet(bi + (int)(payload_off[p] - in0), (int)(payload_off[p + 1] - payload_off[p]),
                         bo + (int)(pkt_off[p] - out0), (int)(pkt_off[p + 1] - pkt_off[p]), whitening, mask, s_crc);
    __syncwarp();
    warp_copy_bytes(pkts + out0, bo, (int)out_bytes, lane);
}

int launch_make_packets(ofdm_handle* h, const uint8_t* payload, const int64_t* payload_off, int32_t n_pkts,
                        int whitening, uint8_t* pkts, const int64_t* pkt_off, cudaStream_t st) {
    make_packets_kernel<<<(n_pkts + 31) / 32, 32, 0, st>>>(payload, payload_off, n_pkts, whitening, pkts, pkt_off,
                                                           h->d_mask, h->d_crctab);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// ---------------------------------------------------------------------------------------------
// K_TX: ofdm_mapper_bcv + ofdm_insert_preamble + fft_vcc(inverse, shifted) + ofdm_cyclic_prefixer
//       + multiply_const(1/sqrt N) + multiply_const(amp)      (ofdm.py:106-117, transmit_path.py:48)
// One group of T = N/E threads builds one OFDM symbol; G groups per CTA.
// ---------------------------------------------------------------------------------------------
struct TxParams {
    const uint8_t* pkts;
    const int64_t* pkt_off;
    const int64_t* sym_off;
    int n_frames;
    int uniform_syms;
    int64_t total_syms;
    int64_t first_frame;
    uint64_t seed;
    float2* out;
    const float2* cst;
    const int16_t* bin2car;
    const float2* tw;
    const float2* pre_time;
    int cp, ncar, nbits, M;
    float s1, amp;
    // several streams in one batch (n_streams > 0): frames [stream_frame0[s], stream_frame0[s+1]) belong to stream s,
    // whose first symbol is written at out[stream_out_off[s]]; its frames are numbered from first_frame again and
    // its pad symbols use seed + s
    const int64_t* stream_frame0;
    const int64_t* stream_out_off;
    int n_streams;
    // options.log taps (ofdm.py:123-129): mapper output (data symbols only), the symbol stream behind
    // ofdm_insert_preamble, and the unscaled IFFT output, fft_length complex values per symbol
    float2* map_tap;             // [data symbols][N]
    float2* pre_tap;             // [symbols][N]
    float2* ifft_tap;            // [symbols][N]
    const float2* pre_freq;      // [N] the known symbol as the mapper-order vector
    const float2* pre_ifft;      // [N] its unscaled IFFT
};

// The carrier map and the bytes of the symbol are staged in shared memory before the first pass: from global memory
// the map lookup and the packet byte it leads to are two dependent loads per carrier (a quarter of this kernel's
// stall samples were long-scoreboard waits on them).
template <int N, bool TAPS>
struct TxLoad {
    const TxParams& p;
    const uint8_t* s_bytes;      // packet bytes from byte0 on (this symbol's share)
    int byte0;
    int pkt_bits;
    int64_t frame_id;
    uint64_t seed;
    int dsym;                    // data symbol number inside the frame
    const float2* s_cst;
    const int16_t* s_b2c;
    float2* map_tap;             // this symbol's row of the mapper tap / of the preambles tap (TAPS only)
    float2* pre_tap;
    __device__ __forceinline__ float2 operator()(int idx, int) const {
        const float2 o = value(idx);
        if (TAPS) {
            const int v = (idx + N / 2) & (N - 1);
            if (map_tap) map_tap[v] = o;
            if (pre_tap) pre_tap[v] = o;
        }
        return o;
    }
    __device__ __forceinline__ float2 value(int idx) const {
        const int v = (idx + N / 2) & (N - 1);          // ifftshift: IFFT input idx holds vector bin v
        const int c = s_b2c[v];
        if (c < 0) return make_float2(0.f, 0.f);
        const int bit0 = (dsym * p.ncar + c) * p.nbits;          // < 8 * 4112: 32-bit math
        uint32_t val;
        if (bit0 + p.nbits <= pkt_bits) {
            const int b = (bit0 >> 3) - byte0, sh = bit0 & 7;
            uint32_t w = s_bytes[b];
            if (sh + p.nbits > 8) w |= ((uint32_t)s_bytes[b + 1]) << 8;
            val = (w >> sh) & ((1u << p.nbits) - 1u);
        } else {
            val = pad_index(seed, (uint64_t)frame_id, (uint32_t)dsym, (uint32_t)c, (uint32_t)p.M);
        }
        return s_cst[val];
    }
};

template <int N, bool TAPS>
struct TxStore {
    float2* dst;                 // start of this symbol (its cyclic prefix)
    int cp;
    float s1, amp;
    float2* ifft_tap;            // this symbol's row of the IFFT tap (TAPS only)
    __device__ __forceinline__ void operator()(int n, float2 v, int) const {
        if (TAPS && ifft_tap) ifft_tap[n] = v;
        // float32 after each multiply_const, like the two upstream blocks
        const float2 o = cscale_x(cscale_x(v, s1), amp);
        dst[cp + n] = o;
        if (n >= N - cp) dst[n - (N - cp)] = o;
    }
};

// Branch-free mapper (every kernel below without taps).  Tables are per FFT INPUT index (the ifftshift is folded in): s_bo[idx] = bit offset of the
// carrier inside a symbol's bit field, -1 for an unused bin; the symbol's bytes are staged as overlapping 16-bit words
// (s_w[b] = byte b | byte b+1 << 8), so a carrier's bits are one load and one shift.  First-pass rows [RLO, RHI) hold no
// carrier at all (rows 7 .. 24 of 32 when occupied_tones / fft_length <= 0.41): compile-time zeros, no mapper work and no
// branch.  PAD: the symbol reaches behind the end of the packet (last symbol of a frame) and fills up with pad symbols.
// pad symbol of the carrier at bit offset b of a symbol's bit field; carrier ordinal = b / nbits, nbits in {1, 2, 3, 4, 6,
// 8}: a shift, then an exact multiply-shift division by 3
__device__ __forceinline__ uint32_t tx_pad_value(uint64_t base, int b, int nbits, int M) {
    const int sh = __ffs(nbits) - 1;
    const unsigned pre = (unsigned)b >> sh;
    const unsigned c = ((nbits >> sh) == 3) ? (pre * 43691u) >> 17 : pre;
    return pad_index_from(base, c, (uint32_t)M);
}
static __device__ __noinline__ uint32_t tx_pad_value_call(uint64_t base, int b, int nbits, int M) { return tx_pad_value(base, b, nbits, M); }
// PADINL: the pad hash inline (the three-pass kernels: 16 points per thread) or behind a call (the warp kernels: 32
// points per thread, where 32 inlined copies cost the instruction cache more than the calls: 512-point tx +3 %)
template <int N, int RLO, int RHI, bool PAD, bool PADINL = false>
struct TxLoadW {
    const int16_t* s_bo;
    const uint16_t* s_w;
    const float2* s_cst;
    unsigned vmask;
    int bitbase, rel0, pkt_bits, nbits, dsym, M;
    int64_t frame_id;
    uint64_t seed;
    __device__ __forceinline__ float2 operator()(int idx, int slot) const {
        if (slot >= RLO && slot < RHI) return make_float2(0.f, 0.f);
        const int bo = s_bo[idx];
        const int b = bo < 0 ? 0 : bo;
        const int rel = rel0 + b;
        uint32_t val = ((uint32_t)s_w[rel >> 3] >> (rel & 7)) & vmask;
        if (PAD && bo >= 0 && bitbase + b + nbits > pkt_bits) {
            const uint64_t base = pad_base(seed, (uint64_t)frame_id, (uint32_t)dsym);
            val = PADINL ? tx_pad_value(base, b, nbits, M) : tx_pad_value_call(base, b, nbits, M);
        }
        const float2 pt = s_cst[val];
        return bo < 0 ? make_float2(0.f, 0.f) : pt;
    }
};

template <int N, int G, bool TAPS>
__global__ void __launch_bounds__(G * (N / FftPlan<N>::E), (FftPlan<N>::E == 8 ? 1024 : 512) / (G * (N / FftPlan<N>::E))) tx_kernel(const TxParams p) {
    constexpr int T = N / FftPlan<N>::E;
    constexpr int SB = fft_smem_elems<N>();
    extern __shared__ float2 smem[];
    float2* s_cst = smem;                                  // [M]
    float2* bufs = smem + 256;                             // G * 2 * SB
    int16_t* s_b2c = (int16_t*)(bufs + (size_t)G * 2 * SB);     // [N]
    const int sym_bytes = (p.ncar * p.nbits + 7) / 8 + 2;       // bytes one symbol can touch (+ straddle)
    const int sb_stride = (sym_bytes + 15) & ~15;
    uint8_t* s_sym = (uint8_t*)(s_b2c + N);                     // [G][sb_stride]
    const int g = threadIdx.x / T;
    const int tid = threadIdx.x - g * T;
    for (int i = threadIdx.x; i < p.M; i += blockDim.x) s_cst[i] = p.cst[i];
    if (TAPS) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) s_b2c[i] = p.bin2car[i];
    } else {                                               // per IFFT input index: bit offset of the carrier, -1 if unused
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const int c = p.bin2car[(i + N / 2) & (N - 1)];
            s_b2c[i] = (int16_t)(c < 0 ? -1 : c * p.nbits);
        }
    }
    __syncthreads();
    float2* bufA = bufs + (size_t)g * 2 * SB;
    float2* bufB = bufA + SB;
    uint8_t* my_bytes = s_sym + (size_t)g * sb_stride * (TAPS ? 1 : 2);     // bytes (TAPS) / overlapping 16-bit words
    const int L = N + p.cp;
    auto bar = [] { __syncthreads(); };
    const unsigned total = (unsigned)p.total_syms;               // launcher guarantees < 2^31
    for (unsigned base = blockIdx.x * G; base < total; base += gridDim.x * G) {
        const unsigned s = base + g;
        const bool active = s < total;
        int f = 0;
        int m = 0;
        if (active) {
            if (p.uniform_syms > 0) {
                f = (int)(s / (unsigned)p.uniform_syms);
                m = (int)(s - (unsigned)f * (unsigned)p.uniform_syms);
            } else {
                int lo = 0, hi = p.n_frames;              // last f with sym_off[f] <= s
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (LDG(p.sym_off + mid) <= (int64_t)s) lo = mid; else hi = mid;
                }
                f = lo;
                m = (int)(s - LDG(p.sym_off + f));
            }
        }
        float2* dst = p.out + (size_t)s * L;
        int64_t frame_id = p.first_frame + f;
        uint64_t seed = p.seed;
        if (p.n_streams > 0 && active) {
            int lo = 0, hi = p.n_streams;                 // last stream with stream_frame0[q] <= f
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (LDG(p.stream_frame0 + mid) <= (int64_t)f) lo = mid; else hi = mid;
            }
            const int64_t f0 = LDG(p.stream_frame0 + lo);
            const int64_t sym0 = p.uniform_syms > 0 ? f0 * p.uniform_syms : LDG(p.sym_off + f0);
            dst = p.out + LDG(p.stream_out_off + lo) + ((int64_t)s - sym0) * L;
            frame_id = p.first_frame + ((int64_t)f - f0);
            seed += (uint64_t)lo;
        }
        const bool data = active && m > 0;
        if (active && m == 0) {
            // ofdm_insert_preamble: the pre-modulated known symbol
            for (int i = tid; i < L; i += T) {
                float2 v = LDG(p.pre_time + i);
                dst[i] = cscale_x(v, p.amp);
            }
            if (TAPS) {
                for (int i = tid; i < N; i += T) {
                    if (p.pre_tap) p.pre_tap[(size_t)s * N + i] = LDG(p.pre_freq + i);
                    if (p.ifft_tap) p.ifft_tap[(size_t)s * N + i] = LDG(p.pre_ifft + i);
                }
            }
        }
        const int64_t o0 = active ? LDG(p.pkt_off + f) : 0;
        const int64_t o1 = active ? LDG(p.pkt_off + f + 1) : 0;
        const int pkt_len = (int)(o1 - o0);
        const int byte0 = data ? ((m - 1) * p.ncar * p.nbits) >> 3 : 0;
        if (data) {                                              // (the barriers after the previous first pass cover my_bytes)
            int nb = pkt_len - byte0;
            if (nb > sym_bytes) nb = sym_bytes;
            const uint8_t* src = p.pkts + o0 + byte0;
            if (TAPS) {
                for (int i = tid; i < nb; i += T) my_bytes[i] = LDG(src + i);
            } else {
                const int left = pkt_len - byte0;
                for (int i = tid; i < nb; i += T)
                    ((uint16_t*)my_bytes)[i] = (uint16_t)(LDG(src + i) | ((i + 1 < left ? (uint32_t)LDG(src + i + 1) : 0u) << 8));
            }
        }
        bar();
        // data symbols before this one in the batch: s minus the preambles of frames 0 .. f
        TxLoad<N, TAPS> ld{p, my_bytes, byte0, pkt_len * 8, frame_id, seed, m - 1, s_cst, s_b2c,
                           (TAPS && p.map_tap) ? p.map_tap + ((size_t)s - (size_t)f - 1) * N : nullptr,
                           (TAPS && p.pre_tap) ? p.pre_tap + (size_t)s * N : nullptr};
        TxStore<N, TAPS> st{dst, p.cp, p.s1, p.amp, (TAPS && p.ifft_tap) ? p.ifft_tap + (size_t)s * N : nullptr};
        using P = FftPlan<N>;
        constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2];
        if constexpr (TAPS) {
            if (data) fft_pass<N, R0, 1, 1>(tid, p.tw, ld, SmemOut{bufA});
        } else {
            const int bitbase = (m - 1) * p.ncar * p.nbits;
            const unsigned vmask = (1u << p.nbits) - 1u;
            if (data && bitbase + p.ncar * p.nbits <= pkt_len * 8) {
                TxLoadW<N, 0, 0, false> ldw{s_b2c, (const uint16_t*)my_bytes, s_cst, vmask, bitbase, bitbase - 8 * byte0,
                                            pkt_len * 8, p.nbits, m - 1, p.M, frame_id, seed};
                fft_pass<N, R0, 1, 1>(tid, p.tw, ldw, SmemOut{bufA});
            } else if (data) {
                TxLoadW<N, 0, 0, true, true> ldw{s_b2c, (const uint16_t*)my_bytes, s_cst, vmask, bitbase, bitbase - 8 * byte0,
                                           pkt_len * 8, p.nbits, m - 1, p.M, frame_id, seed};
                fft_pass<N, R0, 1, 1>(tid, p.tw, ldw, SmemOut{bufA});
            }
        }
        bar();
        if constexpr (P::NP == 2) {
            if (data) fft_pass<N, R1, R0, 1>(tid, p.tw, SmemIn{bufA}, st);
            bar();
        } else {
            if (data) fft_pass<N, R1, R0, 1>(tid, p.tw, SmemIn{bufA}, SmemOut{bufB});
            bar();
            if (data) fft_pass<N, R2, R0 * R1, 1>(tid, p.tw, SmemIn{bufB}, st);
        }
    }
}

template <int N, int G>
static int launch_tx_n(ofdm_handle* h, const TxParams& p, cudaStream_t st) {
    constexpr int T = N / FftPlan<N>::E;
    const int sym_bytes = (p.ncar * p.nbits + 7) / 8 + 2;
    size_t smem = (256 + (size_t)G * 2 * fft_smem_elems<N>()) * sizeof(float2) + (size_t)N * sizeof(int16_t) +
                  (size_t)G * ((sym_bytes + 15) & ~15) * 2;
    const bool taps = p.map_tap || p.pre_tap || p.ifft_tap;
    if (taps) { OFDM_SET_MAX_SMEM((tx_kernel<N, G, true>), smem, h->device); } else { OFDM_SET_MAX_SMEM((tx_kernel<N, G, false>), smem, h->device); }
    const int sms = h->sms;
    int64_t want = (p.total_syms + G - 1) / G;
    int64_t cap = (int64_t)sms * (64 / G);
    int grid = (int)(want < cap ? want : cap);
    if (taps) tx_kernel<N, G, true><<<grid, G * T, smem, st>>>(p);
    else tx_kernel<N, G, false><<<grid, G * T, smem, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

// Warp-plan transmit kernel (N = 512 / 1024): a transform lives on T = N/32 <= 32 lanes of one warp (two symbols per
// warp at 512), radix-32 first pass, ONE shared-memory exchange, __syncwarp() only.  One warp per CTA.
template <int N, bool TAPS, int RLO, int RHI>
__global__ void __launch_bounds__(32, 16) tx_warp_kernel(const TxParams p) {
    using P = typename FftPlanW<N>::type;
    constexpr int T = N / P::E, G = 32 / T;
    constexpr int SB = FFT_PAD32(N) + 2;
    extern __shared__ float2 smem[];
    float2* s_cst = smem;                                  // [M]
    float2* bufs = smem + ((p.M + 15) & ~15);              // G * SB
    int16_t* s_b2c = (int16_t*)(bufs + (size_t)G * SB);    // [N]
    const int sym_bytes = (p.ncar * p.nbits + 7) / 8 + 2;
    const int sb_stride = (sym_bytes + 15) & ~15;
    uint8_t* s_sym = (uint8_t*)(s_b2c + N);                // [G][sb_stride] (TAPS) / uint16 [G][sb_stride] (word staging)
    const int g = threadIdx.x / T;
    const int tid = threadIdx.x - g * T;
    for (int i = threadIdx.x; i < p.M; i += 32) s_cst[i] = p.cst[i];
    if (TAPS) {
        for (int i = threadIdx.x; i < N; i += 32) s_b2c[i] = p.bin2car[i];
    } else {
        for (int i = threadIdx.x; i < N; i += 32) {
            const int c = p.bin2car[(i + N / 2) & (N - 1)];
            s_b2c[i] = (int16_t)(c < 0 ? -1 : c * p.nbits);
        }
    }
    __syncwarp();
    float2* buf = bufs + (size_t)g * SB;
    uint8_t* my_bytes = s_sym + (size_t)g * sb_stride * (TAPS ? 1 : 2);
    const int L = N + p.cp;
    const unsigned total = (unsigned)p.total_syms;
    for (unsigned base = blockIdx.x * G; base < total; base += gridDim.x * G) {
        const unsigned s = base + g;
        const bool active = s < total;
        int f = 0, m = 0;
        if (active) {
            if (p.uniform_syms > 0) {
                f = (int)(s / (unsigned)p.uniform_syms);
                m = (int)(s - (unsigned)f * (unsigned)p.uniform_syms);
            } else {
                int lo = 0, hi = p.n_frames;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (LDG(p.sym_off + mid) <= (int64_t)s) lo = mid; else hi = mid;
                }
                f = lo;
                m = (int)(s - LDG(p.sym_off + f));
            }
        }
        float2* dst = p.out + (size_t)s * L;
        int64_t frame_id = p.first_frame + f;
        uint64_t seed = p.seed;
        if (p.n_streams > 0 && active) {
            int lo = 0, hi = p.n_streams;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (LDG(p.stream_frame0 + mid) <= (int64_t)f) lo = mid; else hi = mid;
            }
            const int64_t f0 = LDG(p.stream_frame0 + lo);
            const int64_t sym0 = p.uniform_syms > 0 ? f0 * p.uniform_syms : LDG(p.sym_off + f0);
            dst = p.out + LDG(p.stream_out_off + lo) + ((int64_t)s - sym0) * L;
            frame_id = p.first_frame + ((int64_t)f - f0);
            seed += (uint64_t)lo;
        }
        const bool data = active && m > 0;
        if (active && m == 0) {
            for (int i = tid; i < L; i += T) dst[i] = cscale_x(LDG(p.pre_time + i), p.amp);
            if (TAPS) {
                for (int i = tid; i < N; i += T) {
                    if (p.pre_tap) p.pre_tap[(size_t)s * N + i] = LDG(p.pre_freq + i);
                    if (p.ifft_tap) p.ifft_tap[(size_t)s * N + i] = LDG(p.pre_ifft + i);
                }
            }
        }
        const int64_t o0 = active ? LDG(p.pkt_off + f) : 0;
        const int64_t o1 = active ? LDG(p.pkt_off + f + 1) : 0;
        const int pkt_len = (int)(o1 - o0);
        const int byte0 = data ? ((m - 1) * p.ncar * p.nbits) >> 3 : 0;
        if (data) {
            int nb = pkt_len - byte0;
            if (nb > sym_bytes) nb = sym_bytes;
            const uint8_t* src = p.pkts + o0 + byte0;
            if (TAPS) {
                for (int i = tid; i < nb; i += T) my_bytes[i] = LDG(src + i);
            } else {
                const int left = pkt_len - byte0;
                for (int i = tid; i < nb; i += T)
                    ((uint16_t*)my_bytes)[i] = (uint16_t)(LDG(src + i) | ((i + 1 < left ? (uint32_t)LDG(src + i + 1) : 0u) << 8));
            }
        }
        __syncwarp();
        TxLoad<N, TAPS> ld{p, my_bytes, byte0, pkt_len * 8, frame_id, seed, m - 1, s_cst, s_b2c,
                           (TAPS && p.map_tap) ? p.map_tap + ((size_t)s - (size_t)f - 1) * N : nullptr,
                           (TAPS && p.pre_tap) ? p.pre_tap + (size_t)s * N : nullptr};
        TxStore<N, TAPS> st{dst, p.cp, p.s1, p.amp, (TAPS && p.ifft_tap) ? p.ifft_tap + (size_t)s * N : nullptr};
        if constexpr (TAPS) {
            if (data) fft_pass<N, P::R[0], 1, 1, decltype(ld), SmemOut32, false, P>(tid, p.tw, ld, SmemOut32{buf});
        } else {
            const int bitbase = (m - 1) * p.ncar * p.nbits;
            const unsigned vmask = (1u << p.nbits) - 1u;
            if (data && bitbase + p.ncar * p.nbits <= pkt_len * 8) {
                TxLoadW<N, RLO, RHI, false> ldw{s_b2c, (const uint16_t*)my_bytes, s_cst, vmask, bitbase, bitbase - 8 * byte0,
                                                pkt_len * 8, p.nbits, m - 1, p.M, frame_id, seed};
                fft_pass<N, P::R[0], 1, 1, decltype(ldw), SmemOut32, false, P>(tid, p.tw, ldw, SmemOut32{buf});
            } else if (data) {
                TxLoadW<N, RLO, RHI, true> ldw{s_b2c, (const uint16_t*)my_bytes, s_cst, vmask, bitbase, bitbase - 8 * byte0,
                                               pkt_len * 8, p.nbits, m - 1, p.M, frame_id, seed};
                fft_pass<N, P::R[0], 1, 1, decltype(ldw), SmemOut32, false, P>(tid, p.tw, ldw, SmemOut32{buf});
            }
        }
        __syncwarp();
        if (data) fft_pass<N, P::R[1], P::R[0], 1, SmemIn32, decltype(st), false, P>(tid, p.tw, SmemIn32{buf}, st);
        __syncwarp();
    }
}

template <int N, int RLO, int RHI>
static int launch_tx_warp(ofdm_handle* h, const TxParams& p, cudaStream_t st) {
    using P = typename FftPlanW<N>::type;
    constexpr int G = 32 / (N / P::E);
    const int sym_bytes = (p.ncar * p.nbits + 7) / 8 + 2;
    size_t smem = (((size_t)p.M + 15) / 16 * 16 + (size_t)G * (FFT_PAD32(N) + 2)) * sizeof(float2) + (size_t)N * sizeof(int16_t) +
                  (size_t)G * ((sym_bytes + 15) & ~15) * 2;
    const bool taps = p.map_tap || p.pre_tap || p.ifft_tap;
    if (taps) { OFDM_SET_MAX_SMEM((tx_warp_kernel<N, true, 16, 16>), smem, h->device); } else { OFDM_SET_MAX_SMEM((tx_warp_kernel<N, false, RLO, RHI>), smem, h->device); }
    int64_t want = (p.total_syms + G - 1) / G;
    int64_t cap = (int64_t)h->sms * 64;
    int grid = (int)(want < cap ? want : cap);
    if (taps) tx_warp_kernel<N, true, 16, 16><<<grid, 32, smem, st>>>(p);
    else tx_warp_kernel<N, false, RLO, RHI><<<grid, 32, smem, st>>>(p);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}

int launch_tx(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames, int64_t first_frame,
              const int64_t* sym_off, int64_t total_syms, int32_t uniform_syms, const int64_t* stream_frame0,
              const int64_t* stream_out_off, int32_t n_streams, float2* out, cudaStream_t st, float2* map_tap,
              float2* pre_tap, float2* ifft_tap) {
    TxParams p;
    p.map_tap = map_tap; p.pre_tap = pre_tap; p.ifft_tap = ifft_tap; p.pre_freq = h->d_pre_freq; p.pre_ifft = h->d_pre_ifft;
    p.stream_frame0 = stream_frame0; p.stream_out_off = stream_out_off; p.n_streams = n_streams;
    p.pkts = pkts; p.pkt_off = pkt_off; p.sym_off = sym_off; p.n_frames = n_frames; p.uniform_syms = uniform_syms;
    p.first_frame = first_frame; p.seed = h->pad_seed; p.out = out; p.cst = h->d_const; p.bin2car = h->d_bin2car;
    p.tw = h->d_tw; p.pre_time = h->d_pre_time; p.cp = h->cp; p.ncar = h->ncar; p.nbits = h->nbits; p.M = h->M;
    p.s1 = (float)(1.0 / sqrt((double)h->N)); p.amp = h->amp;
    p.total_syms = uniform_syms > 0 ? (int64_t)n_frames * uniform_syms : total_syms;
    if (p.total_syms >= (1ll << 31)) { ofdm_set_error("tx: more than 2^31 OFDM symbols in one batch"); return OFDM_E_INVAL; }
    static const bool old_plan = getenv("OFDM_TX_OLD") != nullptr;
    if (!old_plan && h->d_tw_w) {
        p.tw = h->d_tw_w;
        // first-pass rows without a carrier: [7, 25) of 32 for every layout with occupied_tones / fft_length <= 0.41
        const bool narrow = h->tx_row_lo <= 7 && h->tx_row_hi >= 25;
        if (h->N == 512) return narrow ? launch_tx_warp<512, 7, 25>(h, p, st) : launch_tx_warp<512, 16, 16>(h, p, st);
        if (h->N == 1024) return narrow ? launch_tx_warp<1024, 7, 25>(h, p, st) : launch_tx_warp<1024, 16, 16>(h, p, st);
        p.tw = h->d_tw;
    }
    switch (h->N) {
        case 64:   return launch_tx_n<64, 8>(h, p, st);
        case 128:  return launch_tx_n<128, 8>(h, p, st);
        case 256:  return launch_tx_n<256, 8>(h, p, st);
        case 512:  return launch_tx_n<512, 2>(h, p, st);
        case 1024: return launch_tx_n<1024, 4>(h, p, st);
        case 2048: return launch_tx_n<2048, 2>(h, p, st);
        case 4096: return launch_tx_n<4096, 1>(h, p, st);
    }
    ofdm_set_error("tx: unsupported fft_length %d", h->N);
    return OFDM_E_INVAL;
}

// ---------------------------------------------------------------------------------------------
// Synthetic channel (test / bench infrastructure): CFO rotation + counter-based Gaussian noise.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// e^{j(phase0 + w i)}: the phase is reduced in float64 (i runs to 10^9), the sine / cosine taken in float32
__device__ __forceinline__ float2 channel_phasor(int64_t i, double w, double phase0) {
    double r = (phase0 + w * (double)i) * 0.15915494309189533577;        // turns
    r -= rint(r);
    float s, c;
    sincospif(2.0f * (float)r, &s, &c);
    return make_float2(c, s);
}

// complex Gaussian of standard deviation sigma per component, a pure function of (seed, sample index): Box-Muller on
// two 23-bit uniforms cut from one 64-bit hash (mantissa stuffing instead of int -> float conversions: the XU pipe is
// this kernel's bottleneck), fast intrinsics -- it is only noise
__device__ __forceinline__ float2 channel_noise(int64_t i, float sigma, uint64_t seed) {
    const uint64_t hsh = mix64(seed ^ (uint64_t)i * 0xD6E8FEB86659FD93ull);
    const float u1 = __uint_as_float(0x3F800000u | (uint32_t)(hsh >> 41)) - 0.99999994f;      // (0, 1]
    const float u2 = __uint_as_float(0x3F800000u | ((uint32_t)hsh & 0x7FFFFFu)) - 1.0f;      // [0, 1)
    const float rad = sigma * sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

__device__ __forceinline__ float2 channel_sample(float2 v, int64_t i, double w, double phase0, float sigma, uint64_t seed) {
    const float2 ph = channel_phasor(i, w, phase0);
    float2 o = make_float2(v.x * ph.x - v.y * ph.y, v.x * ph.y + v.y * ph.x);
    if (sigma > 0.f) {
        const float2 g = channel_noise(i, sigma, seed);
        o.x += g.x;
        o.y += g.y;
    }
    return o;
}

__global__ void __launch_bounds__(256) channel_kernel(const float2* __restrict__ x, int64_t n, double w, double phase0,
                                                       float sigma, uint64_t seed, float2* __restrict__ y) {
    const bool vec = ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        // two samples per thread and pass: the second phasor is the first one turned by e^{jw}
        double sw, cw;
        sincos(w, &sw, &cw);
        const float2 step = make_float2((float)cw, (float)sw);
        const int64_t n2 = n >> 1;
        for (int64_t k = t; k < n2; k += stride) {
            const float4 q = __ldg((const float4*)x + k);
            const float2 p0 = channel_phasor(2 * k, w, phase0);
            const float2 p1 = make_float2(p0.x * step.x - p0.y * step.y, p0.x * step.y + p0.y * step.x);
            float2 a = make_float2(q.x * p0.x - q.y * p0.y, q.x * p0.y + q.y * p0.x);
            float2 b = make_float2(q.z * p1.x - q.w * p1.y, q.z * p1.y + q.w * p1.x);
            if (sigma > 0.f) {
                const float2 ga = channel_noise(2 * k, sigma, seed), gb = channel_noise(2 * k + 1, sigma, seed);
                a.x += ga.x; a.y += ga.y; b.x += gb.x; b.y += gb.y;
            }
            ((float4*)y)[k] = make_float4(a.x, a.y, b.x, b.y);
        }
        if ((n & 1) && t == 0) y[n - 1] = channel_sample(x[n - 1], n - 1, w, phase0, sigma, seed);
    } else {
        for (int64_t i = t; i < n; i += stride) y[i] = channel_sample(x[i], i, w, phase0, sigma, seed);
    }
}

int launch_channel(ofdm_handle* h, const float2* x, int64_t n, float cfo, double phase0, float sigma, uint64_t seed,
                   float2* y, cudaStream_t st) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)h->sms * 32) blocks = (int64_t)h->sms * 32;
    double w = 2.0 * M_PI * (double)cfo / (double)h->N;
    channel_kernel<<<(int)blocks, 256, 0, st>>>(x, n, w, phase0, sigma, seed, y);
    OFDM_LAUNCH_CHECK();
    return OFDM_OK;
}
