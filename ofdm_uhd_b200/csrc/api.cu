// C ABI of libofdm_b200.so: handles, constant tables, argument checks, dispatch.
// Host-side table construction restates ofdm.py:71-101 (preamble / constellation wiring),
// ofdm_receiver.py~:69-76 (firdes.low_pass design inputs) and SURVEY.md A.3/A.5/A.13.
#include "internal.h"
#include "fft.cuh"
#include <algorithm>
#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include <complex>

static thread_local char g_err[512] = "";

void ofdm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* ofdm_last_error(void) { return g_err; }
extern "C" int ofdm_version(void) { return 100; }

// ---- tables ---------------------------------------------------------------------------------
static const char* KNOWN_B85 =
    "G%Yn+z+C((y^1&>9SHr$Q%^4)KPP8_N!hz@_mT_hMPvTxbqaQkWSi~Wu8B!QR)zkQ)?zPm4aH;<^;58S?h8z>gq=eL_VMpm?DBrxa2xPsU>54"
    ";i5Ru+gfA;GH$bMdh_xJhbN6kQ9)ryE|IZ?)lhgjCkOlAekgb&5m$FbNK%Pmv|0=(<v(p0LR%@q-)c0D`wL6(X^8?*2YAJeZE+~CFpi;G2eEiD"
    "Odn=*7&a>x+PX~&MIuu~7%bh%cQomBX^95JLBM0i)MRP!ug~mrlXB7BpNLLXlee6tz#FW&H?aNNkYkly63%RC$RBee%Aat=1+U^cPDrIKt|8XH"
    "S3gA?LIbFBq63Ljc%e$T{6tVZn#a|2pXi-U}YxJByY%$CtU{u*{3~>*h>h31-=c!WIGcg<p3~_<>J_8m%OCM-bl@>7_oL&`GTCGH;-X2I3?Yl"
    "}b)m2yTpd5)>DM{#Xip|l$!o3)kCK`q9U<LEsg-M(yP~uH(k4z5oFj20cdCeIh^m;Gcvohp>V6iA65@;>N%zpw7P<7^|Fg@F7nw$CRx~H&uR!"
    "4k~i5XL-hcE<7jm>uHxDL<<8tR3xCLADjp691EO9&P=s~@I7i>0L!lW5T9Zhj22RarGG8+Ti-M`_e;Xci|P-7!lnOKc(gdR8LNP`45$hx8#Hay"
    "1_Nc3F)TUXg|=B}+BSL*;>0dPw{g&!t;!18_z1Te0emdM_YY";

// known_symbols_4512_3 (ofdm.py:310-325) as +/-1, decoded from the packed base85 form
static std::vector<int> known_symbols() {
    static const char* alpha =
        "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz!#$%&()*+-;<=>?@^_`{|}~";
    int dec[256];
    for (int i = 0; i < 256; ++i) dec[i] = -1;
    for (int i = 0; i < 85; ++i) dec[(unsigned char)alpha[i]] = i;
    std::vector<uint8_t> bytes;
    size_t n = strlen(KNOWN_B85);
    for (size_t i = 0; i < n; i += 5) {
        uint64_t acc = 0;
        int cnt = 0;
        for (int j = 0; j < 5; ++j) {
            int v = (i + j < n) ? dec[(unsigned char)KNOWN_B85[i + j]] : 84;
            if (i + j < n) ++cnt;
            acc = acc * 85 + (uint64_t)v;
        }
        for (int j = 0; j < cnt - 1; ++j) bytes.push_back((uint8_t)((acc >> (24 - 8 * j)) & 0xFF));
    }
    std::vector<int> ks(4512);
    for (int i = 0; i < 4512; ++i) ks[i] = ((bytes[i >> 3] >> (i & 7)) & 1) ? 1 : -1;
    return ks;
}

// random_mask_tuple (ofdm_packet_utils.py:195-451): PN15 x^15+x^14+1, LSB-first, 4094 bytes + wrap
static void whitening_mask(uint8_t* m) {
    const int nbits = 4094 * 8;
    std::vector<uint8_t> g(nbits, 0);
    for (int i = 0; i < 14; ++i) g[i] = 1;
    for (int n = 15; n < nbits; ++n) g[n] = g[n - 14] ^ g[n - 15];
    memset(m, 0, 4096);
    for (int n = 0; n < nbits; ++n) m[n >> 3] |= (uint8_t)(g[n] << (n & 7));
    m[4094] = m[0];
    m[4095] = m[1];
}

static void crc_table(uint32_t* t) {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? ((c << 1) ^ 0x04C11DB7u) : (c << 1);
        t[i] = c;
    }
}

static std::string carrier_hex(int occ, const char* base) {
    std::string c = (base && *base) ? base : "FE7F";
    int diff = occ - 4 * (int)c.size();
    while (diff > 7) { c = "f" + c + "f"; diff -= 8; }
    if (diff > 0) {
        const char* hx = "0123456789abcdef";
        int dl = (int)ceil(diff / 2.0), dr = diff - dl;
        c = std::string(1, hx[(1 << dl) - 1]) + c + std::string(1, hx[0xF ^ ((1 << dr) - 1)]);
    }
    return c;
}

static int hexval(char ch) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    return ch - 'A' + 10;
}

// gr.firdes.low_pass(1, 1, fc, tw, WIN_HAMMING)
static int firdes_lowpass(double fc, double tw, float* taps, int max_taps) {
    int ntaps = (int)(53.0 / (22.0 * tw));
    if ((ntaps & 1) == 0) ++ntaps;
    if (ntaps > max_taps) return -1;
    int M0 = (ntaps - 1) / 2;
    std::vector<double> t(ntaps);
    double fw = 2 * M_PI * fc;
    for (int n = -M0; n <= M0; ++n) {
        double w = 0.54 - 0.46 * cos(2 * M_PI * (n + M0) / (ntaps - 1));
        t[n + M0] = (n == 0) ? fw / M_PI * w : sin(n * fw) / (n * M_PI) * w;
    }
    double fmax = t[M0];
    for (int n = 1; n <= M0; ++n) fmax += 2 * t[n + M0];
    for (int n = 0; n < ntaps; ++n) taps[n] = (float)(t[n] / fmax);
    return ntaps;
}

template <class T> static int upload(T** dptr, const std::vector<T>& v) {
    OFDM_CUDA_CHECK(cudaMalloc((void**)dptr, v.size() * sizeof(T)));
    OFDM_CUDA_CHECK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// per-pass, thread-ordered twiddle sections (fft.cuh: fft_fill_twiddles)
static std::vector<float2> twiddles(int N) {
    std::vector<float2> tw((size_t)fft_twiddle_elems(N), make_float2(0.f, 0.f));
    fft_fill_twiddles_n(N, tw.data());
    return tw;
}

static bool fft_size_ok(int N) { return N == 64 || N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096; }

extern "C" ofdm_handle* ofdm_create(const ofdm_cfg* cfg) {
    if (!cfg || !cfg->host_constellation) { ofdm_set_error("ofdm_create: null cfg"); return nullptr; }
    const int N = cfg->fft_length, occ = cfg->occupied_tones, cp = cfg->cp_length, M = cfg->constellation_size;
    if (!fft_size_ok(N)) { ofdm_set_error("ofdm_create: fft_length %d not in {64..4096 powers of two}", N); return nullptr; }
    if (occ > N) { ofdm_set_error("ofdm_create: occupied_tones > fft_length"); return nullptr; }   // upstream invalid_argument
    if (occ < 16 || cp < 1 || cp > N) { ofdm_set_error("ofdm_create: bad occupied_tones/cp_length"); return nullptr; }
    int nbits = 0;
    while ((1 << nbits) < M) ++nbits;
    if ((1 << nbits) != M || nbits < 1 || nbits > 8) { ofdm_set_error("ofdm_create: constellation size %d", M); return nullptr; }
    if (cudaSetDevice(cfg->device) != cudaSuccess) { ofdm_set_error("ofdm_create: cudaSetDevice(%d) failed", cfg->device); return nullptr; }

    ofdm_handle* h = new ofdm_handle();
    memset(h, 0, sizeof(*h));
    h->device = cfg->device;
    h->sms = 148;
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, cfg->device);
    h->N = N; h->occ = occ; h->cp = cp; h->L = N + cp; h->M = M; h->nbits = nbits;
    h->zl = (int)ceil((N - occ) / 2.0);
    h->amp = fmaxf(0.f, fminf(cfg->tx_amplitude, 1.f));
    h->pad_seed = cfg->pad_seed;
    int mp = cfg->max_pkt_bytes > 0 ? cfg->max_pkt_bytes : 4096;
    if (mp > 4096) mp = 4096;
    h->pkt_stride = (mp + 15) & ~15;

    // carrier maps (A.3)
    if (cfg->host_carrier_map)
        for (const char* q = cfg->host_carrier_map; *q; ++q)
            if (!isxdigit((unsigned char)*q)) { ofdm_set_error("ofdm_create: carrier map is not a hex string"); delete h; return nullptr; }
    std::string hx = carrier_hex(occ, cfg->host_carrier_map);
    if ((int)hx.size() * 4 > occ + 3) { ofdm_set_error("ofdm_create: carrier map wider than occupied_tones"); delete h; return nullptr; }
    int pad = (N / 4 - (int)hx.size()) / 2;
    std::vector<int16_t> bin2car(N, -1), sinkmap;
    int ord = 0;
    for (size_t i = 0; i < hx.size(); ++i)
        for (int j = 0; j < 4; ++j)
            if ((hexval(hx[i]) >> (3 - j)) & 1) {
                int v = 4 * ((int)i + pad) + j;
                if (v < 0 || v >= N) { ofdm_set_error("ofdm_create: carrier map outside the FFT"); delete h; return nullptr; }
                bin2car[v] = (int16_t)ord++;
                sinkmap.push_back((int16_t)(4 * i + j));
            }
    h->ncar = ord;
    {   // the longest run of 32nd-of-the-band rows of the IFFT input without a carrier, around row 16 (the band edges)
        bool used[32] = {false};
        if (N >= 32) for (int v = 0; v < N; ++v) if (bin2car[v] >= 0) used[(((v + N / 2) % N) * 32) / N] = true;
        h->tx_row_lo = 16; h->tx_row_hi = 16;
        while (h->tx_row_lo > 0 && !used[h->tx_row_lo - 1]) --h->tx_row_lo;
        while (h->tx_row_hi < 32 && !used[h->tx_row_hi]) ++h->tx_row_hi;
        if (used[16] && h->tx_row_hi == 16) h->tx_row_lo = 16;
    }
    if (h->ncar * nbits < 32) { ofdm_set_error("ofdm_create: fewer than 32 bits per OFDM symbol"); delete h; return nullptr; }
    if (h->zl < OFDM_MAX_SHIFT || h->zl + occ + OFDM_MAX_SHIFT + 2 > N) {
        ofdm_set_error("ofdm_create: layout leaves no room for the +/-%d bin coarse search", OFDM_MAX_SHIFT); delete h; return nullptr;
    }

    // known symbol (ofdm.py:71-77) and its difference profile (A.10)
    std::vector<int> known = known_symbols();
    std::vector<float> ks(occ), kd(occ, 0.f);
    for (int i = 0; i < occ; ++i) ks[i] = ((h->zl + i) & 1) ? 0.f : (float)known[i];
    for (int i = 0; i + 2 < occ; i += 2) { float d = ks[i] - ks[i + 2]; kd[i] = d * d; }

    // channel filter (ofdm_receiver.py~:69-76)
    double bw = ((double)occ / (double)N) / 2.0, tb = bw * 0.08;
    h->ntaps = firdes_lowpass(bw + tb, tb, h->h_taps, OFDM_MAX_TAPS);
    if (h->ntaps < 0) { ofdm_set_error("ofdm_create: channel filter needs more than %d taps", OFDM_MAX_TAPS); delete h; return nullptr; }
    // overlap-save size: 1024 (the warp-per-block kernel, 2 radix-32 passes) while at least 3/4 of a block is output
    h->NOS = 1024;
    if (const char* e = getenv("OFDM_FILTER_NOS")) { if (atoi(e) == 2048 || atoi(e) == 4096) h->NOS = atoi(e); }
    while (h->NOS < 4 * h->ntaps) h->NOS *= 2;
    if (h->NOS > 4096) { ofdm_set_error("ofdm_create: channel filter too long"); delete h; return nullptr; }
    std::vector<float2> Hos(h->NOS);
    for (int k = 0; k < h->NOS; ++k) {
        std::complex<double> acc = 0;
        for (int t = 0; t < h->ntaps; ++t)
            acc += (double)h->h_taps[t] * std::polar(1.0, -2 * M_PI * (double)((long)k * t % h->NOS) / h->NOS);
        Hos[k] = make_float2((float)(acc.real() / h->NOS), (float)(acc.imag() / h->NOS));
    }

    // time-domain preamble: N*ifft(ifftshift(X)), CP, * 1/sqrt(N) (float32 after each stage, A.4)
    std::vector<float2> pre(N + cp), pre_ifft(N), pre_freq(N, make_float2(0.f, 0.f));
    {
        std::vector<std::complex<double>> x(N);
        for (int n = 0; n < N; ++n) {
            std::complex<double> acc = 0;
            for (int i = 0; i < occ; ++i) {
                if (ks[i] == 0.f) continue;
                long k = ((long)(h->zl + i) + N / 2) % N;       // ifftshift: in[k] = X[(k+N/2)%N]
                acc += (double)ks[i] * std::polar(1.0, 2 * M_PI * (double)(k * n % N) / N);
            }
            x[n] = acc;
        }
        const float s1 = (float)(1.0 / sqrt((double)N));
        for (int n = 0; n < N + cp; ++n) {
            std::complex<double> v = x[(n + N - cp) % N];
            pre[n] = make_float2((float)v.real() * s1, (float)v.imag() * s1);
        }
        for (int n = 0; n < N; ++n) pre_ifft[n] = make_float2((float)x[n].real(), (float)x[n].imag());
        for (int i = 0; i < occ; ++i) pre_freq[h->zl + i] = make_float2(ks[i], 0.f);
    }

    std::vector<float2> cst(M);
    for (int i = 0; i < M; ++i) cst[i] = make_float2(cfg->host_constellation[2 * i], cfg->host_constellation[2 * i + 1]);
    // A constellation whose points form a full L x L grid of (nearly) uniform levels -- qam.constellation for
    // m >= 64 -- lets the frame sink's slicer look at the 3 x 3 cells around the received point instead of all M
    // points; the result is provably the brute-force one (rx_demod.cu: slice_point).
    std::vector<uint8_t> grid;
    h->grid_L = 0;
    if (M >= 64) {
        std::vector<float> xs, ys;
        for (int i = 0; i < M; ++i) { xs.push_back(cst[i].x); ys.push_back(cst[i].y); }
        std::sort(xs.begin(), xs.end()); xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
        std::sort(ys.begin(), ys.end()); ys.erase(std::unique(ys.begin(), ys.end()), ys.end());
        const int Lg = (int)xs.size();
        bool ok = Lg >= 8 && (int)ys.size() == Lg && Lg * Lg == M;
        double dx = 0, dy = 0;
        if (ok) {
            dx = ((double)xs[Lg - 1] - xs[0]) / (Lg - 1);
            dy = ((double)ys[Lg - 1] - ys[0]) / (Lg - 1);
            ok = dx > 0 && dy > 0;
            for (int i = 0; ok && i < Lg; ++i)
                ok = fabs(xs[i] - (xs[0] + i * dx)) < 1e-4 * dx && fabs(ys[i] - (ys[0] + i * dy)) < 1e-4 * dy;
            // the candidate-search argument below assumes |coordinates| <= 1 (levels normalised by the largest)
            ok = ok && fabs(xs[0]) <= 1.0001 && fabs(xs[Lg - 1]) <= 1.0001 && fabs(ys[0]) <= 1.0001 && fabs(ys[Lg - 1]) <= 1.0001;
        }
        if (ok) {
            grid.assign((size_t)M, 255);
            std::vector<int> seen((size_t)M, 0);
            for (int i = 0; ok && i < M; ++i) {
                const int ix = (int)(std::lower_bound(xs.begin(), xs.end(), cst[i].x) - xs.begin());
                const int iy = (int)(std::lower_bound(ys.begin(), ys.end(), cst[i].y) - ys.begin());
                if (seen[(size_t)iy * Lg + ix]++) ok = false;
                grid[(size_t)iy * Lg + ix] = (uint8_t)i;
            }
        }
        if (ok) {
            h->grid_L = Lg;
            h->grid_x0 = xs[0]; h->grid_y0 = ys[0];
            h->grid_inv_dx = (float)(1.0 / dx); h->grid_inv_dy = (float)(1.0 / dy);
        }
    }
    std::vector<uint8_t> mask(4096);
    whitening_mask(mask.data());
    // CRC table, then x^(8*128*j) mod P (j < 32) and x^(8*r) mod P (r < 128): the shift factors with which the
    // warp-per-packet CRC kernel combines the partial CRCs of 128-byte slices
    std::vector<uint32_t> crct(OFDM_CRC_SLICE + 3 * 256);
    crc_table(crct.data());
    // slicing tables: T_k[i] = state after byte i and k zero bytes (four bytes per step in the framing / CRC kernels)
    for (int k = 1; k < 4; ++k)
        for (int i = 0; i < 256; ++i) {
            const uint32_t v = crct[(k == 1 ? 0 : OFDM_CRC_SLICE + (k - 2) * 256) + i];
            crct[OFDM_CRC_SLICE + (k - 1) * 256 + i] = crct[v >> 24] ^ (v << 8);
        }
    {
        uint32_t s = 1u;                                            // the polynomial 1 = x^0
        for (int r = 0; r <= 128 * 31; ++r) {
            if (r < 128) crct[256 + 32 + r] = s;
            if ((r & 127) == 0) crct[256 + (r >> 7)] = s;
            s = crct[s >> 24] ^ (s << 8);                           // times x^8 mod P (one zero byte)
        }
    }

    int rc = 0;
    rc |= upload(&h->d_const, cst);
    rc |= upload(&h->d_bin2car, bin2car);
    rc |= upload(&h->d_sinkmap, sinkmap);
    rc |= upload(&h->d_ks, ks);
    rc |= upload(&h->d_kd, kd);
    rc |= upload(&h->d_tw, twiddles(N));
    if (N == 512 || N == 1024) {
        std::vector<float2> tw((size_t)fft_twiddle_elems(N), make_float2(0.f, 0.f));
        if (N == 512) fft_fill_twiddles<512, FftPlanW512>(tw.data()); else fft_fill_twiddles<1024, FftPlanW1024>(tw.data());
        rc |= upload(&h->d_tw_w, tw);
    }
    if (h->NOS == 1024) {
        std::vector<float2> tw((size_t)fft_twiddle_elems(1024), make_float2(0.f, 0.f));
        fft_fill_twiddles<1024, FftPlanW1024>(tw.data());
        rc |= upload(&h->d_tw_os, tw);
    } else rc |= upload(&h->d_tw_os, twiddles(h->NOS));
    rc |= upload(&h->d_Hos, Hos);
    rc |= upload(&h->d_pre_time, pre);
    rc |= upload(&h->d_pre_freq, pre_freq);
    rc |= upload(&h->d_pre_ifft, pre_ifft);
    rc |= upload(&h->d_mask, mask);
    rc |= upload(&h->d_crctab, crct);
    if (h->grid_L) rc |= upload(&h->d_grid, grid);
    if (rc) { ofdm_destroy(h); return nullptr; }
    return h;
}

extern "C" void ofdm_destroy(ofdm_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_const); cudaFree(h->d_bin2car); cudaFree(h->d_sinkmap); cudaFree(h->d_ks); cudaFree(h->d_kd);
    cudaFree(h->d_tw); cudaFree(h->d_tw_w); cudaFree(h->d_tw_os); cudaFree(h->d_Hos); cudaFree(h->d_pre_time); cudaFree(h->d_pre_freq); cudaFree(h->d_pre_ifft); cudaFree(h->d_mask);
    cudaFree(h->d_crctab); cudaFree(h->d_grid); cudaFree(h->d_Hks_half); cudaFree(h->d_Hks_full); cudaFree(h->d_tw_os_alt[0]); cudaFree(h->d_tw_os_alt[1]);
    delete h;
}

extern "C" int ofdm_set_tx_amplitude(ofdm_handle* h, float ampl) {
    if (!h) return OFDM_E_INVAL;
    h->amp = fmaxf(0.f, fminf(ampl, 1.f));
    return OFDM_OK;
}

extern "C" int ofdm_get_layout(const ofdm_handle* h, int32_t* o) {
    if (!h || !o) return OFDM_E_INVAL;
    o[0] = h->zl; o[1] = h->ncar; o[2] = h->nbits; o[3] = h->L; o[4] = h->ntaps; o[5] = h->NOS; o[6] = h->pkt_stride; o[7] = 0;
    return OFDM_OK;
}

extern "C" int ofdm_get_chan_taps(const ofdm_handle* h, float* t, int32_t max_taps) {
    if (!h || !t || max_taps < h->ntaps) return OFDM_E_INVAL;
    memcpy(t, h->h_taps, sizeof(float) * h->ntaps);
    return h->ntaps;
}

extern "C" int32_t ofdm_packet_len(int32_t payload_len, int pad_for_usrp) {
    int32_t len = 4 + payload_len + 4 + 1;
    if (pad_for_usrp) len = (len + 15) & ~15;      // _npadding_bytes with sps = bps = 1 (ofdm.py:143)
    return len;
}

extern "C" int32_t ofdm_frame_symbols(const ofdm_handle* h, int32_t pkt_len) {
    if (!h) return OFDM_E_INVAL;
    int64_t bits = 8ll * pkt_len, per = (int64_t)h->ncar * h->nbits;
    int32_t nd = pkt_len > 0 ? (int32_t)((bits + per - 1) / per) : 1;
    return 1 + (nd < 1 ? 1 : nd);
}

#define NEED(h) do { if (!(h)) { ofdm_set_error("null handle"); return OFDM_E_INVAL; } \
                     OFDM_CUDA_CHECK(cudaSetDevice((h)->device)); } while (0)

extern "C" int ofdm_make_packets(ofdm_handle* h, const uint8_t* payload, const int64_t* payload_off, int32_t n_pkts,
                                 int whitening, uint8_t* pkts, const int64_t* pkt_off, void* stream) {
    NEED(h);
    if (n_pkts <= 0) return OFDM_OK;
    return launch_make_packets(h, payload, payload_off, n_pkts, whitening, pkts, pkt_off, (cudaStream_t)stream);
}

extern "C" int ofdm_tx_modulate_batch(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames,
                                      int64_t first_frame, const int64_t* sym_off, int64_t total_syms,
                                      int32_t uniform_syms, float* out_iq, void* stream) {
    NEED(h);
    if (n_frames <= 0) return OFDM_OK;
    if (uniform_syms <= 0 && !sym_off) { ofdm_set_error("tx: sym_off required for ragged frames"); return OFDM_E_INVAL; }
    return launch_tx(h, pkts, pkt_off, n_frames, first_frame, sym_off, total_syms, uniform_syms, nullptr, nullptr, 0,
                     (float2*)out_iq, (cudaStream_t)stream);
}

extern "C" int ofdm_tx_modulate_taps(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames,
                                     int64_t first_frame, const int64_t* sym_off, int64_t total_syms, int32_t uniform_syms,
                                     float* out_iq, float* mapper_out, float* preambles_out, float* ifft_out, void* stream) {
    NEED(h);
    if (n_frames <= 0) return OFDM_OK;
    if (uniform_syms <= 0 && !sym_off) { ofdm_set_error("tx: sym_off required for ragged frames"); return OFDM_E_INVAL; }
    return launch_tx(h, pkts, pkt_off, n_frames, first_frame, sym_off, total_syms, uniform_syms, nullptr, nullptr, 0,
                     (float2*)out_iq, (cudaStream_t)stream, (float2*)mapper_out, (float2*)preambles_out, (float2*)ifft_out);
}

extern "C" int ofdm_tx_modulate_streams(ofdm_handle* h, const uint8_t* pkts, const int64_t* pkt_off, int32_t n_frames,
                                        int64_t first_frame, const int64_t* sym_off, int64_t total_syms, int32_t uniform_syms,
                                        const int64_t* stream_frame0, const int64_t* stream_out_off, int32_t n_streams,
                                        float* out_iq, void* stream) {
    NEED(h);
    if (n_frames <= 0) return OFDM_OK;
    if (uniform_syms <= 0 && !sym_off) { ofdm_set_error("tx: sym_off required for ragged frames"); return OFDM_E_INVAL; }
    if (n_streams < 1 || !stream_frame0 || !stream_out_off) { ofdm_set_error("tx streams: bad stream tables"); return OFDM_E_INVAL; }
    return launch_tx(h, pkts, pkt_off, n_frames, first_frame, sym_off, total_syms, uniform_syms, stream_frame0, stream_out_off,
                     n_streams, (float2*)out_iq, (cudaStream_t)stream);
}

extern "C" size_t ofdm_rx_workspace_bytes(const ofdm_handle* h, int64_t n, int32_t max_frames) {
    if (!h) return 0;
    size_t need = 0;
    RxWorkspace ws;
    rx_workspace_layout(h, single_stream(n), max_frames, nullptr, 0, &ws, &need);
    return need;
}

extern "C" size_t ofdm_rx_workspace_bytes_batch(const ofdm_handle* h, int32_t n_streams, int64_t total_samples,
                                                int64_t max_stream_samples, int32_t max_frames_per_stream) {
    if (!h || n_streams < 1) return 0;
    StreamSet ss;
    ss.S = n_streams; ss.off = nullptr; ss.n_max = max_stream_samples; ss.n_total = total_samples;
    size_t need = 0;
    RxWorkspace ws;
    rx_workspace_layout(h, ss, max_frames_per_stream, nullptr, 0, &ws, &need);
    return need;
}

static int get_ws(ofdm_handle* h, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws) {
    if (!io || !io->workspace) { ofdm_set_error("rx: null io/workspace"); return OFDM_E_INVAL; }
    size_t need = 0;
    int rc = rx_workspace_layout(h, ss, io->max_frames, io->workspace, io->workspace_bytes, ws, &need);
    if (rc) ofdm_set_error("rx: workspace too small (%zu < %zu)", io->workspace_bytes, need);
    return rc;
}

extern "C" void* ofdm_rx_workspace_ptr(const ofdm_handle* h, const ofdm_rx_io* io, int64_t n, int which) {
    if (!h || !io) return nullptr;
    RxWorkspace ws;
    size_t need = 0;
    if (rx_workspace_layout(h, single_stream(n), io->max_frames, io->workspace, io->workspace_bytes, &ws, &need)) return nullptr;
    switch (which) {
        case 0: return ws.y;
        case 1: return ws.mf;
        case 2: return ws.first_nan;
        case 3: return ws.phi0;
        case 4: return ws.step;
        case 5: return ws.vbase;
        case 6: return ws.sess_nvec;
        case 7: return ws.n_nco;
        case 8: return ws.nco_idx;
        case 9: return ws.nco_ang;
        default: return nullptr;
    }
}

extern "C" int ofdm_rx_chan_filter(ofdm_handle* h, const float* x, int64_t n, float* y, void* stream) {
    NEED(h);
    if (n <= 0) return OFDM_OK;
    return launch_chan_filter(h, (const float2*)x, single_stream(n), (float2*)y, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_sync_metric(ofdm_handle* h, const float* y, int64_t n, float* mf, int64_t* first_nan, void* stream) {
    NEED(h);
    return launch_sync_metric(h, (const float2*)y, n, mf, first_nan, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_peak_detect(ofdm_handle* h, const float* y, const float* mf, int64_t n, const int64_t* first_nan,
                                   ofdm_rx_io* io, void* stream) {
    NEED(h);
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(n), io, &ws);
    if (rc) return rc;
    OFDM_CUDA_CHECK(cudaMemsetAsync(ws.nco_init, 0, sizeof(double), (cudaStream_t)stream));
    if ((rc = launch_nco_mode(h, &ws, 1, -2.0, (cudaStream_t)stream))) return rc;
    return launch_peak_detect(h, (const float2*)y, mf, n, first_nan, io, &ws, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_plan(ofdm_handle* h, int64_t n, ofdm_rx_io* io, void* stream) {
    NEED(h);
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(n), io, &ws);
    if (rc) return rc;
    return launch_plan(h, single_stream(n), io, &ws, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_stage(ofdm_handle* h, const float* y, int64_t n, ofdm_rx_io* io, int32_t stage, void* stream) {
    NEED(h);
    RxWorkspace ws;
    const StreamSet ss = single_stream(n);
    int rc = get_ws(h, ss, io, &ws);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (stage) {
        case 0: rc = launch_sync_stream(h, (const float2*)y, ss, io, &ws, 1, st); break;
        case 1: rc = launch_sync_stream(h, (const float2*)y, ss, io, &ws, 2, st); break;
        case 2: return launch_trig_compact(h, (const float2*)y, ss, io, &ws, st);
        case 3: return launch_demod(h, (const float2*)y, ss, io, &ws, st, 1);
        case 4: return launch_demod(h, (const float2*)y, ss, io, &ws, st, 2);
        default: ofdm_set_error("ofdm_rx_stage: stage %d", stage); return OFDM_E_INVAL;
    }
    if (rc == 1) { ofdm_set_error("ofdm_rx_stage: this layout has no streaming synchroniser"); return OFDM_E_INVAL; }
    return rc;
}

extern "C" int ofdm_rx_demod(ofdm_handle* h, const float* y, int64_t n, ofdm_rx_io* io, void* stream) {
    NEED(h);
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(n), io, &ws);
    if (rc) return rc;
    return launch_demod(h, (const float2*)y, single_stream(n), io, &ws, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_finish(ofdm_handle* h, ofdm_rx_io* io, void* stream) {
    NEED(h);
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(0), io, &ws);
    if (rc) return rc;
    return launch_finish(h, 1, io, &ws, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_liveness(const int32_t* n_frames, const int64_t* vbase, const int32_t* sess_nvec, int32_t max_frames,
                                int32_t* scratch, uint8_t* live, int force_general, void* stream) {
    if (!n_frames || !vbase || !sess_nvec || !scratch || !live || max_frames < 1) {
        ofdm_set_error("ofdm_rx_liveness: null argument or max_frames < 1");
        return OFDM_E_INVAL;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return launch_liveness(sms, 1, n_frames, vbase, (int64_t)max_frames + 1, sess_nvec, max_frames, scratch, scratch + max_frames,
                           scratch + 2 * max_frames, live, force_general, (cudaStream_t)stream);
}

// ofdm_sync_pn: the streaming kernel pair where the layout allows it (N >= 128, cp <= N/2), else -- single stream
// only -- the tile-parallel metric kernel followed by the detector kernel
static int rx_sync(ofdm_handle* h, const float2* y, const StreamSet& ss, ofdm_rx_io* io, RxWorkspace* ws, int force_fused,
                   cudaStream_t st) {
    OFDM_CUDA_CHECK(cudaMemsetAsync(io->status, 0, sizeof(uint32_t) * ss.S, st));
    OFDM_CUDA_CHECK(cudaMemsetAsync(ws->nco_init, 0, sizeof(double) * ss.S, st));     // sample_and_hold starts at 0
    { int rc0 = launch_nco_mode(h, ws, ss.S, -2.0, st); if (rc0) return rc0; }        // the NCO follows the triggers, -2/N
    (void)force_fused;
    int rc = launch_sync_stream(h, y, ss, io, ws, 3, st);
    if (rc == 0) return launch_trig_compact(h, y, ss, io, ws, st);
    if (rc < 0) return rc;
    if (ss.S > 1 || ss.off) {
        ofdm_set_error("rx batch: this layout (fft_length %d, cp_length %d) has no streaming synchroniser; feed its streams one by one", h->N, h->cp);
        return OFDM_E_INVAL;
    }
    if ((rc = launch_sync_metric(h, y, ss.n_max, ws->mf, ws->first_nan, st))) return rc;
    return launch_peak_detect(h, y, ws->mf, ss.n_max, ws->first_nan, io, ws, st);
}

extern "C" int ofdm_rx_sync(ofdm_handle* h, const float* y, int64_t n, ofdm_rx_io* io, void* stream) {
    NEED(h);
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(n), io, &ws);
    if (rc) return rc;
    return rx_sync(h, (const float2*)y, single_stream(n), io, &ws, 1, (cudaStream_t)stream);
}

static int rx_chain(ofdm_handle* h, const float2* x, const StreamSet& ss, ofdm_rx_io* io, cudaStream_t st) {
    RxWorkspace ws;
    int rc = get_ws(h, ss, io, &ws);
    if (rc) return rc;
    if ((rc = launch_chan_filter(h, x, ss, ws.y, st))) return rc;
    if ((rc = rx_sync(h, ws.y, ss, io, &ws, 0, st))) return rc;
    if ((rc = launch_plan(h, ss, io, &ws, st))) return rc;
    if ((rc = launch_demod(h, ws.y, ss, io, &ws, st))) return rc;
    return launch_finish(h, ss.S, io, &ws, st);
}

extern "C" int ofdm_rx_demodulate(ofdm_handle* h, const float* x, int64_t n, ofdm_rx_io* io, void* stream) {
    NEED(h);
    return rx_chain(h, (const float2*)x, single_stream(n), io, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_demodulate_batch(ofdm_handle* h, const float* x, const int64_t* stream_off, int32_t n_streams,
                                        int64_t total_samples, int64_t max_stream_samples, ofdm_rx_io* io, void* stream) {
    NEED(h);
    if (n_streams < 1 || !stream_off || total_samples < 0 || max_stream_samples < 0 || max_stream_samples > total_samples) {
        ofdm_set_error("ofdm_rx_demodulate_batch: bad stream set");
        return OFDM_E_INVAL;
    }
    if (n_streams > 65535) { ofdm_set_error("ofdm_rx_demodulate_batch: more than 65535 streams in one call"); return OFDM_E_INVAL; }
    StreamSet ss;
    ss.S = n_streams; ss.off = stream_off; ss.n_max = max_stream_samples; ss.n_total = total_samples;
    return rx_chain(h, (const float2*)x, ss, io, (cudaStream_t)stream);
}

extern "C" size_t ofdm_rx_sync_alt_scratch_bytes(const ofdm_handle* h, int64_t n) {
    (void)h;
    return sync_alt_scratch_bytes(n);
}

static int alt_mode(const char* sync) {
    if (sync && !strcmp(sync, "pnac")) return 1;
    if (sync && !strcmp(sync, "ml")) return 2;
    ofdm_set_error("sync %s: this entry point serves \"pnac\" and \"ml\" (\"pn\": ofdm_rx_sync, \"fixed\": ofdm_rx_sync_fixed)", sync ? sync : "(null)");
    return 0;
}

extern "C" int ofdm_rx_sync_alt(ofdm_handle* h, const float* y, int64_t n, const char* host_sync, float snr_db, ofdm_rx_io* io,
                                void* scratch, size_t scratch_bytes, void* stream) {
    NEED(h);
    const int mode = alt_mode(host_sync);
    if (!mode) return OFDM_E_INVAL;
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(n), io, &ws);
    if (rc) return rc;
    return launch_sync_alt(h, (const float2*)y, n, mode, snr_db, io, &ws, scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_demodulate_alt(ofdm_handle* h, const float* x, int64_t n, const char* host_sync, float snr_db, ofdm_rx_io* io,
                                      void* scratch, size_t scratch_bytes, void* stream) {
    NEED(h);
    const int mode = alt_mode(host_sync);
    if (!mode) return OFDM_E_INVAL;
    RxWorkspace ws;
    const StreamSet ss = single_stream(n);
    int rc = get_ws(h, ss, io, &ws);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = launch_chan_filter(h, (const float2*)x, ss, ws.y, st))) return rc;
    if ((rc = launch_sync_alt(h, ws.y, n, mode, snr_db, io, &ws, scratch, scratch_bytes, st))) return rc;
    if ((rc = launch_plan(h, ss, io, &ws, st))) return rc;
    if ((rc = launch_demod(h, ws.y, ss, io, &ws, st))) return rc;
    return launch_finish(h, 1, io, &ws, st);
}

extern "C" int ofdm_rx_sync_fixed(ofdm_handle* h, int64_t n, int32_t nsymbols, float freq_offset, ofdm_rx_io* io,
                                  void* stream) {
    NEED(h);
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(n), io, &ws);
    if (rc) return rc;
    if ((rc = launch_nco_mode(h, &ws, 1, -2.0, (cudaStream_t)stream))) return rc;
    return launch_sync_fixed(h, n, nsymbols, freq_offset, io, &ws, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_demodulate_fixed(ofdm_handle* h, const float* x, int64_t n, int32_t nsymbols, float freq_offset,
                                        ofdm_rx_io* io, void* stream) {
    NEED(h);
    RxWorkspace ws;
    const StreamSet ss = single_stream(n);
    int rc = get_ws(h, ss, io, &ws);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // chan_filt = gr.multiply_const_cc(1.0): the capture itself is what the sampler reads
    if ((rc = launch_nco_mode(h, &ws, 1, -2.0, st))) return rc;
    if ((rc = launch_sync_fixed(h, n, nsymbols, freq_offset, io, &ws, st))) return rc;
    if ((rc = launch_plan(h, ss, io, &ws, st))) return rc;
    if ((rc = launch_demod(h, (const float2*)x, ss, io, &ws, st))) return rc;
    return launch_finish(h, 1, io, &ws, st);
}

extern "C" int ofdm_rx_nco_taps(ofdm_handle* h, const float* y, int64_t n, ofdm_rx_io* io, float* nco_out, float* sigmix_out,
                                void* stream) {
    NEED(h);
    RxWorkspace ws;
    int rc = get_ws(h, single_stream(n), io, &ws);
    if (rc) return rc;
    return launch_nco_taps(h, (const float2*)y, n, io, &ws, (float2*)nco_out, (float2*)sigmix_out, (cudaStream_t)stream);
}

extern "C" int ofdm_rx_compact(ofdm_handle* h, const ofdm_rx_io* io, int32_t n_streams, uint8_t* out_bytes,
                               int64_t out_capacity, int64_t* msg_off, int32_t* msg_frame, uint32_t* ok_bits, int64_t* totals,
                               int64_t* scratch, void* stream) {
    NEED(h);
    if (!io || n_streams < 1 || !out_bytes || !msg_off || !msg_frame || !ok_bits || !totals || !scratch || out_capacity < 0) {
        ofdm_set_error("ofdm_rx_compact: null argument");
        return OFDM_E_INVAL;
    }
    return launch_compact(h, io, n_streams, out_bytes, out_capacity, msg_off, msg_frame, ok_bits, totals, scratch,
                          (cudaStream_t)stream);
}

extern "C" int ofdm_channel(ofdm_handle* h, const float* x, int64_t n, float cfo, double phase0, float sigma,
                            uint64_t seed, float* y, void* stream) {
    NEED(h);
    if (n <= 0) return OFDM_OK;
    return launch_channel(h, (const float2*)x, n, cfo, phase0, sigma, seed, (float2*)y, (cudaStream_t)stream);
}

// ---- sensing ----------------------------------------------------------------------------------
extern "C" ofdm_sense_handle* ofdm_sense_create(int32_t N, int32_t device) {
    if (!fft_size_ok(N)) { ofdm_set_error("ofdm_sense_create: fft_size %d unsupported", N); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { ofdm_set_error("ofdm_sense_create: cudaSetDevice failed"); return nullptr; }
    ofdm_sense_handle* s = new ofdm_sense_handle();
    memset(s, 0, sizeof(*s));
    s->device = device;
    s->sms = 148;
    cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device);
    s->N = N;
    std::vector<float> w(N);
    for (int i = 0; i < N; ++i) {                       // gnuradio window.blackmanharris (A.13)
        double a = 2 * M_PI * (i + 0.5) / (N - 1);
        w[i] = (float)(0.35875 - 0.48829 * cos(a) + 0.14128 * cos(2 * a) - 0.01168 * cos(3 * a));
    }
    if (upload(&s->d_win, w) || upload(&s->d_tw, twiddles(N))) { ofdm_sense_destroy(s); return nullptr; }
    return s;
}

extern "C" void ofdm_sense_destroy(ofdm_sense_handle* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->d_win); cudaFree(s->d_tw);
    delete s;
}

extern "C" int ofdm_sense(ofdm_sense_handle* s, const float* x, int64_t n_frames, int shift, int32_t tune_delay,
                          int32_t dwell_delay, float* maxhold, void* stream) {
    NEED(s);
    if (tune_delay < 0 || dwell_delay < 1) { ofdm_set_error("ofdm_sense: bad delays"); return OFDM_E_INVAL; }
    return launch_sense(s, (const float2*)x, n_frames, shift, tune_delay, dwell_delay, maxhold, nullptr, (cudaStream_t)stream);
}

extern "C" int ofdm_sense_fft(ofdm_sense_handle* s, const float* x, int64_t n_frames, int shift, float* out, void* stream) {
    NEED(s);
    return launch_sense(s, (const float2*)x, n_frames, shift, 0, 1, nullptr, (float2*)out, (cudaStream_t)stream);
}

extern "C" int ofdm_sense_decide(ofdm_sense_handle* s, const float* maxhold, int32_t n_avg, double threshold,
                                 double* avg_inorder, uint8_t* free_bits, char* hex, void* stream) {
    NEED(s);
    if (n_avg < 1) return OFDM_E_INVAL;
    return launch_sense_decide(s, maxhold, n_avg, threshold, avg_inorder, free_bits, hex, (cudaStream_t)stream);
}

extern "C" int ofdm_sense_hop(ofdm_sense_handle* s, const double* avg_inorder, const uint8_t* free_bits,
                              int32_t required_index, int32_t* out, void* stream) {
    NEED(s);
    if (!avg_inorder || !free_bits || !out) {
        ofdm_set_error("ofdm_sense_hop: null argument");
        return OFDM_E_INVAL;
    }
    return launch_sense_hop(s, avg_inorder, free_bits, required_index, out, (cudaStream_t)stream);
}
