/* C port of oracle/ofdm_oracle.py -- the CPU restatement of the reference's OFDM hot path
 * (GNU Radio 3.6 blocks as wired by /root/reference/ofdm.py:62-118,202-261 and ofdm_receiver.py~:69-142;
 * SURVEY.md Appendix A), used as the multi-threaded CPU baseline of bench.py.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY: nothing in ofdm_uhd_b200/ links or calls this file.
 * PARITY UNPINNED against upstream binaries (see the header of ofdm_oracle.py); this port is validated
 * against the NumPy oracle by tests/test_c_port.py (decoded bytes, CRC verdicts, trigger indices exact).
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: every float op is rounded individually, like the oracle).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { float re, im; } cf;
typedef struct { double re, im; } cd;

typedef struct {
    int32_t N, occ, cp, M;
    const float* constellation;      /* 2*M */
    float amp;
    uint64_t pad_seed;
} oc_cfg;

/* ---------------------------------------------------------------- tables */
static const char* KNOWN_B85 =
    "G%Yn+z+C((y^1&>9SHr$Q%^4)KPP8_N!hz@_mT_hMPvTxbqaQkWSi~Wu8B!QR)zkQ)?zPm4aH;<^;58S?h8z>gq=eL_VMpm?DBrxa2xPsU>54"
    ";i5Ru+gfA;GH$bMdh_xJhbN6kQ9)ryE|IZ?)lhgjCkOlAekgb&5m$FbNK%Pmv|0=(<v(p0LR%@q-)c0D`wL6(X^8?*2YAJeZE+~CFpi;G2eEiD"
    "Odn=*7&a>x+PX~&MIuu~7%bh%cQomBX^95JLBM0i)MRP!ug~mrlXB7BpNLLXlee6tz#FW&H?aNNkYkly63%RC$RBee%Aat=1+U^cPDrIKt|8XH"
    "S3gA?LIbFBq63Ljc%e$T{6tVZn#a|2pXi-U}YxJByY%$CtU{u*{3~>*h>h31-=c!WIGcg<p3~_<>J_8m%OCM-bl@>7_oL&`GTCGH;-X2I3?Yl"
    "}b)m2yTpd5)>DM{#Xip|l$!o3)kCK`q9U<LEsg-M(yP~uH(k4z5oFj20cdCeIh^m;Gcvohp>V6iA65@;>N%zpw7P<7^|Fg@F7nw$CRx~H&uR!"
    "4k~i5XL-hcE<7jm>uHxDL<<8tR3xCLADjp691EO9&P=s~@I7i>0L!lW5T9Zhj22RarGG8+Ti-M`_e;Xci|P-7!lnOKc(gdR8LNP`45$hx8#Hay"
    "1_Nc3F)TUXg|=B}+BSL*;>0dPw{g&!t;!18_z1Te0emdM_YY";

static int g_known[4512];
static uint8_t g_mask[4096];
static uint32_t g_crc[256];
static int g_tables = 0;

static void init_tables(void) {
    if (g_tables) return;
    static const char* alpha = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz!#$%&()*+-;<=>?@^_`{|}~";
    int dec[256];
    for (int i = 0; i < 256; ++i) dec[i] = 0;
    for (int i = 0; i < 85; ++i) dec[(unsigned char)alpha[i]] = i;
    uint8_t bytes[600];
    int nb = 0;
    size_t n = strlen(KNOWN_B85);
    for (size_t i = 0; i + 5 <= n; i += 5) {
        uint64_t acc = 0;
        for (int j = 0; j < 5; ++j) acc = acc * 85 + (uint64_t)dec[(unsigned char)KNOWN_B85[i + j]];
        for (int j = 0; j < 4; ++j) bytes[nb++] = (uint8_t)(acc >> (24 - 8 * j));
    }
    for (int i = 0; i < 4512; ++i) g_known[i] = ((bytes[i >> 3] >> (i & 7)) & 1) ? 1 : -1;
    /* whitening mask: PN15 x^15+x^14+1, LSB first, 4094 bytes + wrap (ofdm_packet_utils.py:195-451) */
    static uint8_t g[4094 * 8];
    memset(g, 0, sizeof(g));
    for (int i = 0; i < 14; ++i) g[i] = 1;
    for (int k = 15; k < 4094 * 8; ++k) g[k] = g[k - 14] ^ g[k - 15];
    memset(g_mask, 0, sizeof(g_mask));
    for (int k = 0; k < 4094 * 8; ++k) g_mask[k >> 3] |= (uint8_t)(g[k] << (k & 7));
    g_mask[4094] = g_mask[0];
    g_mask[4095] = g_mask[1];
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? ((c << 1) ^ 0x04C11DB7u) : (c << 1);
        g_crc[i] = c;
    }
    g_tables = 1;
}

static uint32_t crc32_gr(const uint8_t* p, int n) {
    uint32_t crc = 0xFFFFFFFFu;
    for (int i = 0; i < n; ++i) crc = g_crc[(p[i] ^ (crc >> 24)) & 0xFF] ^ (crc << 8);
    return ~crc;
}

/* make_packet with pad_for_usrp = False (transmit_path.py:47): returns the packet length */
int oc_make_packet(const uint8_t* payload, int plen, uint8_t* out) {
    init_tables();
    int L = plen + 4;
    uint32_t v = (uint32_t)(L & 0x0FFF);
    out[0] = (uint8_t)(v >> 8); out[1] = (uint8_t)v; out[2] = (uint8_t)(v >> 8); out[3] = (uint8_t)v;
    uint32_t c = crc32_gr(payload, plen);
    int o = 0;
    for (int i = 0; i < plen; ++i, ++o) out[4 + o] = payload[i] ^ g_mask[o];
    for (int i = 0; i < 4; ++i, ++o) out[4 + o] = (uint8_t)(c >> (24 - 8 * i)) ^ g_mask[o];
    out[4 + o] = 0x55 ^ g_mask[o];
    return plen + 9;
}

/* ---------------------------------------------------------------- layout */
typedef struct {
    int N, occ, cp, L, zl, M, nbits, ncar, ntaps;
    int* tx_map;      /* [ncar] index into the N-wide vector */
    int* sink_map;    /* [ncar] index into the occ-wide vector */
    float* ks;        /* [occ] */
    float* kd;        /* [occ] */
    float* taps;      /* [ntaps] */
    cf* cst;          /* [M] */
    cd* tw;           /* [N] forward twiddles for the symbol FFT */
} layout;

static void carrier_hex(int occ, char* out) {
    char buf[2048];
    strcpy(buf, "FE7F");
    int diff = occ - 16;
    while (diff > 7) {
        char t[2048];
        t[0] = 'f'; strcpy(t + 1, buf); strcat(t, "f");
        strcpy(buf, t);
        diff -= 8;
    }
    if (diff > 0) {
        const char* hx = "0123456789abcdef";
        int dl = (int)ceil(diff / 2.0), dr = diff - dl;
        char t[2048];
        t[0] = hx[(1 << dl) - 1]; strcpy(t + 1, buf);
        size_t l = strlen(t);
        t[l] = hx[0xF ^ ((1 << dr) - 1)]; t[l + 1] = 0;
        strcpy(buf, t);
    }
    strcpy(out, buf);
}

static int hexval(char c) { return (c >= '0' && c <= '9') ? c - '0' : ((c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10); }

static layout* layout_new(const oc_cfg* c) {
    init_tables();
    layout* y = (layout*)calloc(1, sizeof(layout));
    y->N = c->N; y->occ = c->occ; y->cp = c->cp; y->L = c->N + c->cp; y->M = c->M;
    y->zl = (int)ceil((c->N - c->occ) / 2.0);
    while ((1 << y->nbits) < c->M) y->nbits++;
    char hx[2048];
    carrier_hex(c->occ, hx);
    int nh = (int)strlen(hx), pad = (c->N / 4 - nh) / 2;
    y->tx_map = (int*)malloc(sizeof(int) * c->occ);
    y->sink_map = (int*)malloc(sizeof(int) * c->occ);
    for (int i = 0; i < nh; ++i)
        for (int j = 0; j < 4; ++j)
            if ((hexval(hx[i]) >> (3 - j)) & 1) { y->tx_map[y->ncar] = 4 * (i + pad) + j; y->sink_map[y->ncar] = 4 * i + j; y->ncar++; }
    y->ks = (float*)calloc(c->occ, sizeof(float));
    y->kd = (float*)calloc(c->occ, sizeof(float));
    for (int i = 0; i < c->occ; ++i) y->ks[i] = ((y->zl + i) & 1) ? 0.f : (float)g_known[i];
    for (int i = 0; i + 2 < c->occ; i += 2) { float d = y->ks[i] - y->ks[i + 2]; y->kd[i] = d * d; }
    /* firdes.low_pass Hamming (ofdm_receiver.py~:69-76) */
    double bw = ((double)c->occ / (double)c->N) / 2.0, tb = bw * 0.08, fc = bw + tb;
    int nt = (int)(53.0 / (22.0 * tb));
    if ((nt & 1) == 0) ++nt;
    y->ntaps = nt;
    y->taps = (float*)malloc(sizeof(float) * nt);
    int M0 = (nt - 1) / 2;
    double* t = (double*)malloc(sizeof(double) * nt);
    double fw = 2 * M_PI * fc;
    for (int k = -M0; k <= M0; ++k) {
        double w = 0.54 - 0.46 * cos(2 * M_PI * (k + M0) / (nt - 1));
        t[k + M0] = (k == 0) ? fw / M_PI * w : sin(k * fw) / (k * M_PI) * w;
    }
    double fmax = t[M0];
    for (int k = 1; k <= M0; ++k) fmax += 2 * t[k + M0];
    for (int k = 0; k < nt; ++k) y->taps[k] = (float)(t[k] / fmax);
    free(t);
    y->cst = (cf*)malloc(sizeof(cf) * c->M);
    for (int i = 0; i < c->M; ++i) { y->cst[i].re = c->constellation[2 * i]; y->cst[i].im = c->constellation[2 * i + 1]; }
    y->tw = (cd*)malloc(sizeof(cd) * c->N);
    for (int i = 0; i < c->N; ++i) { y->tw[i].re = cos(2 * M_PI * i / c->N); y->tw[i].im = -sin(2 * M_PI * i / c->N); }
    return y;
}

static void layout_free(layout* y) {
    free(y->tx_map); free(y->sink_map); free(y->ks); free(y->kd); free(y->taps); free(y->cst); free(y->tw); free(y);
}

/* in-place radix-2 FFT in double; sign = -1 forward, +1 backward (unnormalised); tw = forward twiddles of size n */
static void fft_d(cd* a, int n, int sign, const cd* tw) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cd t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                cd w = tw[k * step];
                if (sign > 0) w.im = -w.im;
                cd u = a[i + k], v = a[i + k + len / 2];
                cd t = { v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re };
                a[i + k].re = u.re + t.re; a[i + k].im = u.im + t.im;
                a[i + k + len / 2].re = u.re - t.re; a[i + k + len / 2].im = u.im - t.im;
            }
    }
}

static uint32_t pad_index(uint64_t seed, uint64_t frame, uint32_t sym, uint32_t car, uint32_t M) {
    uint64_t x = seed ^ (frame << 32) ^ ((uint64_t)sym << 16) ^ (uint64_t)car;
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z & (uint64_t)(M - 1));
}

int oc_frame_symbols(const oc_cfg* c, int pkt_len) {
    layout* y = layout_new(c);
    int64_t per = (int64_t)y->ncar * y->nbits;
    int nd = pkt_len > 0 ? (int)((8ll * pkt_len + per - 1) / per) : 1;
    if (nd < 1) nd = 1;
    layout_free(y);
    return 1 + nd;
}

/* mapper -> preamble -> IFFT -> CP -> *1/sqrt(N) -> *amp  (A.4); returns the number of samples written */
static int64_t tx_frames(const layout* y, const oc_cfg* c, const uint8_t* pkts, const int64_t* pkt_off, int F,
                         int64_t first_frame, cf* out) {
    const int N = y->N, cp = y->cp;
    cd* buf = (cd*)malloc(sizeof(cd) * N);
    const float s1 = (float)(1.0 / sqrt((double)N));
    float amp = c->amp < 0.f ? 0.f : (c->amp > 1.f ? 1.f : c->amp);
    int64_t pos = 0;
    for (int f = 0; f < F; ++f) {
        const uint8_t* pk = pkts + pkt_off[f];
        int len = (int)(pkt_off[f + 1] - pkt_off[f]);
        int64_t bits = 8ll * len, per = (int64_t)y->ncar * y->nbits;
        int nd = len > 0 ? (int)((bits + per - 1) / per) : 1;
        if (nd < 1) nd = 1;
        for (int m = 0; m <= nd; ++m) {
            for (int i = 0; i < N; ++i) { buf[i].re = 0; buf[i].im = 0; }
            if (m == 0) {
                for (int i = 0; i < y->occ; ++i) buf[(y->zl + i + N / 2) % N].re = y->ks[i];
            } else {
                for (int cidx = 0; cidx < y->ncar; ++cidx) {
                    int64_t b0 = ((int64_t)(m - 1) * y->ncar + cidx) * y->nbits;
                    uint32_t val;
                    if (b0 + y->nbits <= bits) {
                        int64_t bb = b0 >> 3; int sh = (int)(b0 & 7);
                        uint32_t w = pk[bb];
                        if (sh + y->nbits > 8) w |= ((uint32_t)pk[bb + 1]) << 8;
                        val = (w >> sh) & ((1u << y->nbits) - 1u);
                    } else val = pad_index(c->pad_seed, (uint64_t)(first_frame + f), (uint32_t)(m - 1), (uint32_t)cidx, (uint32_t)y->M);
                    cd* d = &buf[(y->tx_map[cidx] + N / 2) % N];
                    d->re = y->cst[val].re; d->im = y->cst[val].im;
                }
            }
            fft_d(buf, N, +1, y->tw);
            cf* o = out + pos;
            for (int i = 0; i < N; ++i) {
                float re = ((float)buf[i].re * s1) * amp, im = ((float)buf[i].im * s1) * amp;
                o[cp + i].re = re; o[cp + i].im = im;
                if (i >= N - cp) { o[i - (N - cp)].re = re; o[i - (N - cp)].im = im; }
            }
            pos += N + cp;
        }
    }
    free(buf);
    return pos;
}

int64_t oc_tx(const oc_cfg* c, const uint8_t* pkts, const int64_t* pkt_off, int F, int64_t first_frame, float* out) {
    layout* y = layout_new(c);
    int64_t n = tx_frames(y, c, pkts, pkt_off, F, first_frame, (cf*)out);
    layout_free(y);
    return n;
}

/* ---------------------------------------------------------------- receiver */
/* sum_{k<w} v[n-k], zero history, float64 without subtraction: blocks of w, prefix + suffix of the previous block
 * (same association as _sliding_sum64 in ofdm_oracle.py) */
static void sliding_sum64(const float* v, int64_t n, int w, float* out, double* pre, double* suf) {
    int64_t nb = (n + w - 1) / w;
    for (int64_t b = 0; b < nb; ++b) {
        int64_t lo = b * w, hi = lo + w < n ? lo + w : n;
        double run = 0;
        for (int64_t i = lo; i < hi; ++i) { run += (double)v[i]; pre[i] = run; }
        run = 0;
        for (int64_t i = lo + w - 1; i >= lo; --i) { if (i < n) run += (double)v[i]; if (i < n) suf[i] = run; }
    }
    for (int64_t i = 0; i < n; ++i) {
        int64_t b = i / w, k = i - b * w;
        double s = pre[i];
        if (b >= 1 && k < w - 1) s += suf[(b - 1) * w + k + 1];
        out[i] = (float)s;
    }
}

typedef struct {
    int64_t n_trig, n_vec, n_pkts;
} rx_counts;

typedef struct {
    /* frame acquisition */
    float *Hr, *Hi; int delta, cnt;
    /* frame sink */
    int state; uint32_t bitbuf; int nbitbuf; uint32_t header; int hdr_cnt; float freq, phase; float *dr, *di;
    int pktlen, pktcnt; uint8_t pkt[4096];
} rx_state;

static void expj32(float ph, float* c, float* s) { *c = (float)cos((double)ph); *s = (float)sin((double)ph); }

static void coarse_comp(const layout* y, int delta, int cnt, float* c, float* s) {
    float a = (float)(-2.0 * M_PI * delta * y->cp);
    float ph = (a / (float)y->N) * (float)cnt;
    expj32(ph, c, s);
}

static inline void cmulf(float ar, float ai, float br, float bi, float* r, float* i) {
    float t1 = ar * br, t2 = ai * bi, t3 = ar * bi, t4 = ai * br;
    *r = t1 - t2; *i = t3 + t4;
}

static inline void cdivf(float ar, float ai, float br, float bi, float* r, float* i) {
    float t = br * br + bi * bi;
    float nr = ar * br + ai * bi, ni = ai * br - ar * bi;
    *r = nr / t; *i = ni / t;
}

int64_t oc_rx(const oc_cfg* c, const float* xin, int64_t n, int64_t* trig_out, float* ang_out, int64_t max_trig,
              uint8_t* pkt_bytes, int32_t stride, int32_t* pkt_len, uint8_t* pkt_ok, int64_t max_pkts, int64_t* counts) {
    layout* y = layout_new(c);
    const int N = y->N, cp = y->cp, L = y->L, occ = y->occ, zl = y->zl, h = N / 2;
    const cf* x = (const cf*)xin;
    cf* yf = (cf*)calloc((size_t)(n > 0 ? n : 1), sizeof(cf));
    /* ---- channel filter: overlap-save in float64 (fft_filter_ccc, A.5) ---- */
    {
        int nos = 1024;
        while (nos < 4 * y->ntaps) nos *= 2;
        cd* tw = (cd*)malloc(sizeof(cd) * nos);
        for (int i = 0; i < nos; ++i) { tw[i].re = cos(2 * M_PI * i / nos); tw[i].im = -sin(2 * M_PI * i / nos); }
        cd* Hf = (cd*)calloc(nos, sizeof(cd));
        for (int t = 0; t < y->ntaps; ++t) Hf[t].re = (double)y->taps[t];
        fft_d(Hf, nos, -1, tw);
        cd* b = (cd*)malloc(sizeof(cd) * nos);
        int hist = y->ntaps - 1, V = nos - hist;
        for (int64_t blk = 0; blk * V < n; ++blk) {
            int64_t in0 = blk * V - hist;
            for (int i = 0; i < nos; ++i) {
                int64_t gi = in0 + i;
                if (gi >= 0 && gi < n) { b[i].re = x[gi].re; b[i].im = x[gi].im; } else { b[i].re = 0; b[i].im = 0; }
            }
            fft_d(b, nos, -1, tw);
            for (int i = 0; i < nos; ++i) {
                cd t = { b[i].re * Hf[i].re - b[i].im * Hf[i].im, b[i].re * Hf[i].im + b[i].im * Hf[i].re };
                b[i] = t;
            }
            fft_d(b, nos, +1, tw);
            for (int i = hist; i < nos; ++i) {
                int64_t o = blk * V + (i - hist);
                if (o < n) { yf[o].re = (float)(b[i].re / nos); yf[o].im = (float)(b[i].im / nos); }
            }
        }
        free(tw); free(Hf); free(b);
    }
    /* ---- Schmidl-Cox metric (A.6) ---- */
    float* cre = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* cim = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* en = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* Pr = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* Pi = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* R = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    float* mf = (float*)malloc(sizeof(float) * (size_t)(n + 1));
    double* pre = (double*)malloc(sizeof(double) * (size_t)(n + 1));
    double* suf = (double*)malloc(sizeof(double) * (size_t)(n + 1));
    for (int64_t i = 0; i < n; ++i) {
        float yr = yf[i].re, yi = yf[i].im, dr = 0.f, di = 0.f;
        if (i >= h) { dr = yf[i - h].re; di = yf[i - h].im; }
        float a1 = yr * dr, a2 = yi * di, a3 = yi * dr, a4 = yr * di;
        cre[i] = a1 + a2; cim[i] = a3 - a4;
        float e1 = yr * yr, e2 = yi * yi;
        en[i] = e1 + e2;
    }
    sliding_sum64(cre, n, h, Pr, pre, suf);
    sliding_sum64(cim, n, h, Pi, pre, suf);
    sliding_sum64(en, n, h, R, pre, suf);
    {
        const double tap = (double)(float)(1.0 / cp);
        double run = 0;
        for (int64_t i = 0; i < n; ++i) {
            float p1 = Pr[i] * Pr[i], p2 = Pi[i] * Pi[i];
            float num = p1 + p2, den = R[i] * R[i];
            float Mt = num / den;
            run += (double)Mt;
            pre[i] = run;
            double s = (i >= cp) ? run - pre[i - cp] : run;
            float sf = (float)(s * tap);
            mf[i] = sf + (-1.0f);
        }
    }
    /* ---- peak detector (A.7), whole-stream state machine ---- */
    int64_t nt = 0;
    int64_t* trig = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n / 2 + 2));
    {
        const double a1 = (double)0.001f, a2 = 1.0 - a1;
        double avg = 0.0;
        int state = 0;
        float peak = -INFINITY;
        int64_t ind = 0, i = 0;
        while (i < n) {
            float v = mf[i];
            float thr = (float)avg * 0.2f;
            if (state == 0) {
                if (v > thr) state = 1;
                else { avg = a1 * (double)v + a2 * avg; ++i; }
            } else {
                if (v > peak) { peak = v; ind = i; avg = a1 * (double)v + a2 * avg; ++i; }
                else if (v > thr) { avg = a1 * (double)v + a2 * avg; ++i; }
                else { trig[nt++] = ind; state = 0; peak = -INFINITY; }
            }
        }
    }
    float* ang = (float*)malloc(sizeof(float) * (size_t)(nt + 1));
    for (int64_t k = 0; k < nt; ++k) ang[k] = (float)atan2((double)Pi[trig[k]], (double)Pr[trig[k]]);
    for (int64_t k = 0; k < nt && k < max_trig; ++k) { if (trig_out) trig_out[k] = trig[k]; if (ang_out) ang_out[k] = ang[k]; }
    /* ---- NCO phase before each trigger (A.8) ---- */
    double* phi0 = (double*)calloc((size_t)(nt + 1), sizeof(double));
    double* step = (double*)calloc((size_t)(nt + 1), sizeof(double));
    for (int64_t k = 0; k < nt; ++k) step[k] = (-2.0 / N) * (double)ang[k];
    for (int64_t k = 1; k < nt; ++k) phi0[k] = phi0[k - 1] + step[k - 1] * (double)(trig[k] - trig[k - 1]);
    /* ---- sampler + FFT + acquisition + sink, call by call (A.9-A.11) ---- */
    rx_state st;
    memset(&st, 0, sizeof(st));
    st.Hr = (float*)malloc(sizeof(float) * occ); st.Hi = (float*)calloc(occ, sizeof(float));
    st.dr = (float*)malloc(sizeof(float) * y->ncar); st.di = (float*)calloc(y->ncar, sizeof(float));
    for (int i = 0; i < occ; ++i) st.Hr[i] = 1.f;
    st.cnt = 1;
    cd* fb = (cd*)malloc(sizeof(cd) * N);
    float* Sr = (float*)malloc(sizeof(float) * N); float* Si = (float*)malloc(sizeof(float) * N);
    float* er_ = (float*)malloc(sizeof(float) * occ); float* ei_ = (float*)malloc(sizeof(float) * occ);
    int64_t npk = 0, nvec = 0;
    {
        int64_t pos = 0, lo = 0;
        int sstate = 0, timeout = 0;
        while (pos + L + N < n) {
            while (lo < nt && trig[lo] < pos + N) ++lo;
            int flag = 0;
            int64_t vs;
            if (lo < nt && trig[lo] <= pos + L + N) {
                vs = trig[lo] - N + 1; flag = 1; timeout = 1000; sstate = 2; pos = vs;
            } else if (sstate == 2) {
                vs = pos + L; if (timeout-- == 0) sstate = 0; pos += L;   /* post-decrement: up to 1001 data vectors */
            } else { pos += L + 1; continue; }
            ++nvec;
            /* derotate + FFT + shift */
            int64_t kk = -1;
            { int64_t a = 0, b = nt; while (a < b) { int64_t m = (a + b) >> 1; if (trig[m] <= vs) a = m + 1; else b = m; } kk = a - 1; }
            for (int i = 0; i < N; ++i) {
                int64_t s = vs + i;
                while (kk + 1 < nt && trig[kk + 1] <= s) ++kk;
                double ph = kk >= 0 ? phi0[kk] + step[kk] * (double)(s - trig[kk] + 1) : 0.0;
                float cs = (float)cos(ph), sn = (float)sin(ph);
                float zr, zi;
                cmulf(yf[s].re, yf[s].im, cs, sn, &zr, &zi);
                fb[i].re = zr; fb[i].im = zi;
            }
            fft_d(fb, N, -1, y->tw);
            for (int i = 0; i < N; ++i) { int k = (i + N / 2) % N; Sr[k] = (float)fb[i].re; Si[k] = (float)fb[i].im; }
            /* acquisition */
            if (flag) {
                st.cnt = 1;
                float best = 0.f; int index = 0;
                for (int sh = zl - 4; sh < zl + 4; ++sh) {
                    double acc = 0;
                    for (int j = 0; j < occ; ++j) {
                        int p = sh + j;
                        float sd = 0.f;
                        if (p >= 0 && p < N - 2) { float d1 = Sr[p] - Sr[p + 2], d2 = Si[p] - Si[p + 2]; float q1 = d1 * d1, q2 = d2 * d2; sd = q1 + q2; }
                        acc += (double)y->kd[j] * (double)sd;
                    }
                    float sf = (float)acc;
                    if (sf > best) { best = sf; index = sh; }
                }
                st.delta = index - zl;
                float c1, s1;
                coarse_comp(y, st.delta, 1, &c1, &s1);
                for (int i = 0; i < occ; i += 2) {
                    float br, bi;
                    cmulf(c1, s1, Sr[i + zl + st.delta], Si[i + zl + st.delta], &br, &bi);
                    cdivf(y->ks[i], 0.f, br, bi, &st.Hr[i], &st.Hi[i]);
                }
                for (int i = 1; i + 1 < occ; i += 2) {
                    st.Hr[i] = (st.Hr[i + 1] + st.Hr[i - 1]) * 0.5f;
                    st.Hi[i] = (st.Hi[i + 1] + st.Hi[i - 1]) * 0.5f;
                }
                if ((occ & 1) == 0) { st.Hr[occ - 1] = st.Hr[occ - 2]; st.Hi[occ - 1] = st.Hi[occ - 2]; }
            }
            {
                float cc, cs;
                coarse_comp(y, st.delta, st.cnt, &cc, &cs);
                for (int i = 0; i < occ; ++i) {
                    float tr, ti;
                    cmulf(st.Hr[i], st.Hi[i], cc, cs, &tr, &ti);
                    cmulf(tr, ti, Sr[i + zl + st.delta], Si[i + zl + st.delta], &er_[i], &ei_[i]);
                }
                st.cnt++;
                if (st.cnt == 1000) st.cnt = 1;
            }
            /* frame sink */
            if (st.state == 0) {
                if (flag) {
                    st.state = 1; st.bitbuf = 0; st.nbitbuf = 0; st.header = 0; st.hdr_cnt = 0; st.freq = 0.f; st.phase = 0.f;
                    for (int i = 0; i < y->ncar; ++i) { st.dr[i] = 1.f; st.di[i] = 0.f; }
                }
                continue;
            }
            uint8_t data[4200];
            int nd = 0;
            {
                float car_r, car_i;
                expj32(st.phase, &car_r, &car_i);
                double sre = 0, sim = 0;
                for (int cidx = 0; cidx < y->ncar; ++cidx) {
                    int i = y->sink_map[cidx];
                    float tr, ti, rr, ri;
                    cmulf(er_[i], ei_[i], car_r, car_i, &tr, &ti);
                    cmulf(tr, ti, st.dr[cidx], st.di[cidx], &rr, &ri);
                    int b = 0; float bestd = 0.f;
                    for (int k = 0; k < y->M; ++k) {
                        float dx = rr - y->cst[k].re, dy = ri - y->cst[k].im;
                        float q1 = dx * dx, q2 = dy * dy, dd = q1 + q2;
                        if (k == 0 || dd < bestd) { bestd = dd; b = k; }
                    }
                    float clr = y->cst[b].re, cli = y->cst[b].im;
                    float e1 = rr * clr, e2 = ri * cli, e3 = ri * clr, e4 = rr * cli;
                    sre += (double)(e1 + e2); sim += (double)(e3 - e4);
                    float n1 = rr * rr, n2 = ri * ri;
                    if (n1 + n2 > 0.001f) {
                        float qr, qi;
                        cdivf(clr, cli, rr, ri, &qr, &qi);
                        float u1 = qr - st.dr[cidx], u2 = qi - st.di[cidx];
                        float g1 = 0.05f * u1, g2 = 0.05f * u2;
                        st.dr[cidx] = st.dr[cidx] + g1; st.di[cidx] = st.di[cidx] + g2;
                    }
                    st.bitbuf |= ((uint32_t)b) << st.nbitbuf;
                    st.nbitbuf += y->nbits;
                    while (st.nbitbuf >= 8) { data[nd++] = (uint8_t)(st.bitbuf & 0xFF); st.bitbuf >>= 8; st.nbitbuf -= 8; }
                }
                float angle = (float)atan2((double)(float)sim, (double)(float)sre);
                float fgain = 0.015625f * angle;
                st.freq = st.freq - fgain;
                float p1 = st.phase + st.freq, p2 = 0.25f * angle;
                float ph = p1 - p2;
                if ((double)ph >= 2 * M_PI) ph = (float)((double)ph - 2 * M_PI);
                if ((double)ph < 0) ph = (float)((double)ph + 2 * M_PI);
                st.phase = ph;
            }
            int j = 0;
            if (st.state == 1) {
                while (j < nd) {
                    st.header = (st.header << 8) | data[j];
                    ++j;
                    if (++st.hdr_cnt == 4) {
                        if (((st.header >> 16) ^ (st.header & 0xFFFF)) == 0) {
                            st.state = 2; st.pktlen = (int)((st.header >> 16) & 0x0FFF); st.pktcnt = 0;
                            while (j < nd && st.pktcnt < st.pktlen) st.pkt[st.pktcnt++] = data[j++];
                            if (st.pktcnt == st.pktlen) goto deliver;
                        } else st.state = 0;
                        break;
                    }
                }
                continue;
            }
            while (j < nd) {
                st.pkt[st.pktcnt++] = data[j++];
                if (st.pktcnt == st.pktlen) goto deliver;
            }
            continue;
        deliver:
            st.state = 0;
            if (npk < max_pkts) {
                /* unmake_packet: dewhiten + CRC */
                uint8_t body[4096];
                int len = st.pktlen;
                for (int i = 0; i < len; ++i) body[i] = st.pkt[i] ^ g_mask[i];
                int ok = 0;
                if (len >= 4) {
                    uint32_t want = ((uint32_t)body[len - 4] << 24) | ((uint32_t)body[len - 3] << 16) | ((uint32_t)body[len - 2] << 8) | body[len - 1];
                    ok = crc32_gr(body, len - 4) == want;
                }
                if (pkt_len) pkt_len[npk] = len;
                if (pkt_ok) pkt_ok[npk] = (uint8_t)ok;
                if (pkt_bytes) memcpy(pkt_bytes + (size_t)npk * stride, body, (size_t)(len < stride ? len : stride));
            }
            ++npk;
        }
    }
    if (counts) { counts[0] = nt; counts[1] = nvec; counts[2] = npk; }
    free(st.Hr); free(st.Hi); free(st.dr); free(st.di); free(fb); free(Sr); free(Si); free(er_); free(ei_);
    free(phi0); free(step); free(ang); free(trig);
    free(cre); free(cim); free(en); free(Pr); free(Pi); free(R); free(mf); free(pre); free(suf); free(yf);
    layout_free(y);
    return npk;
}

/* ---------------------------------------------------------------- timed multi-threaded loopback */
typedef struct {
    const oc_cfg* cfg; int frames, psize; double snr_db, cfo; uint64_t seed;
    double t_mod, t_demod; int64_t samples, npk, nok;
} job;

static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void* job_run(void* arg) {
    job* j = (job*)arg;
    const oc_cfg* c = j->cfg;
    layout* y = layout_new(c);
    int F = j->frames, ps = j->psize, plen = ps + 9;
    uint8_t* pay = (uint8_t*)malloc((size_t)F * ps);
    for (int f = 0; f < F; ++f) {
        for (int i = 0; i < ps; ++i) pay[(size_t)f * ps + i] = (uint8_t)(mix64(j->seed ^ ((uint64_t)f << 20) ^ i) & 0xFF);
        pay[(size_t)f * ps] = (uint8_t)(f >> 8); pay[(size_t)f * ps + 1] = (uint8_t)f; pay[(size_t)f * ps + 2] = 0; pay[(size_t)f * ps + 3] = 0;
    }
    int64_t per = (int64_t)y->ncar * y->nbits;
    int nsym = 1 + (int)((8ll * plen + per - 1) / per);
    int64_t nsig = (int64_t)F * nsym * y->L, lead = 2 * y->L, n = nsig + 2 * lead;
    uint8_t* pk = (uint8_t*)malloc((size_t)F * plen);
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (F + 1));
    cf* x = (cf*)calloc((size_t)n, sizeof(cf));
    double t0 = now_s();
    for (int f = 0; f <= F; ++f) off[f] = (int64_t)f * plen;
    for (int f = 0; f < F; ++f) oc_make_packet(pay + (size_t)f * ps, ps, pk + (size_t)f * plen);
    tx_frames(y, c, pk, off, F, 0, x + lead);
    j->t_mod = now_s() - t0;
    /* channel (untimed) */
    double p = 0;
    for (int64_t i = lead; i < lead + nsig; ++i) p += (double)x[i].re * x[i].re + (double)x[i].im * x[i].im;
    p /= (double)nsig;
    double sigma = sqrt(p / pow(10.0, j->snr_db / 10.0) / 2.0), w = 2 * M_PI * j->cfo / c->N;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t hsh = mix64(j->seed * 77 + (uint64_t)i);
        double u1 = ((double)(hsh >> 40) + 0.5) / 16777216.0, u2 = ((double)((hsh >> 8) & 0xFFFFFF) + 0.5) / 16777216.0;
        double rad = sigma * sqrt(-2.0 * log(u1));
        double cr = cos(w * i), sr = sin(w * i);
        double re = x[i].re * cr - x[i].im * sr + rad * cos(2 * M_PI * u2), im = x[i].re * sr + x[i].im * cr + rad * sin(2 * M_PI * u2);
        x[i].re = (float)re; x[i].im = (float)im;
    }
    uint8_t* ok = (uint8_t*)calloc((size_t)F + 16, 1);
    int64_t counts[3];
    t0 = now_s();
    int64_t npk = oc_rx(c, (const float*)x, n, NULL, NULL, 0, NULL, 0, NULL, ok, F + 16, counts);
    j->t_demod = now_s() - t0;
    j->samples = nsig; j->npk = npk; j->nok = 0;
    for (int64_t i = 0; i < npk && i < F + 16; ++i) j->nok += ok[i];
    free(pay); free(pk); free(off); free(x); free(ok);
    layout_free(y);
    return NULL;
}

/* out: [0] samples total, [1] seconds (max over threads of t_mod + t_demod), [2] packets, [3] crc ok,
 *      [4] max t_mod, [5] max t_demod */
int oc_loopback_mt2(const oc_cfg* c, int frames_per_thread, int psize, int nthreads, double snr_db, const double* cfos, double* out);
int oc_loopback_mt(const oc_cfg* c, int frames_per_thread, int psize, int nthreads, double snr_db, double cfo, double* out) {
    double* cf_ = (double*)malloc(sizeof(double) * (nthreads > 0 ? nthreads : 1));
    for (int t = 0; t < nthreads; ++t) cf_[t] = cfo;
    int rc = oc_loopback_mt2(c, frames_per_thread, psize, nthreads, snr_db, cf_, out);
    free(cf_);
    return rc;
}

/* the same with one carrier-frequency offset per thread: every thread works on a piece of one 10 000-frame CFO segment of
 * the bench capture (bench.py redraws the offset per segment), cfos[t] = that segment's offset */
int oc_loopback_mt2(const oc_cfg* c, int frames_per_thread, int psize, int nthreads, double snr_db, const double* cfos, double* out) {
    init_tables();
    job* jobs = (job*)calloc(nthreads, sizeof(job));
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].cfg = c; jobs[t].frames = frames_per_thread; jobs[t].psize = psize; jobs[t].snr_db = snr_db; jobs[t].cfo = cfos[t];
        jobs[t].seed = 1234 + 7919 * (uint64_t)t;
        pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    double tot = 0, tm = 0, td = 0;
    int64_t s = 0, npk = 0, nok = 0;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        double tt = jobs[t].t_mod + jobs[t].t_demod;
        if (tt > tot) tot = tt;
        if (jobs[t].t_mod > tm) tm = jobs[t].t_mod;
        if (jobs[t].t_demod > td) td = jobs[t].t_demod;
        s += jobs[t].samples; npk += jobs[t].npk; nok += jobs[t].nok;
    }
    out[0] = (double)s; out[1] = tot; out[2] = (double)npk; out[3] = (double)nok; out[4] = tm; out[5] = td;
    free(jobs); free(th);
    return 0;
}
