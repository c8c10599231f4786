// CPU check of the two arithmetic arguments the metric kernels (ofdm_uhd_b200/csrc/rx_sync_stream.cu, common.cuh)
// rest on -- a restatement of the device code, lane by lane, since shuffles and MUFU.RCP do not exist here; the device
// code itself is compared bit for bit with __fdiv_rn by ofdm_selftest_packed_math and with the oracle by the -m gpu tests.
//
//  1. fdiv_inrange (common.cuh): for operands whose biased exponents lie in [80, 175) the sequence
//       r0 ~ 1/d;  t = fma(-d, r0, 1);  r = fma(r0, t, r0);  q0 = n*r;  e = fma(-d, q0, n);  q = fma(r, e, q0)
//     returns the correctly rounded n / d for ANY starting reciprocal within 2 ulp of 1/d (MUFU.RCP is within 1).
//  2. bfly_scan (rx_sync_stream.cu): five xor-exchanges with integer-masked routing give every lane the sum of the
//     lanes in front of it and the sum of the lanes behind it; additions only (zeros stay exactly zero, integers stay
//     exact) and a NaN in lane j reaches `fwd` of the lanes behind j and `bwd` of the lanes in front of j, no other.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

static float as_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t as_u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint64_t as_u64(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static double as_d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

static float div_seq(float n, float d, float r0) {
    const float t = fmaf(-d, r0, 1.0f);
    const float r = fmaf(r0, t, r0);
    const float q0 = n * r;
    const float e = fmaf(-d, q0, n);
    return fmaf(r, e, q0);
}

static void bfly(const double (&run)[32], double (&fwd)[32], double (&bwd)[32], double (&tot)[32]) {
    double s[32], f[32] = {0}, b[32] = {0};
    for (int l = 0; l < 32; ++l) s[l] = run[l];
    for (int k = 0; k < 5; ++k) {
        double o[32];
        for (int l = 0; l < 32; ++l) o[l] = s[l ^ (1 << k)];
        for (int l = 0; l < 32; ++l) {
            const uint32_t hi = (uint32_t)(as_u64(o[l]) >> 32), lo = (uint32_t)as_u64(o[l]);
            const uint32_t up = 0u - (uint32_t)((l >> k) & 1);
            f[l] += as_d(((uint64_t)(hi & up) << 32) | (lo & up));
            b[l] += as_d(((uint64_t)(hi & ~up) << 32) | (lo & ~up));
            s[l] += o[l];
        }
    }
    for (int l = 0; l < 32; ++l) { fwd[l] = f[l]; bwd[l] = b[l]; tot[l] = s[l]; }
}

int main() {
    std::mt19937_64 g(20260101);
    // ---- 1. the division sequence -------------------------------------------------------------------------------
    long bad = 0, cases = 0;
    for (int trial = 0; trial < 4000000; ++trial) {
        const uint64_t h = g();
        uint32_t en = 80u + (uint32_t)(h & 0xFF) % 95u, ed = 80u + (uint32_t)((h >> 8) & 0xFF) % 95u;
        if (trial % 7 == 0) { en = (trial & 1) ? 80u : 174u; ed = (trial & 2) ? 80u : 174u; }       // the corners of the window
        uint32_t mn = (uint32_t)(h >> 16) & 0x7FFFFFu, md = (uint32_t)(h >> 40) & 0x7FFFFFu;
        if (trial % 11 == 0) md = 0x7FFFFFu;                                                         // all-ones divisor mantissa
        if (trial % 13 == 0) md = 0;                                                                 // power-of-two divisor
        if (trial % 17 == 0) mn = md;                                                                // quotient a power of two
        const float n = as_f((en << 23) | mn), d = as_f((ed << 23) | md);
        const float want = n / d;
        const float rc = (float)(1.0 / (double)d);
        for (int du = -2; du <= 2; ++du) {
            const float r0 = as_f(as_u(rc) + (uint32_t)du);
            ++cases;
            if (as_u(div_seq(n, d, r0)) != as_u(want)) {
                if (++bad < 5) printf("division: n=%a d=%a r0 off by %d ulp: %a, want %a\n", n, d, du, div_seq(n, d, r0), want);
            }
        }
    }
    if (bad) { printf("division sequence: %ld of %ld cases differ from n / d\n", bad, cases); return 1; }
    printf("division sequence ok (%ld cases)\n", cases);

    // ---- 2. the scan ---------------------------------------------------------------------------------------------
    for (int trial = 0; trial < 2000; ++trial) {
        double run[32], fwd[32], bwd[32], tot[32];
        const int mode = trial % 4;
        for (int l = 0; l < 32; ++l) {
            const int64_t v = (int64_t)(g() % 2000001) - 1000000;
            run[l] = (mode == 1 && (l < 9 || l > 20)) ? 0.0 : (double)v;              // integers: every order of addition is exact
        }
        int nan_at = -1;
        if (mode == 2) { nan_at = (int)(g() % 32); run[nan_at] = NAN; }
        if (mode == 3) { nan_at = (int)(g() % 32); run[nan_at] = INFINITY; }
        bfly(run, fwd, bwd, tot);
        for (int l = 0; l < 32; ++l) {
            double f = 0.0, b = 0.0, t = 0.0;
            bool f_bad = false, b_bad = false;
            for (int j = 0; j < 32; ++j) {
                if (j == nan_at) { if (j < l) f_bad = true; if (j > l) b_bad = true; continue; }
                if (j < l) f += run[j];
                if (j > l) b += run[j];
                t += run[j];
            }
            const bool okf = f_bad ? !std::isfinite(fwd[l]) : (as_u64(fwd[l]) == as_u64(f));
            const bool okb = b_bad ? !std::isfinite(bwd[l]) : (as_u64(bwd[l]) == as_u64(b));
            const bool okt = nan_at >= 0 ? !std::isfinite(tot[l]) : (tot[l] == t && as_u64(tot[l]) == as_u64(tot[0]));
            if (!okf || !okb || !okt) {
                printf("scan: trial %d lane %d: fwd %g (want %g%s) bwd %g (want %g%s) tot %g\n", trial, l, fwd[l], f,
                       f_bad ? ", poisoned" : "", bwd[l], b, b_bad ? ", poisoned" : "", tot[l]);
                return 1;
            }
        }
    }
    printf("bfly_scan ok\n");
    return 0;
}
