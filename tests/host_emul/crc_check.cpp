// CPU check of crc32_step4 (ofdm_uhd_b200/csrc/common.cuh): CRC-32 four bytes per step over the slicing tables
// T0..T3, built the way ofdm_create builds them (api.cu), against the byte-at-a-time table CRC
// (digital.crc32, digital_swig.py:3151-3168: MSB-first 0x04C11DB7, init / final all ones; check value 0xFC891918... of
// "123456789" under this bit order is asserted by tests/test_tables.py against the oracle).
#define OFDM_HOST_EMUL
#include "../../ofdm_uhd_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main() {
    std::vector<uint32_t> t(1024);
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        t[i] = c;
    }
    for (int k = 1; k < 4; ++k)
        for (int i = 0; i < 256; ++i) {
            const uint32_t v = t[(k - 1) * 256 + i];
            t[k * 256 + i] = t[v >> 24] ^ (v << 8);
        }
    srand(7);
    for (int trial = 0; trial < 2000; ++trial) {
        const int n = 4 * (rand() % 300);
        std::vector<uint8_t> b(n);
        for (auto& x : b) x = (uint8_t)rand();
        uint32_t a = 0xFFFFFFFFu, c = 0xFFFFFFFFu;
        for (int i = 0; i < n; ++i) a = t[(b[i] ^ (a >> 24)) & 0xFF] ^ (a << 8);
        for (int i = 0; i < n; i += 4)
            c = crc32_step4(c, ((uint32_t)b[i] << 24) | ((uint32_t)b[i + 1] << 16) | ((uint32_t)b[i + 2] << 8) | b[i + 3], t.data());
        if (a != c) { printf("mismatch at trial %d (n=%d): %08x vs %08x\n", trial, n, a, c); return 1; }
    }
    printf("crc32_step4 ok\n");
    return 0;
}
