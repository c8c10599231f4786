// Device self-test of the packed-fp32 helpers in common.cuh: every "exact" helper (the ones the decision paths use)
// must return, bit for bit, what its definition in individually rounded scalar float32 operations returns.  ptxas is
// known to contract packed multiply -> packed add pairs into FFMA2 (common.cuh), so this is checked on the device the
// library runs on, over random operands that include zeros, denormals and huge values.
#include "internal.h"
#include "common.cuh"

__device__ __forceinline__ uint64_t st_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// a float with a random sign, a random mantissa and an exponent drawn from a window around 1 (most of the time),
// around the denormal range, or around 1e18; sometimes an exact zero
__device__ __forceinline__ float st_float(uint64_t h) {
    const uint32_t mant = (uint32_t)h & 0x7FFFFFu, sign = (uint32_t)(h >> 23) & 1u;
    const uint32_t sel = (uint32_t)(h >> 24) & 15u, e = (uint32_t)(h >> 28) & 31u;
    uint32_t ex;
    if (sel == 0) return sign ? -0.0f : 0.0f;
    if (sel == 1) ex = e >> 2;                       // denormals and the smallest normals
    else if (sel == 2) ex = 170u + e;                // ~1e13 .. 1e22
    else ex = 111u + e;                              // 2^-16 .. 2^15
    return __uint_as_float((sign << 31) | (ex << 23) | mant);
}

__device__ __forceinline__ bool st_same(float a, float b) {
    return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b);
}
__device__ __forceinline__ bool st_same2(float2 a, float2 b) { return st_same(a.x, b.x) && st_same(a.y, b.y); }

__global__ void __launch_bounds__(256) selftest_kernel(int64_t n, uint64_t seed, unsigned long long* bad) {
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i - (threadIdx.x & 31) < n; i += (int64_t)gridDim.x * blockDim.x) {   // whole warps (fdiv_block votes)
        const uint64_t h0 = st_mix(seed ^ (uint64_t)i), h1 = st_mix(h0), h2 = st_mix(h1), h3 = st_mix(h2);
        const float2 a = make_float2(st_float(h0), st_float(h1)), b = make_float2(st_float(h2), st_float(h3));
        const float s = st_float(st_mix(h3));
        // definitions in scalar, individually rounded operations
        const float2 r_mul = make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                                         __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
        const float2 r_mulc = make_float2(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                                          __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y)));
        const float r_norm = __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y));
        const float2 r_sub = make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y));
        const float2 r_add = make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
        const float2 r_scale = make_float2(__fmul_rn(a.x, s), __fmul_rn(a.y, s));
        const float r_t = __fadd_rn(__fmul_rn(b.x, b.x), __fmul_rn(b.y, b.y));
        const float2 r_div = make_float2(__fdiv_rn(r_mulc.x, r_t), __fdiv_rn(r_mulc.y, r_t));
        // the slicer distance and the DFE update as rx_demod.cu composes them
        const float2 e = csub_x(a, b);
        const float r_dist = __fadd_rn(__fmul_rn(r_sub.x, r_sub.x), __fmul_rn(r_sub.y, r_sub.y));
        const float2 dq = cscale_x(csub_x(a, b), 0.05f);
        const float2 r_dfe = make_float2(__fadd_rn(b.x, __fmul_rn(0.05f, __fsub_rn(a.x, b.x))),
                                         __fadd_rn(b.y, __fmul_rn(0.05f, __fsub_rn(a.y, b.y))));
        const float2 g_dfe = make_float2(fadd_rn(b.x, dq.x), fadd_rn(b.y, dq.y));
        // (a + b) * 0.5 (channel-estimate interpolation): add, then multiply -- nothing to contract
        const float2 r_mid = make_float2(__fmul_rn(__fadd_rn(a.x, b.x), 0.5f), __fmul_rn(__fadd_rn(a.y, b.y), 0.5f));
        unsigned m = 0;
        m |= st_same2(cmul_x(a, b), r_mul) ? 0u : 1u;
        m |= st_same2(cmulc_x(a, b), r_mulc) ? 0u : 2u;
        m |= st_same(norm_x(a), r_norm) ? 0u : 4u;
        m |= st_same2(csub_x(a, b), r_sub) ? 0u : 8u;
        m |= st_same2(cadd_x(a, b), r_add) ? 0u : 16u;
        m |= st_same2(cscale_x(a, s), r_scale) ? 0u : 32u;
        m |= st_same2(cdiv_x(a, b), r_div) ? 0u : 64u;
        m |= st_same(norm_x(e), r_dist) ? 0u : 128u;
        m |= st_same2(g_dfe, r_dfe) ? 0u : 256u;
        m |= st_same2(cscale_x(cadd_x(a, b), 0.5f), r_mid) ? 0u : 512u;
        // fdiv_block: in-range operands take the branch-free chains, anything else the compiler's division -- whole warps
        // at a time, so every eighth warp draws all its operands from the in-range window
        {
            const bool tame = ((threadIdx.x >> 5) & 7) == 0;
            float nn[4] = {fabsf(a.x), fabsf(a.y), fabsf(b.x), fabsf(s)}, dd[4] = {fabsf(b.y), fabsf(s), fabsf(a.y), fabsf(a.x)}, qq[4];
            if (tame) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    nn[j] = __uint_as_float(((100u + ((uint32_t)(h1 >> (8 * j)) & 63u)) << 23) | (__float_as_uint(nn[j]) & 0x7FFFFFu));
                    dd[j] = __uint_as_float(((100u + ((uint32_t)(h2 >> (8 * j)) & 63u)) << 23) | (__float_as_uint(dd[j]) & 0x7FFFFFu));
                }
            }
            fdiv_block<4>(nn, dd, qq);
#pragma unroll
            for (int j = 0; j < 4; ++j) m |= st_same(qq[j], __fdiv_rn(nn[j], dd[j])) ? 0u : 1024u;
        }
        if (m) { ++local; atomicOr(bad + 1, (unsigned long long)m); }
    }
    if (local) atomicAdd(bad, local);
}

extern "C" int ofdm_selftest_packed_math(int32_t device, int64_t n, uint64_t seed, int64_t* host_out2) {
    if (!host_out2 || n < 1) { ofdm_set_error("ofdm_selftest_packed_math: bad argument"); return OFDM_E_INVAL; }
    OFDM_CUDA_CHECK(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    OFDM_CUDA_CHECK(cudaMalloc((void**)&d, 2 * sizeof(unsigned long long)));
    cudaMemset(d, 0, 2 * sizeof(unsigned long long));
    int sms = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    selftest_kernel<<<sms * 8, 256>>>(n, seed, d);
    cudaError_t e = cudaGetLastError();
    unsigned long long h[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { ofdm_set_error("ofdm_selftest_packed_math: %s", cudaGetErrorString(e)); return OFDM_E_CUDA; }
    host_out2[0] = (int64_t)h[0];
    host_out2[1] = (int64_t)h[1];
    return OFDM_OK;
}
