// Pipe-rate microbenchmark for the instruction classes the sync kernel leans on (sm_100a).
// One CTA of 128 threads per SM x 8 CTAs/SM, long dependent-free unrolled chains; prints warp-instr/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 32768
template <int OP> __global__ void k(float* out, float a, double b) {
    float f0 = a + threadIdx.x, f1 = a * 2 + threadIdx.x, f2 = a * 3, f3 = a * 5;
    double d0 = b + threadIdx.x, d1 = b * 2, d2 = b * 3, d3 = b * 5;
    unsigned u0 = threadIdx.x, u1 = 7;
    unsigned long long q0 = 0x3f8000003f800000ull + threadIdx.x, q1 = 0x3f8000013f800001ull, q2 = 0x3f8000023f800002ull,
                       q3 = 0x3f8000033f800003ull, qa = 0x3f8000103f800010ull;
#pragma unroll 8
    for (int i = 0; i < ITER; ++i) {
        if (OP == 0) { f0 = __fmaf_rn(f0, a, f1); f1 = __fmaf_rn(f1, a, f2); f2 = __fmaf_rn(f2, a, f3); f3 = __fmaf_rn(f3, a, f0); }
        if (OP == 1) { d0 = __dadd_rn(d0, d1); d1 = __dadd_rn(d1, d2); d2 = __dadd_rn(d2, d3); d3 = __dadd_rn(d3, d0); }
        if (OP == 2) { d0 = (double)f0; d1 = (double)f1; d2 = (double)f2; d3 = (double)f3;
                       f0 = __fadd_rn(f0, __double2float_rn(d3) ); asm volatile("" : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3));
                       f1 += 1.f; f2 += 1.f; f3 += 1.f; }           // 4 F2F.F64.F32 + 1 F2F.F32.F64 + 4 FADD
        if (OP == 3) { f0 = __double2float_rn(d0); f1 = __double2float_rn(d1); f2 = __double2float_rn(d2); f3 = __double2float_rn(d3);
                       asm volatile("" : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3)); d0 += 1.0; d1 += 1.0; d2 += 1.0; d3 += 1.0; }
        if (OP == 4) { f0 = __fdiv_rn(f1, f0 + 2.f); f1 = __fdiv_rn(f2, f1 + 2.f); f2 = __fdiv_rn(f3, f2 + 2.f); f3 = __fdiv_rn(f0, f3 + 2.f); }
        if (OP == 7) { d0 = (double)f0; d1 = (double)f1; d2 = (double)f2; d3 = (double)f3;
                       asm volatile("" : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3));
                       f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; }
        if (OP == 5) { u0 = __shfl_up_sync(0xffffffffu, u0, 1); u1 = __shfl_up_sync(0xffffffffu, u1, 2);
                       u0 += u1; u1 ^= u0; }
        if (OP == 8) {   // packed fp32: 4 independent FFMA2 chains (two fp32 FMAs per lane each)
            asm volatile("fma.rn.f32x2 %0, %0, %4, %1;\n\tfma.rn.f32x2 %1, %1, %4, %2;\n\tfma.rn.f32x2 %2, %2, %4, %3;\n\tfma.rn.f32x2 %3, %3, %4, %0;"
                         : "+l"(q0), "+l"(q1), "+l"(q2), "+l"(q3) : "l"(qa)); }
        if (OP == 9) {   // FADD2
            asm volatile("add.rn.f32x2 %0, %0, %1;\n\tadd.rn.f32x2 %1, %1, %2;\n\tadd.rn.f32x2 %2, %2, %3;\n\tadd.rn.f32x2 %3, %3, %0;"
                         : "+l"(q0), "+l"(q1), "+l"(q2), "+l"(q3)); }
        if (OP == 10) {  // 2 FFMA2 + 2 scalar FFMA interleaved
            asm volatile("fma.rn.f32x2 %0, %0, %2, %1;\n\tfma.rn.f32x2 %1, %1, %2, %0;" : "+l"(q0), "+l"(q1) : "l"(qa));
            f0 = __fmaf_rn(f0, a, f1); f1 = __fmaf_rn(f1, a, f0); }
        if (OP == 11) {  // FADD x4 (scalar)
            f0 = __fadd_rn(f0, f1); f1 = __fadd_rn(f1, f2); f2 = __fadd_rn(f2, f3); f3 = __fadd_rn(f3, f0); }
        if (OP == 12) {  // 2 FFMA2 + 2 integer (alu pipe) ops
            asm volatile("fma.rn.f32x2 %0, %0, %2, %1;\n\tfma.rn.f32x2 %1, %1, %2, %0;" : "+l"(q0), "+l"(q1) : "l"(qa));
            u0 = (u0 ^ u1) + 3u; u1 = (u1 & u0) + 5u; }
        if (OP == 13) {  // 2 scalar FFMA + 2 integer ops
            f0 = __fmaf_rn(f0, a, f1); f1 = __fmaf_rn(f1, a, f0);
            u0 = (u0 ^ u1) + 3u; u1 = (u1 & u0) + 5u; }
        if (OP == 6) { d0 = fma(d0, d1, d2); d1 = fma(d1, d2, d3); d2 = fma(d2, d3, d0); d3 = fma(d3, d0, d1); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + (float)(d0 + d1 + d2 + d3) + u0 + u1 + (float)(q0 ^ q1 ^ q2 ^ q3);
}
template <int OP> void run(const char* name, double ops_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 128 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<148 * 8, 128>>>(out, 1.0001f, 1.0000001);
    cudaEventRecord(a);
    k<OP><<<148 * 8, 128>>>(out, 1.0001f, 1.0000001);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double warp_inst = 148.0 * 8 * 4 * ITER * ops_per_iter;
    double clk = ms * 1e-3 * khz * 1e3;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SM (of the counted op)\n", name, ms, warp_inst / clk / 148.0);
    cudaFree(out);
}
int main() {
    run<0>("FFMA x4", 4);
    run<1>("DADD x4", 4);
    run<6>("DFMA x4", 4);
    run<2>("F2F f32->f64 x4 (+1 back)", 5);
    run<3>("F2F f64->f32 x4 (+4 DADD)", 4);
    run<4>("fdiv_rn x4 (+4 FADD)", 4);
    run<7>("F2F f32->f64 x4 (+4 FADD)", 4);
    run<5>("SHFL x2", 2);
    run<11>("FADD x4", 4);
    run<8>("FFMA2 x4", 4);
    run<9>("FADD2 x4", 4);
    run<10>("FFMA2 x2 + FFMA x2", 4);
    run<12>("FFMA2 x2 + int x4 (count FFMA2)", 2);
    run<13>("FFMA x2 + int x4 (count FFMA)", 2);
    return 0;
}
