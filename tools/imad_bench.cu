// imad_bench.cu -- measures the integer-multiply roofline of the GPU it runs on: sustained
// 32x32+64->64 multiply-add ("limb-MAC") rate of IMAD.WIDE.U32 chains, plus the achieved limb-MAC
// rate of this repo's fe_mul / fe_sq.  Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../anonymous-credit-tokens_b200/csrc/fe25519.cuh"

#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k_imad(u32* out, u32 a0, u32 b0) {
    u32 a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    u64 acc[8];
    u32 lo[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = i; lo[i] = i; }
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a), "r"(b));
            if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
            if (MODE == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a), "r"(b));
        }
        if (MODE == 3) {
            // carry-chained pairs exactly as in fe_row_chain (4 fused IMAD.WIDE.X + addc) x 2
            u32* p = reinterpret_cast<u32*>(acc);
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                         "madc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                         "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                         "madc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                         : "+r"(p[0]), "+r"(p[1]), "+r"(p[2]), "+r"(p[3]), "+r"(p[4]), "+r"(p[5]), "+r"(p[6]), "+r"(p[7]) : "r"(a), "r"(b));
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                         "madc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                         "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                         "madc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                         : "+r"(p[8]), "+r"(p[9]), "+r"(p[10]), "+r"(p[11]), "+r"(p[12]), "+r"(p[13]), "+r"(p[14]), "+r"(p[15]) : "r"(a), "r"(b));
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i] + lo[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)s ^ (u32)(s >> 32);
}
#define FE_ITER 512
template <int MODE>
__global__ void __launch_bounds__(256) k_fe(u32* out, u32 seed) {
    fe x[2];
    for (int k = 0; k < 2; k++) for (int i = 0; i < 8; i++) x[k].v[i] = seed * (threadIdx.x + 1 + k) + i * 0x9e3779b9u + blockIdx.x;
#pragma unroll 1
    for (int it = 0; it < FE_ITER; it++) {
        if (MODE == 0) { x[0] = fe_mul(x[0], x[1]); x[1] = fe_mul(x[1], x[0]); }
        else { x[0] = fe_sq(x[0]); x[1] = fe_sq(x[1]); }
    }
    u32 s = 0;
    for (int i = 0; i < 8; i++) s ^= x[0].v[i] ^ x[1].v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
static double time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    double best = 1e30;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    u32* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    double ops = (double)blocks * threads * ITER * 8;
    double t0 = time_ms([&] { k_imad<0><<<blocks, threads>>>(out, 3, 5); });
    double t1 = time_ms([&] { k_imad<1><<<blocks, threads>>>(out, 3, 5); });
    double t2 = time_ms([&] { k_imad<2><<<blocks, threads>>>(out, 3, 5); });
    double t3 = time_ms([&] { k_imad<3><<<blocks, threads>>>(out, 3, 5); });
    double fops = (double)blocks * threads * FE_ITER * 2;
    double t4 = time_ms([&] { k_fe<0><<<blocks, threads>>>(out, 7); });
    double t5 = time_ms([&] { k_fe<1><<<blocks, threads>>>(out, 7); });
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, "
           "\"imad_lo_gops\": %.1f, \"imad_wide_gops\": %.1f, \"imad_hi_gops\": %.1f, \"imad_wide_cc_gops\": %.1f, "
           "\"fe_mul_gops\": %.2f, \"fe_sq_gops\": %.2f, \"fe_mul_limbmac_g\": %.1f, \"fe_sq_limbmac_g\": %.1f}\n",
           p.name, sms, clk, ops / t0 / 1e6, ops / t1 / 1e6, ops / t2 / 1e6, ops / t3 / 1e6,
           fops / t4 / 1e6, fops / t5 / 1e6, fops * 72 / t4 / 1e6, fops * 72 / t5 / 1e6);
    return 0;
}
