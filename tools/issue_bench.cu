// issue_bench.cu -- how many integer instructions per clock can one SM sub-partition issue?
// Independent accumulator streams (no dependency within 8 instructions), 16 warps per SM (4 per sub-partition) or 32,
// operands varied so that nothing folds.  Modes: 0 IADD3 (3 register sources)  1 IADD (2 register sources)
// 2 LOP3 (3 sources)  3 IMAD.WIDE.U32 with a 64-bit addend  4 alternating IADD3 / IMAD.WIDE (the mix of a field multiply)
// 5 LOP3 with 2 register sources  6 IMAD (32-bit) with 3 register sources  7 IMAD.WIDE without addend  8-11 other mixes;
// warpsN = N warps per sub-partition
// Output: one JSON object; inst/clk/SMSP = instructions / (elapsed x SM clock x SMs x 4).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef uint32_t u32; typedef unsigned long long u64;
#define ITER 2048
template <int MODE>
__global__ void __launch_bounds__(128) k(u32* out, u32 a0, u32 b0) {
    u32 a = a0 + threadIdx.x, b = b0 ^ blockIdx.x, c = a0 * 3u + 1u;
    u32 r[8]; u64 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = i + a; w[i] = i + b; }
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                // every instruction reads results of the previous round (8 instructions back): nothing folds, ILP = 8
                u32 x = r[(i + 1) & 7], y = r[(i + 2) & 7];
                if (MODE == 0) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(r[i]) : "r"(x), "r"(y));   // IADD3, 3 register sources
                if (MODE == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(x));                                          // IADD3 with RZ, 2 sources
                if (MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(x), "r"(y));                        // LOP3, 3 sources
                if (MODE == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 1) & 7]), "r"((u32)(w[(i + 2) & 7] >> 32)));   // IMAD.WIDE, 4 source words
                if (MODE == 4) { if (i & 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 2) & 7]), "r"((u32)(w[(i + 4) & 7] >> 32)));
                                 else asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(r[i]) : "r"(x), "r"(y)); }
                if (MODE == 5) asm volatile("xor.b32 %0, %1, %2;" : "=r"(r[i]) : "r"(x), "r"(r[(i + 3) & 7]));                      // LOP3, 2 sources
                if (MODE == 6) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(x), "r"(y));                            // IMAD, 3 sources
                if (MODE == 7) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[(i + 1) & 7]), "r"((u32)(w[(i + 2) & 7] >> 32)));   // IMAD.WIDE, no addend
                if (MODE == 8) { if ((i & 3) == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 4) & 7]), "r"(x));
                                 else asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(r[i]) : "r"(x), "r"(y)); }        // 3 IADD3 : 1 IMAD.WIDE
                if (MODE == 9) { if ((i & 3) == 0) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(r[i]) : "r"(x), "r"(y));
                                 else asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 1) & 7]), "r"((u32)(w[(i + 2) & 7] >> 32))); } // 1 IADD3 : 3 IMAD.WIDE
                if (MODE == 10) { if (i & 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 2) & 7]), "r"((u32)(w[(i + 4) & 7] >> 32)));
                                  else asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(x)); }                                          // 1 two-source add : 1 IMAD.WIDE
                if (MODE == 11) { if (i & 1) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[(i + 2) & 7]), "r"((u32)(w[(i + 4) & 7] >> 32)));
                                  else asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(r[i]) : "r"(x), "r"(y)); }      // 1 IADD3 : 1 IMAD.WIDE without addend
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += w[i] + r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)s ^ (u32)(s >> 32);
}
template <typename F> static double time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u32* out; cudaMalloc(&out, 148 * 8 * 128 * 4 * 2);
    const char* names[12] = {"iadd3_3reg", "iadd_2reg", "lop3_3reg", "imad_wide_64acc", "mix_iadd3_imadwide", "lop3_2reg", "imad_3reg",
                             "imad_wide_noacc", "mix_3iadd3_1wide", "mix_1iadd3_3wide", "mix_add2_wide", "mix_iadd3_wide_noacc"};
    printf("{\"sms\": %d, \"clock_khz\": %d", p.multiProcessorCount, clk_khz);
    for (int bps = 4; bps <= 8; bps *= 2) {
        int grid = p.multiProcessorCount * bps;
        double inst = (double)grid * 4 /*warps*/ * ITER * 32.0;   // warp instructions of the measured kind per launch
        double ms[12];
        ms[0] = time_ms([&] { k<0><<<grid, 128>>>(out, 3, 5); });
        ms[1] = time_ms([&] { k<1><<<grid, 128>>>(out, 3, 5); });
        ms[2] = time_ms([&] { k<2><<<grid, 128>>>(out, 3, 5); });
        ms[3] = time_ms([&] { k<3><<<grid, 128>>>(out, 3, 5); });
        ms[4] = time_ms([&] { k<4><<<grid, 128>>>(out, 3, 5); });
        ms[5] = time_ms([&] { k<5><<<grid, 128>>>(out, 3, 5); });
        ms[6] = time_ms([&] { k<6><<<grid, 128>>>(out, 3, 5); });
        ms[7] = time_ms([&] { k<7><<<grid, 128>>>(out, 3, 5); });
        ms[8] = time_ms([&] { k<8><<<grid, 128>>>(out, 3, 5); });
        ms[9] = time_ms([&] { k<9><<<grid, 128>>>(out, 3, 5); });
        ms[10] = time_ms([&] { k<10><<<grid, 128>>>(out, 3, 5); });
        ms[11] = time_ms([&] { k<11><<<grid, 128>>>(out, 3, 5); });
        for (int m = 0; m < 12; m++)
            printf(", \"%s_warps%d\": %.4f", names[m], bps, inst / (ms[m] * 1e-3) / ((double)clk_khz * 1e3) / (p.multiProcessorCount * 4.0));
    }
    printf("}\n");
    return 0;
}
