// fe29_bench.cu -- dev microbenchmark: GF(2^255-19) on 9 unsaturated 29-bit limbs (plain IMAD.WIDE.U32
// column sums, no carry flags) versus the 8x32 saturated representation (IMAD.WIDE.U32.X carry chains),
// in the same (4 doublings + 1 addition) window step of the range kernel.  Also measures the raw
// throughput of the IMAD.WIDE flavours.  JSON lines on stdout.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../anonymous-credit-tokens_b200/csrc/ge25519.cuh"

#define NT 128
#ifndef ITERS
#define ITERS 256
#endif

// ------------------------------------------------------------------ 9 x 29-bit limbs
struct f29 { u32 v[9]; };
#define M29 0x1fffffffu
// 2^261 = 1216 (mod p)
__device__ __forceinline__ f29 f29_add(const f29& a, const f29& b) {
    f29 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] = a.v[i] + b.v[i];
    return r;
}
// a - b + 2*(2^261 - 1216): needs b limbs <= 2^30 - 2432
__device__ __forceinline__ f29 f29_sub(const f29& a, const f29& b) {
    f29 r;
    r.v[0] = a.v[0] + (0x40000000u - 2432u) - b.v[0];
#pragma unroll
    for (int i = 1; i < 9; i++) r.v[i] = a.v[i] + (0x40000000u - 2u) - b.v[i];
    return r;
}
// weak normalisation of limbs < 2^32: every limb < 2^29 + 2^14 afterwards
__device__ __forceinline__ f29 f29_norm(const f29& a) {
    f29 r;
    u32 c8 = a.v[8] >> 29;
    r.v[0] = (a.v[0] & M29) + 1216u * c8;
#pragma unroll
    for (int i = 1; i < 9; i++) r.v[i] = (a.v[i] & M29) + (a.v[i - 1] >> 29);
    return r;
}
__device__ __forceinline__ f29 f29_carry(u64* c) {
    // fold columns 9..16: c[k] * 2^(29k) = 1216 * lo32 * 2^(29(k-9)) + 9728 * hi32 * 2^(29(k-8))
#pragma unroll
    for (int k = 9; k < 17; k++) {
        u32 lo = (u32)c[k], hi = (u32)(c[k] >> 32);
        c[k - 9] += (u64)lo * 1216u;
        c[k - 8] += (u64)hi * 9728u;
    }
    f29 r;
    u64 t = c[0];
    r.v[0] = (u32)t & M29; t >>= 29;
#pragma unroll
    for (int k = 1; k < 9; k++) { t += c[k]; r.v[k] = (u32)t & M29; t >>= 29; }
    // t < 2^36: wrap around
    u64 u = (u64)r.v[0] + (u64)(u32)t * 1216u + (((u64)(u32)(t >> 32) * 1216u) << 32);
    r.v[0] = (u32)u & M29;
    r.v[1] += (u32)(u >> 29);
    return r;
}
__device__ __forceinline__ f29 f29_mul_inl(const f29& a, const f29& b) {
    u64 c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) c[i + j] += (u64)a.v[i] * b.v[j];
    return f29_carry(c);
}
__device__ __forceinline__ f29 f29_sq_inl(const f29& a) {
    u64 c[17];
    u32 a2[9];
#pragma unroll
    for (int i = 0; i < 9; i++) a2[i] = a.v[i] << 1;
#pragma unroll
    for (int k = 0; k < 17; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c[2 * i] += (u64)a.v[i] * a.v[i];
#pragma unroll
        for (int j = i + 1; j < 9; j++) c[i + j] += (u64)a2[i] * a.v[j];
    }
    return f29_carry(c);
}
__device__ __noinline__ f29 f29_mul(f29 a, f29 b) { return f29_mul_inl(a, b); }
__device__ __noinline__ f29 f29_sq(f29 a) { return f29_sq_inl(a); }

__device__ f29 f29_from_fe(const fe& s) {   // s < 2^256
    f29 r;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        int bit = 29 * k, w = bit >> 5, sh = bit & 31;
        u64 x = s.v[w];
        if (w + 1 < 8) x |= (u64)s.v[w + 1] << 32;
        r.v[k] = (u32)(x >> sh) & M29;
    }
    return r;
}
__device__ fe f29_to_fe(const f29& a) {
    // full carry, fold bits >= 256 (2^256 = 38), pack
    u64 t = 0;
    u32 l[9];
    for (int k = 0; k < 9; k++) { t += a.v[k]; l[k] = (u32)t & M29; t >>= 29; }
    // value = sum l[k] 2^(29k) + t * 2^261
    u32 w[9];
    for (int i = 0; i < 9; i++) w[i] = 0;
    for (int k = 0; k < 9; k++) {
        int bit = 29 * k, wi = bit >> 5, sh = bit & 31;
        u64 x = (u64)l[k] << sh;
        w[wi] |= (u32)x;
        w[wi + 1] |= (u32)(x >> 32);
    }
    // w[8] holds bits 256..260; plus t*2^261 = t * 32 * 2^256
    u64 hi = (u64)w[8] + t * 32u;
    u64 c = hi * 38u;
    fe r;
    for (int i = 0; i < 8; i++) { c += w[i]; r.v[i] = (u32)c; c >>= 32; }
    c *= 38u;
    for (int i = 0; i < 8; i++) { c += r.v[i]; r.v[i] = (u32)c; c >>= 32; }
    return r;
}
struct g29 { f29 X, Y, Z, T; };
struct g29c { f29 YpX, YmX, Z, T2d; };
__device__ __forceinline__ g29 g29_dbl(const g29& p, bool want_t) {
    f29 XX = f29_sq(p.X), YY = f29_sq(p.Y), ZZ = f29_sq(p.Z);
    f29 XpY2 = f29_sq(f29_add(p.X, p.Y));
    f29 Yc = f29_norm(f29_add(YY, XX)), Zc = f29_sub(YY, XX);
    f29 Xc = f29_sub(XpY2, Yc);
    f29 Tc = f29_norm(f29_sub(f29_add(f29_add(ZZ, ZZ), XX), YY));
    g29 r;
    r.X = f29_mul(Xc, Tc); r.Y = f29_mul(Zc, Yc); r.Z = f29_mul(Zc, Tc);
    r.T = p.T;
    if (want_t) r.T = f29_mul(Xc, Yc);
    return r;
}
__device__ __forceinline__ g29 g29_add(const g29& p, const g29c& q) {
    f29 PP = f29_mul(f29_add(p.Y, p.X), q.YpX);
    f29 MM = f29_mul(f29_sub(p.Y, p.X), q.YmX);
    f29 TT = f29_mul(p.T, q.T2d);
    f29 ZZ = f29_mul(p.Z, q.Z);
    f29 ZZ2 = f29_add(ZZ, ZZ);
    f29 E = f29_sub(PP, MM), H = f29_add(PP, MM), G = f29_add(ZZ2, TT), F = f29_norm(f29_sub(ZZ2, TT));
    g29 r;
    r.X = f29_mul(E, F); r.Y = f29_mul(G, H); r.Z = f29_mul(G, F); r.T = f29_mul(E, H);
    return r;
}

__device__ ge start_point(u32 salt) {
    ge B = ge_basepoint();
    ge P = B;
    for (u32 i = 0; i < (salt & 7u) + 1; i++) P = ge_add(ge_dbl_t(P), B);
    return P;
}
template <int BPS, int WITH_ADD>
__global__ void __launch_bounds__(NT, BPS) k_sat(u32* out) {
    ge P = start_point(threadIdx.x + blockIdx.x);
    ge_cached q = ge_to_cached(start_point(threadIdx.x * 3 + 1));
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
        for (int d = 0; d < 4; d++) P = ge_dbl_x<false>(P, d == 3);
        if (WITH_ADD) P = ge_add_cached_x<false>(P, q);
    }
    u32 w[8];
    ristretto_encode_(w, &P);
    for (int i = 0; i < 8; i++) out[(blockIdx.x * NT + threadIdx.x) * 8 + i] = w[i];
}
template <int BPS, int WITH_ADD>
__global__ void __launch_bounds__(NT, BPS) k_29(u32* out) {
    g29 P;
    g29c q;
    {
        ge P0 = start_point(threadIdx.x + blockIdx.x);
        ge_cached q0 = ge_to_cached(start_point(threadIdx.x * 3 + 1));
        P.X = f29_from_fe(P0.X); P.Y = f29_from_fe(P0.Y); P.Z = f29_from_fe(P0.Z); P.T = f29_from_fe(P0.T);
        q.YpX = f29_from_fe(q0.YpX); q.YmX = f29_from_fe(q0.YmX); q.Z = f29_from_fe(q0.Z); q.T2d = f29_from_fe(q0.T2d);
    }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
        for (int d = 0; d < 4; d++) P = g29_dbl(P, d == 3);
        if (WITH_ADD) P = g29_add(P, q);
    }
    ge R;
    R.X = f29_to_fe(P.X); R.Y = f29_to_fe(P.Y); R.Z = f29_to_fe(P.Z); R.T = f29_to_fe(P.T);
    u32 w[8];
    ristretto_encode_(w, &R);
    for (int i = 0; i < 8; i++) out[(blockIdx.x * NT + threadIdx.x) * 8 + i] = w[i];
}

// ------------------------------------------------------------------ raw IMAD.WIDE flavours
#define WITER 2048
template <int MODE>
__global__ void __launch_bounds__(256) k_wide(u32* out, u32 s) {
    u32 a[8], b[8];
    u32 acc[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = s * (threadIdx.x + i + 1); b[i] = s + blockIdx.x * 7 + i; }
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = i;
#pragma unroll 1
    for (int it = 0; it < WITER; it++) {
        if (MODE == 0) {   // 8 independent plain multiply-adds, distinct operands
            u64* q = reinterpret_cast<u64*>(acc);
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[i]) : "r"(a[i]), "r"(b[7 - i]));
        }
        if (MODE == 1) {   // two carry chains of 4 (cc-out, X, X, X-in-only)
            asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                         "madc.lo.cc.u32 %2, %9, %13, %2;\n\tmadc.hi.cc.u32 %3, %9, %13, %3;\n\t"
                         "madc.lo.cc.u32 %4, %10, %14, %4;\n\tmadc.hi.cc.u32 %5, %10, %14, %5;\n\t"
                         "madc.lo.cc.u32 %6, %11, %15, %6;\n\tmadc.hi.u32 %7, %11, %15, %7;"
                         : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
            asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                         "madc.lo.cc.u32 %2, %9, %13, %2;\n\tmadc.hi.cc.u32 %3, %9, %13, %3;\n\t"
                         "madc.lo.cc.u32 %4, %10, %14, %4;\n\tmadc.hi.cc.u32 %5, %10, %14, %5;\n\t"
                         "madc.lo.cc.u32 %6, %11, %15, %6;\n\tmadc.hi.u32 %7, %11, %15, %7;"
                         : "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(acc[12]), "+r"(acc[13]), "+r"(acc[14]), "+r"(acc[15])
                         : "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
        }
        if (MODE == 2) {   // 8 independent multiply-adds each producing a carry-out that nobody chains (cc-out only)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u32 cy;
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, 0, 0;"
                             : "+r"(acc[2 * i]), "+r"(acc[2 * i + 1]), "=r"(cy) : "r"(a[i]), "r"(b[7 - i]));
                a[i] ^= cy;
            }
        }
    }
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= acc[i];
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename F>
static double best_ms(F launch, int reps = 3) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 1e30;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); exit(1); }
    return best;
}

template <int BPS, int WITH_ADD>
static void compare(int sms, u32* d_out, u32* h1, u32* h2) {
    int grid = sms * BPS * 4;
    size_t n = (size_t)grid * NT * 8;
    double tr = best_ms([&] { k_sat<BPS, WITH_ADD><<<grid, NT>>>(d_out); });
    cudaMemcpy(h1, d_out, n * 4, cudaMemcpyDeviceToHost);
    double ts = best_ms([&] { k_29<BPS, WITH_ADD><<<grid, NT>>>(d_out); });
    cudaMemcpy(h2, d_out, n * 4, cudaMemcpyDeviceToHost);
    int same = 1;
    for (size_t i = 0; i < n; i++) if (h1[i] != h2[i]) { same = 0; break; }
    double steps = (double)grid * NT * ITERS;
    printf("{\"bps\": %d, \"with_add\": %d, \"sat_ms\": %.3f, \"f29_ms\": %.3f, \"sat_Gsteps\": %.3f, \"f29_Gsteps\": %.3f, \"speedup\": %.3f, \"same\": %d}\n",
           BPS, WITH_ADD, tr, ts, steps / tr * 1e-6, steps / ts * 1e-6, tr / ts, same);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    size_t maxn = (size_t)sms * 8 * 4 * NT * 8;
    u32* d_out;
    cudaMalloc(&d_out, maxn * 4 > (size_t)sms * 8 * 256 * 4 ? maxn * 4 : (size_t)sms * 8 * 256 * 4);
    u32* h1 = (u32*)malloc(maxn * 4);
    u32* h2 = (u32*)malloc(maxn * 4);
    int blocks = sms * 8;
    double m0 = best_ms([&] { k_wide<0><<<blocks, 256>>>(d_out, 3); });
    double m1 = best_ms([&] { k_wide<1><<<blocks, 256>>>(d_out, 3); });
    double m2 = best_ms([&] { k_wide<2><<<blocks, 256>>>(d_out, 3); });
    double ops = (double)blocks * 256 * WITER * 8;
    printf("{\"wide_plain_T\": %.3f, \"wide_chain_T\": %.3f, \"wide_ccout_T\": %.3f}\n", ops / m0 * 1e-9, ops / m1 * 1e-9, ops / m2 * 1e-9);
    compare<3, 1>(sms, d_out, h1, h2);
    compare<4, 1>(sms, d_out, h1, h2);
    compare<4, 0>(sms, d_out, h1, h2);
    compare<5, 1>(sms, d_out, h1, h2);
    return 0;
}
