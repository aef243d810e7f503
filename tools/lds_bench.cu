// lds_bench.cu -- dev microbenchmark: point doubling/addition throughput with field elements passed
//   (R) in registers through the fe_mul / fe_sq call ABI (what the engine does), versus
//   (S) in shared-memory slots: fe_mul_s(dst, a, b) loads its operands with LDS.128 and stores the
//       result with STS.128, so no register marshalling moves land on the integer-multiply pipe.
// Prints JSON lines with doublings/s and (4 dbl + 1 add)/s for several blocks-per-SM settings and
// checks that both formulations give identical results.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../anonymous-credit-tokens_b200/csrc/ge25519.cuh"

#define NT 128
#ifndef ITERS
#define ITERS 256
#endif

// ---------------- shared-memory slot machine ----------------
// slot s of thread t: two uint4 at sm[(2s+h)*NT + t]  (conflict-free 128-bit accesses)
__device__ __forceinline__ fe lds_fe(const uint4* base, u32 slot) {
    uint4 a = base[(2 * slot) * NT], b = base[(2 * slot + 1) * NT];
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fe(uint4* base, u32 slot, const fe& f) {
    base[(2 * slot) * NT] = make_uint4(f.v[0], f.v[1], f.v[2], f.v[3]);
    base[(2 * slot + 1) * NT] = make_uint4(f.v[4], f.v[5], f.v[6], f.v[7]);
}
// d = op(a1, a2) * b     mode 0: a1, 1: a1 + a2, 2: a1 - a2
__device__ __noinline__ void fe_mul_s(uint4* base, u32 d, u32 a1, u32 a2, u32 b, u32 mode) {
    fe x = lds_fe(base, a1), y = lds_fe(base, b);
    if (mode) {
        fe z = lds_fe(base, a2);
        x = (mode == 1) ? fe_add(x, z) : fe_sub(x, z);
    }
    sts_fe(base, d, fe_mul_inl(x, y));
}
// d = op(a1, a2)^2
__device__ __noinline__ void fe_sq_s(uint4* base, u32 d, u32 a1, u32 a2, u32 mode) {
    fe x = lds_fe(base, a1);
    if (mode) x = fe_add(x, lds_fe(base, a2));
    sts_fe(base, d, fe_sq_inl(x));
}
// doubling middle: A=XX B=YY C=ZZ D=(X+Y)^2  ->  A:=Yc=B+A, B:=Zc=B-A, D:=Xc=D-Yc, C:=Tc=2C-Zc
__device__ __noinline__ void dbl_mid_s(uint4* base, u32 A, u32 B, u32 Cc, u32 D) {
    fe a = lds_fe(base, A), b = lds_fe(base, B), c = lds_fe(base, Cc), d = lds_fe(base, D);
    fe yc = fe_add(b, a), zc = fe_sub(b, a);
    fe xc = fe_sub(d, yc), tc = fe_sub(fe_add(c, c), zc);
    sts_fe(base, A, yc); sts_fe(base, B, zc); sts_fe(base, D, xc); sts_fe(base, Cc, tc);
}
// addition middle: PP,MM,TT,ZZ -> PP:=E=PP-MM, MM:=H=PP+MM, TT:=G=2ZZ+TT, ZZ:=F=2ZZ-TT
__device__ __noinline__ void add_mid_s(uint4* base, u32 PP, u32 MM, u32 TT, u32 ZZ) {
    fe pp = lds_fe(base, PP), mm = lds_fe(base, MM), tt = lds_fe(base, TT), zz = lds_fe(base, ZZ);
    fe zz2 = fe_add(zz, zz);
    sts_fe(base, PP, fe_sub(pp, mm)); sts_fe(base, MM, fe_add(pp, mm));
    sts_fe(base, TT, fe_add(zz2, tt)); sts_fe(base, ZZ, fe_sub(zz2, tt));
}
// slots: 0..3 = X,Y,Z,T of the running point; 4..7 = YpX,YmX,Zq,T2d of the addend; 8..11 temporaries
enum { SX = 0, SY, SZ, ST, QP, QM, QZ, QT, T0, T1, T2, T3, NSLOT };
__device__ __forceinline__ void dbl_s(uint4* b, bool want_t) {
    fe_sq_s(b, T0, SX, 0, 0); fe_sq_s(b, T1, SY, 0, 0); fe_sq_s(b, T2, SZ, 0, 0); fe_sq_s(b, T3, SX, SY, 1);
    dbl_mid_s(b, T0, T1, T2, T3);                 // T0=Yc T1=Zc T3=Xc T2=Tc
    fe_mul_s(b, SX, T3, 0, T2, 0); fe_mul_s(b, SY, T0, 0, T1, 0); fe_mul_s(b, SZ, T1, 0, T2, 0);
    if (want_t) fe_mul_s(b, ST, T3, 0, T0, 0);
}
__device__ __forceinline__ void add_s(uint4* b) {
    fe_mul_s(b, T0, SY, SX, QP, 1); fe_mul_s(b, T1, SY, SX, QM, 2); fe_mul_s(b, T2, ST, 0, QT, 0); fe_mul_s(b, T3, SZ, 0, QZ, 0);
    add_mid_s(b, T0, T1, T2, T3);                 // T0=E T1=H T2=G T3=F
    fe_mul_s(b, SX, T0, 0, T3, 0); fe_mul_s(b, SY, T1, 0, T2, 0); fe_mul_s(b, SZ, T2, 0, T3, 0); fe_mul_s(b, ST, T0, 0, T1, 0);
}

__device__ ge start_point(u32 salt) {
    ge B = ge_basepoint();
    ge P = B;
    for (u32 i = 0; i < (salt & 7u) + 1; i++) P = ge_add(ge_dbl_t(P), B);
    return P;
}

template <int BPS, int WITH_ADD>
__global__ void __launch_bounds__(NT, BPS) k_reg(u32* out) {
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
__global__ void __launch_bounds__(NT, BPS) k_smem(u32* out) {
    extern __shared__ uint4 sm[];
    uint4* b = sm + threadIdx.x;
    {
        ge P = start_point(threadIdx.x + blockIdx.x);
        ge_cached q = ge_to_cached(start_point(threadIdx.x * 3 + 1));
        sts_fe(b, SX, P.X); sts_fe(b, SY, P.Y); sts_fe(b, SZ, P.Z); sts_fe(b, ST, P.T);
        sts_fe(b, QP, q.YpX); sts_fe(b, QM, q.YmX); sts_fe(b, QZ, q.Z); sts_fe(b, QT, q.T2d);
    }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
        for (int d = 0; d < 4; d++) dbl_s(b, d == 3);
        if (WITH_ADD) add_s(b);
    }
    ge P;
    P.X = lds_fe(b, SX); P.Y = lds_fe(b, SY); P.Z = lds_fe(b, SZ); P.T = lds_fe(b, ST);
    u32 w[8];
    ristretto_encode_(w, &P);
    for (int i = 0; i < 8; i++) out[(blockIdx.x * NT + threadIdx.x) * 8 + i] = w[i];
}

template <typename K>
static double run(K kern, int grid, size_t smem, u32* d_out, int reps = 3) {
    if (smem) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 1e30;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(a);
        kern<<<grid, NT, smem>>>(d_out);
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
    double tr = run(k_reg<BPS, WITH_ADD>, grid, 0, d_out);
    cudaMemcpy(h1, d_out, n * 4, cudaMemcpyDeviceToHost);
    size_t smem = (size_t)NSLOT * 2 * NT * 16;
    double ts = run(k_smem<BPS, WITH_ADD>, grid, smem, d_out);
    cudaMemcpy(h2, d_out, n * 4, cudaMemcpyDeviceToHost);
    int same = 1;
    for (size_t i = 0; i < n; i++) if (h1[i] != h2[i]) { same = 0; break; }
    double steps = (double)grid * NT * ITERS;
    printf("{\"bps\": %d, \"with_add\": %d, \"reg_ms\": %.3f, \"smem_ms\": %.3f, \"reg_Gsteps\": %.3f, \"smem_Gsteps\": %.3f, \"speedup\": %.3f, \"same\": %d}\n",
           BPS, WITH_ADD, tr, ts, steps / tr * 1e-6, steps / ts * 1e-6, tr / ts, same);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    size_t maxn = (size_t)sms * 6 * 4 * NT * 8;
    u32* d_out;
    cudaMalloc(&d_out, maxn * 4);
    u32* h1 = (u32*)malloc(maxn * 4);
    u32* h2 = (u32*)malloc(maxn * 4);
    compare<3, 0>(sms, d_out, h1, h2);
    compare<4, 0>(sms, d_out, h1, h2);
    compare<3, 1>(sms, d_out, h1, h2);
    compare<4, 1>(sms, d_out, h1, h2);
    return 0;
}
