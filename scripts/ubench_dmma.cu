// FP64 tensor-core (DMMA, mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16) issue-rate micro-benchmarks on B200, alone and
// mixed with the MUFU / FP32 work of the fused likelihood kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dmma ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1024

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int MODE>
__global__ void k(const double* in, double* out) {
    const int t = threadIdx.x;
    double a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = in[t + i]; b[i] = in[t + 8 + i]; }
    double d2[8][2], d4[4][4];
    float fa[8], fb[8];
    for (int i = 0; i < 8; ++i) { d2[i][0] = d2[i][1] = 0.0; fa[i] = (float)a[i]; fb[i] = (float)b[i] * 1e-3f; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d4[i][j] = 0.0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {           // 8 independent m8n8k4
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(d2[i], a[i], b[i]);
        }
        if (MODE == 1) {           // 4 independent m16n8k4
#pragma unroll
            for (int i = 0; i < 4; ++i) { double aa[2] = {a[i], a[i + 4]}; dmma1684(d4[i], aa, b[i]); }
        }
        if (MODE == 2) {           // 4 independent m16n8k8
#pragma unroll
            for (int i = 0; i < 4; ++i) { double aa[4] = {a[i], a[i + 4], a[(i + 1) & 7], a[(i + 5) & 7]}; double bb[2] = {b[i], b[i + 4]}; dmma1688(d4[i], aa, bb); }
        }
        if (MODE == 3) {           // 4 independent m16n8k16
#pragma unroll
            for (int i = 0; i < 4; ++i) { double bb[4] = {b[i], b[i + 4], b[(i + 1) & 7], b[(i + 5) & 7]}; dmma16816(d4[i], a, bb); }
        }
        if (MODE == 4) {           // kernel-like: 3 x m8n8k4 per 16 (element,detector-pair) MUFU pairs ... 1 DMMA : 32 MUFU
#pragma unroll
            for (int i = 0; i < 2; ++i) dmma884(d2[i], a[i], b[i]);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i]));
        }
        if (MODE == 5) {           // 2 DMMA : 64 MUFU : 128 FFMA
#pragma unroll
            for (int i = 0; i < 2; ++i) dmma884(d2[i], a[i], b[i]);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i]));
                    fb[i] = fmaf(fb[i], fa[(i + 1) & 7], fb[(i + 2) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 3) & 7], fb[(i + 5) & 7]);
                }
        }
        if (MODE == 6) {           // 64 MUFU : 128 FFMA (no DMMA)
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i]));
                    fb[i] = fmaf(fb[i], fa[(i + 1) & 7], fb[(i + 2) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 3) & 7], fb[(i + 5) & 7]);
                }
        }
        if (MODE == 7) {           // 12 DMMA : 48 MUFU : 48 DFMA(detector angle) : 200 FFMA  == 8 bins x 32 walkers of the kernel
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(d2[i], a[i], b[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma884(d2[i], a[i + 4], b[i + 4]);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a[i] = fma(a[i], b[0], d2[i][r & 1]);
                    asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i]));
                    fb[i] = fmaf(fb[i], fa[(i + 1) & 7], fb[(i + 2) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 3) & 7], fb[(i + 5) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 4) & 7], fb[(i + 6) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 5) & 7], fb[(i + 7) & 7]);
                }
        }
        if (MODE >= 10 && MODE <= 13) {   // 1 DMMA + N independent FFMA (N = 0, 8, 16, 32): does the FP64 tensor op block dispatch?
            constexpr int N = MODE == 10 ? 0 : MODE == 11 ? 8 : MODE == 12 ? 16 : 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                dmma884(d2[g], a[g], b[g]);
#pragma unroll
                for (int i = 0; i < N; ++i) fa[i & 7] = fmaf(fa[i & 7], fb[0], fb[1]);
            }
        }
        if (MODE == 14) {                 // DFMA (shared multiplier) + FFMA (shared operands), 32 pairs
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) { a[i] = fma(a[i], b[0], b[1]); fa[i] = fmaf(fa[i], fb[0], fb[1]); }
        }
        if (MODE == 15) {                 // DFMA (shared multiplier) + 2 FFMA
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) { a[i] = fma(a[i], b[0], b[1]); fa[i] = fmaf(fa[i], fb[0], fb[1]); fb[i & 3 | 4] = fmaf(fb[i & 3 | 4], fb[2], fb[3]); }
        }
        if (MODE == 16) {                 // 1 DMMA + 4 DFMA, x4
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                dmma884(d2[g], a[g], b[g]);
#pragma unroll
                for (int i = 0; i < 4; ++i) a[4 + i] = fma(a[4 + i], b[0], b[1]);
            }
        }
        if (MODE == 17) {                 // 12 DMMA + 48 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(d2[i], a[i], b[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma884(d2[i], a[i + 4], b[i + 4]);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i]));
        }
        if (MODE == 18) {                 // 48 DFMA + 48 MUFU + 192 FFMA (mode 7 without the DMMA)
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a[i] = fma(a[i], b[0], d2[i][r & 1]);
                    asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i]));
                    fb[i] = fmaf(fb[i], fa[(i + 1) & 7], fb[(i + 2) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 3) & 7], fb[(i + 5) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 4) & 7], fb[(i + 6) & 7]);
                    fb[i] = fmaf(fb[i], fa[(i + 5) & 7], fb[(i + 7) & 7]);
                }
        }
        if (MODE == 19) {                 // 12 DMMA + 192 FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(d2[i], a[i], b[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma884(d2[i], a[i + 4], b[i + 4]);
#pragma unroll
            for (int r = 0; r < 24; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) fb[i] = fmaf(fb[i], fa[(i + 1) & 7], fb[(i + 2) & 7]);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += d2[i][0] + d2[i][1] + fa[i] + fb[i] + a[i];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d4[i][j];
    out[blockIdx.x * blockDim.x + t] = s;
}

template <int MODE> void run(const char* name, double fma_per_thread_iter, double groups_per_iter, const double* in, double* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 1; w <= 8; w *= 2) {     // warps per SMSP
        int threads = 128, blocks = sms * w;
        k<MODE><<<blocks, threads>>>(in, out);
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(in, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)threads * blocks * ITERS * fma_per_thread_iter;
        double cyc_per_group = (ms * 1e-3 * 1.93e9) / (ITERS * groups_per_iter * w);   // SMSP cycles per warp-group
        printf("%-44s warps/SMSP=%d  %.3f ms  %.3e FMA/s (%.1f FMA/clk/SM)  %.2f SMSP-cycles per warp-group\n", name, w, ms,
               fma / (ms * 1e-3), fma / (ms * 1e-3) / sms / 1.93e9, cyc_per_group);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 148 * 16 * 128 * 8 * 2);
    double h[512]; for (int i = 0; i < 512; ++i) h[i] = 1e-3 * (1.0 + 1e-3 * i); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount;
    run<0>("DMMA m8n8k4 x8 (group = 1 mma)", 8 * 8.0, 8, in, out, sms);
    run<1>("DMMA m16n8k4 x4 (group = 1 mma)", 4 * 16.0, 4, in, out, sms);
    run<2>("DMMA m16n8k8 x4 (group = 1 mma)", 4 * 32.0, 4, in, out, sms);
    run<3>("DMMA m16n8k16 x4 (group = 1 mma)", 4 * 64.0, 4, in, out, sms);
    run<4>("2 DMMA884 + 64 MUFU (group = all)", 2 * 8.0, 1, in, out, sms);
    run<5>("2 DMMA884 + 64 MUFU + 128 FFMA (group = all)", 2 * 8.0, 1, in, out, sms);
    run<6>("64 MUFU + 128 FFMA (group = all)", 0, 1, in, out, sms);
    run<7>("12 DMMA884+48 DFMA+48 MUFU+192 FFMA (=8 bins)", 12 * 8.0 + 48, 1, in, out, sms);
    run<10>("4 x (1 DMMA)            (group = 1 mma)", 4 * 8.0, 4, in, out, sms);
    run<11>("4 x (1 DMMA +  8 FFMA)  (group = 1 mma)", 4 * 8.0, 4, in, out, sms);
    run<12>("4 x (1 DMMA + 16 FFMA)  (group = 1 mma)", 4 * 8.0, 4, in, out, sms);
    run<13>("4 x (1 DMMA + 32 FFMA)  (group = 1 mma)", 4 * 8.0, 4, in, out, sms);
    run<14>("DFMA + FFMA, shared operands (group = pair)", 32, 32, in, out, sms);
    run<15>("DFMA + 2 FFMA, shared operands (group = triple)", 32, 32, in, out, sms);
    run<16>("4 x (1 DMMA + 4 DFMA)   (group = 1 mma)", 4 * 12.0, 4, in, out, sms);
    run<17>("12 DMMA + 48 MUFU (group = all)", 96, 1, in, out, sms);
    run<18>("48 DFMA + 48 MUFU + 192 FFMA (group = all)", 48, 1, in, out, sms);
    run<19>("12 DMMA + 192 FFMA (group = all)", 96, 1, in, out, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
