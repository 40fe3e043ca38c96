// Operand-bandwidth micro-benchmarks (B200): how many SMSP cycles one warp instruction costs as a function of the number
// and width of its register operands.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_operands ubench_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

template <int MODE>
__global__ void k(const double* in, double* out) {
    const int t = threadIdx.x;
    double a[8], b[8], c[8];
    float fa[8], fb[8], fc[8];
    u64 pa[8], pb[8], pc[8];
    for (int i = 0; i < 8; ++i) {
        a[i] = in[t + i]; b[i] = in[t + 8 + i]; c[i] = in[t + 16 + i];
        fa[i] = (float)a[i]; fb[i] = (float)b[i]; fc[i] = (float)c[i] * 1e-3f;
        pa[i] = pack2(fa[i], fb[i]); pb[i] = pack2(fb[i], fc[i]); pc[i] = pack2(fc[i], fa[i]);
    }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) fa[i] = fmaf(fa[i], fb[i], fc[i]);                         // FFMA 3 distinct
                if (MODE == 1) fa[i] = fmaf(fa[i], fb[0], fc[i]);                         // FFMA shared multiplier
                if (MODE == 2) fa[i] = fmaf(fa[i], fb[0], fc[0]);                         // FFMA two shared
                if (MODE == 3) pa[i] = fma2(pa[i], pb[i], pc[i]);                         // FFMA2 3 distinct pairs
                if (MODE == 4) pa[i] = fma2(pa[i], pb[0], pc[i]);                         // FFMA2 shared multiplier pair
                if (MODE == 5) pa[i] = fma2(pa[i], pack2(fb[i], fb[i]), pc[i]);           // FFMA2 scalar-broadcast multiplier
                if (MODE == 6) pa[i] = fma2(pb[i], pack2(fb[0], fb[0]), pa[i]);           // acc = d * s(shared bcast) + acc
                if (MODE == 7) pa[i] = fma2(pa[i], pb[0], pc[0]);                         // FFMA2 two shared
                if (MODE == 8) a[i] = fma(a[i], b[i], c[i]);                              // DFMA 3 distinct
                if (MODE == 9) a[i] = fma(a[i], b[0], c[i]);                              // DFMA shared multiplier
                if (MODE == 10) a[i] = fma(a[i], b[i], c[0]);                             // DFMA shared addend
                if (MODE == 11) a[i] = fma(a[i], b[0], c[0]);                             // DFMA two shared
                if (MODE == 12) { a[i] = fma(a[i], b[0], c[i]); fa[i] = fmaf(fa[i], fb[i], fc[i]); }          // DFMA(4) + FFMA(3)
                if (MODE == 13) { a[i] = fma(a[i], b[0], c[i]); pa[i] = fma2(pa[i], pb[i], pc[i]); }          // DFMA(4) + FFMA2(6)
                if (MODE == 14) { a[i] = fma(a[i], b[0], c[i]); fa[i] = fmaf(fa[i], fb[i], fc[i]); fc[i] = fmaf(fc[i], fb[(i + 1) & 7], fa[(i + 2) & 7]); }
                if (MODE == 15) fa[i] = fa[i] * fb[i];                                    // FMUL 2 regs
                if (MODE == 16) fa[i] = fa[i] * 1.0000001f;                               // FMUL imm
                if (MODE == 17) a[i] = fma(a[i], 1.0000001, c[i]);                        // DFMA imm multiplier
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + fa[i] + fc[i] + __longlong_as_double(pa[i]);
    out[blockIdx.x * blockDim.x + t] = s;
}
template <int MODE> void run(const char* name, const double* in, double* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int w = 4;
    int threads = 128, blocks = sms * w;
    k<MODE><<<blocks, threads>>>(in, out);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(in, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-52s %.3f ms  %.2f SMSP-cycles per warp-group\n", name, ms, (ms * 1e-3 * 1.93e9) / (ITERS * 32.0 * w));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 148 * 16 * 128 * 8 * 2);
    double h[512]; for (int i = 0; i < 512; ++i) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount;
    run<0>("FFMA  r,r,r distinct", in, out, sms);
    run<1>("FFMA  shared multiplier", in, out, sms);
    run<2>("FFMA  two shared", in, out, sms);
    run<3>("FFMA2 3 distinct pairs", in, out, sms);
    run<4>("FFMA2 shared multiplier pair", in, out, sms);
    run<5>("FFMA2 scalar-broadcast multiplier (distinct)", in, out, sms);
    run<6>("FFMA2 acc = d * bcast(shared) + acc", in, out, sms);
    run<7>("FFMA2 two shared", in, out, sms);
    run<8>("DFMA  3 distinct", in, out, sms);
    run<9>("DFMA  shared multiplier", in, out, sms);
    run<10>("DFMA  shared addend", in, out, sms);
    run<11>("DFMA  two shared", in, out, sms);
    run<12>("DFMA shared-mult + FFMA distinct (pair)", in, out, sms);
    run<13>("DFMA shared-mult + FFMA2 distinct (pair)", in, out, sms);
    run<14>("DFMA shared-mult + 2 FFMA distinct (triple)", in, out, sms);
    run<15>("FMUL r,r", in, out, sms);
    run<16>("FMUL r,imm", in, out, sms);
    run<17>("DFMA imm multiplier", in, out, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
