// Issue-rate micro-benchmarks (B200): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void k(const double* in, double* out) {
    const int t = threadIdx.x;
    double a[8], b[8], c[8];
    float fa[8], fb[8];
    for (int i = 0; i < 8; ++i) { a[i] = in[t + i]; b[i] = in[t + 8 + i]; c[i] = in[t + 16 + i]; fa[i] = (float)a[i]; fb[i] = (float)b[i]; }
    int ia[8]; for (int i = 0; i < 8; ++i) ia[i] = (int)a[i];
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fma(a[i], b[i], c[i]);                     // 3 distinct register pairs
                if (MODE == 1) a[i] = fma(a[i], b[0], c[i]);                     // Horner-like: shared multiplier
                if (MODE == 2) a[i] = fma(a[i], 1.0000001, 1e-9);                // constants
                if (MODE == 3) { a[i] = fma(a[i], b[i], c[i]); fa[i] = fmaf(fa[i], fb[i], fb[(i+1)&7]); }  // DFMA + FFMA co-issue
                if (MODE == 4) { a[i] = (double)fa[i]; fa[i] = fa[i] + fb[i]; }  // F2F.F64.F32 (+FADD)
                if (MODE == 5) { fa[i] = (float)a[i]; a[i] = a[i] + b[i]; }      // F2F.F32.F64 (+DADD)
                if (MODE == 6) { fa[i] = __int2float_rn(ia[i]); ia[i] = ia[i] + __float_as_int(fa[i]); } // I2FP + IADD
                if (MODE == 7) { asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(*(unsigned long long*)&a[i]) : "l"(*(unsigned long long*)&b[i]), "l"(*(unsigned long long*)&c[i])); }
                if (MODE == 8) { a[i] = fma(a[i], b[i], c[i]); fa[i] = __int2float_rn(ia[i]); ia[i] += 3; }   // DFMA + I2FP + IADD
                if (MODE == 9) { a[i] = fma(a[i], b[i], c[i]); double d = (double)fa[i]; fa[i] += fb[i]; c[i] += d*1e-30; } // DFMA+F2F+FADD+DFMA
                if (MODE == 10) a[i] = a[i] * b[i];                              // DMUL 2 regs
                if (MODE == 12) { asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i])); }                       // MUFU.SIN
                if (MODE == 13) { a[i] = fma(a[i], b[0], c[i]); asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i])); }
                if (MODE == 14) { a[i] = fma(a[i], b[0], c[i]); c[i] = fma(c[i], b[1], a[(i+1)&7]); a[i] = fma(a[i], b[2], c[i]);
                                  asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i])); }                         // 3 DFMA : 1 MUFU
                if (MODE == 15) { a[i] = fma(a[i], b[0], c[i]); c[i] = fma(c[i], b[1], a[(i+1)&7]); a[i] = fma(a[i], b[2], c[i]);
                                  asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(fa[i]));
                                  fb[i] = fmaf(fb[i], fa[(i+1)&7], fb[(i+2)&7]); fb[i] = fmaf(fb[i], fa[(i+3)&7], fb[(i+5)&7]); }  // + 2 FFMA
                if (MODE == 11) { a[i] = fma(a[i], b[i], c[i]); c[i] = fma(c[i], b[(i+1)&7], a[(i+3)&7]); }
            }
        }
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i] + c[i] + fa[i] + ia[i];
    out[blockIdx.x * blockDim.x + t] = s;
}
template <int MODE> void run(const char* name, double per_iter, const double* in, double* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 1; w <= 16; w *= 2) {     // warps per SMSP
        int threads = 128, blocks = sms * w;
        k<MODE><<<blocks, threads>>>(in, out);
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(in, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double instr = (double)threads * blocks * ITERS * 32.0 * per_iter;
        printf("%-34s warps/SMSP=%2d  %.3f ms  %.3e thread-instr/s  (%.2f /clk/SM @1.93GHz)\n", name, w, ms, instr / (ms * 1e-3), instr / (ms * 1e-3) / 148 / 1.93e9);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 148 * 16 * 128 * 8 * 2);
    double h[512]; for (int i = 0; i < 512; ++i) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount;
    run<0>("DFMA r,r,r distinct", 1, in, out, sms);
    run<1>("DFMA shared multiplier", 1, in, out, sms);
    run<2>("DFMA const operands", 1, in, out, sms);
    run<10>("DMUL r,r", 1, in, out, sms);
    run<11>("DFMA x2 cross-dependent", 2, in, out, sms);
    run<3>("DFMA + FFMA pair (count pairs)", 1, in, out, sms);
    run<4>("F2F.F64.F32 + FADD (count pairs)", 1, in, out, sms);
    run<5>("F2F.F32.F64 + DADD (count pairs)", 1, in, out, sms);
    run<6>("I2FP + IADD (count pairs)", 1, in, out, sms);
    run<7>("FFMA2 r,r,r", 1, in, out, sms);
    run<8>("DFMA + I2FP + IADD (triples)", 1, in, out, sms);
    run<9>("DFMA+F2F+FADD+DFMA (quads)", 1, in, out, sms);
    run<12>("MUFU.SIN", 1, in, out, sms);
    run<13>("DFMA + MUFU (pairs)", 1, in, out, sms);
    run<14>("3 DFMA + MUFU (quads)", 1, in, out, sms);
    run<15>("3 DFMA + MUFU + 2 FFMA (groups)", 1, in, out, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
