// Exhaustive error statistics of the FP32 phasor paths used by the fused kernel, against FP64 sincospi.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../bajes_b200/csrc -o sincos_error sincos_error.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "kernels.cuh"
using namespace bj;
// stats per octant: [0] sum amp err (c^ c + s^ s - 1), [1] sum phase err (s^ c - c^ s), [2] sum amp^2, [3] sum ph^2, [4] max abs
template <int SC>
__global__ void k(double* stats, int shift) {
    __shared__ double acc[8][5];
    if (threadIdx.x < 40) ((double*)acc)[threadIdx.x] = 0.0;
    __syncthreads();
    const long n = 1L << (32 - shift);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int32_t b = (int32_t)((uint32_t)(i << shift) + (uint32_t)(0x9E3779B9u >> (32 - shift)) * (shift ? 1u : 0u));
        double cq, sq;
        scaled_phasor<SC>(b, 1.0, 1.0f, cq, sq);
        double s, c;
        sincospi(2.0 * ((double)b * 2.3283064365386963e-10), &s, &c);
        const double ea = cq * c + sq * s - 1.0, ep = sq * c - cq * s;
        const int o = ((uint32_t)b) >> 29;
        atomicAdd(&acc[o][0], ea); atomicAdd(&acc[o][1], ep);
        atomicAdd(&acc[o][2], ea * ea); atomicAdd(&acc[o][3], ep * ep);
        const double m = fmax(fabs(cq - c), fabs(sq - s));
        // max via atomicMax on the bit pattern (non-negative doubles order like integers)
        atomicMax((unsigned long long*)&acc[o][4], (unsigned long long)__double_as_longlong(m));
    }
    __syncthreads();
    if (threadIdx.x < 40) {
        int o = threadIdx.x / 5, j = threadIdx.x % 5;
        if (j < 4) atomicAdd(&stats[o * 5 + j], acc[o][j]);
        else atomicMax((unsigned long long*)&stats[o * 5 + 4], (unsigned long long)__double_as_longlong(acc[o][4]));
    }
}
template <int SC> void run(const char* name, int shift) {
    double* d; cudaMalloc(&d, 40 * 8); cudaMemset(d, 0, 40 * 8);
    k<SC><<<148 * 8, 256>>>(d, shift);
    double h[40]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = (double)(1L << (32 - shift)), no = n / 8;
    double ta = 0, tp = 0, ta2 = 0, tp2 = 0, mx = 0;
    printf("%s (%.0f angles)\n", name, n);
    for (int o = 0; o < 8; ++o) {
        printf("  octant %d: mean amp err %+.3e  mean phase err %+.3e  rms amp %.3e  rms phase %.3e  max abs %.3e\n", o,
               h[o*5]/no, h[o*5+1]/no, sqrt(h[o*5+2]/no), sqrt(h[o*5+3]/no), h[o*5+4]);
        ta += h[o*5]; tp += h[o*5+1]; ta2 += h[o*5+2]; tp2 += h[o*5+3]; mx = fmax(mx, h[o*5+4]);
    }
    printf("  ALL     : mean amp err %+.3e  mean phase err %+.3e  rms amp %.3e  rms phase %.3e  max abs %.3e\n",
           ta/n, tp/n, sqrt(ta2/n), sqrt(tp2/n), mx);
    cudaFree(d);
}
int main() {
    run<BJ_SINCOS_MUFU>("MUFU   ", 4);
    run<BJ_SINCOS_POLY32>("POLY32 ", 4);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
