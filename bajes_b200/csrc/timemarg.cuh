// Time-shift marginalisation (GWLikelihood.log_like with marg_time_shift=True, bajes/pipe/log_like.py:135-156):
//
//   dh_arr = sum_ifo fft(dh_ifo)            dh_ifo[k] = (4/T) conj(d_k) h_k / S_k on the FULL one-sided axis (zeros
//                                            outside the band), numpy.fft.fft of length Nfr
//   R      = logsumexp_m( Re dh_arr[m] - ln Nfr )          or  ln I0|dh_arr[m]|  with marg_phi_ref
//   logL   = -1/2 (hh + dd) + R - logZ_noise
//
// The FFT is linear, so the detectors are summed BEFORE it: one transform per walker.  This is the one variant
// where a per-bin array has to exist in HBM (W x Nfr complex128); it is produced in sub-batches of walkers by
// `timemarg_integrand_kernel` (same phase / phasor machinery as the fused kernel, one thread per walker), transformed
// by a batched cuFFT Z2Z plan (library call: a plain FFT), and reduced by `timemarg_reduce_kernel`.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include "kernels.cuh"

namespace bj {

// ---- integrand: S[w][band_offset + k] = sum_i amp C_i fix (dr_i + i di_i) Q (c_i - i s_i) -----------------------
template <int NIFO, int MODEL, int SC>
__global__ void __launch_bounds__(kThreads) timemarg_integrand_kernel(
    DeviceConfig cfg, const unsigned char* __restrict__ rec, const ChunkDesc* __restrict__ chunks, long n_bins,
    const double* __restrict__ coef, long stride, long w0, long n_sub, long n_full, long band_offset,
    double2* __restrict__ out) {
    constexpr int RB = RecM<NIFO>::kBytes;
    const long wl_local = (long)blockIdx.y * kThreads + threadIdx.x;
    if (wl_local >= n_sub) return;
    const long w = w0 + wl_local;

    WalkerCoefReg r;
    load_coef(r, coef, stride, w);
    const double amp = coef[C_AMP * stride + w];
    double tau[NIFO], kr[NIFO], ki[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) {
        tau[i] = coef[(C_TAU0 + i) * stride + w];
        const double cr = coef[(C_CR0 + i) * stride + w], ci = coef[(C_CI0 + i) * stride + w];
        // amp * C_i * (table scale / calibrated phasor bias)
        kr[i] = amp * (cr * cfg.dh_fix_re - ci * cfg.dh_fix_im);
        ki[i] = amp * (cr * cfg.dh_fix_im + ci * cfg.dh_fix_re);
    }
    const ChunkDesc chunk = chunks[blockIdx.x];
    double2* row = out + (size_t)wl_local * n_full + band_offset;
    for (int b = 0; b < chunk.len; ++b) {
        const long k = (long)chunk.start + b;
        if (k >= n_bins) break;                                   // zero-weight padding record
        const unsigned char* rp = rec + (size_t)k * RB;
        const double2 h01 = *reinterpret_cast<const double2*>(rp);       // u, L
        const double2 h23 = *reinterpret_cast<const double2*>(rp + 16);  // w5, f
        struct { double x, y, z, w; } h = {h01.x, h01.y, h23.x, h23.y};
        double turns, Q;
        eval_point(r, h.x, h.y, h.z, turns, Q);
        const float Qf = (float)Q;
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            float cq, sq;
            scaled_phasor_f32<SC>(robust_angle(fma(h.w, tau[i], turns)), Qf, cq, sq);
            const float4 d = *reinterpret_cast<const float4*>(rp + 48 + 16 * i);   // dr, di, di, -dr
            const float tr = fmaf(d.y, sq, d.x * cq);                                // (dr + i di)(cq - i sq)
            const float ti = fmaf(d.w, sq, d.z * cq);
            sr += kr[i] * (double)tr - ki[i] * (double)ti;
            si += kr[i] * (double)ti + ki[i] * (double)tr;
        }
        row[k] = make_double2(sr, si);
    }
}

// ---- R_w = logsumexp_m( g(S_w[m]) ) - ln n_full,   g = Re  or  ln I0|.| ------------------------------------------
__global__ void __launch_bounds__(256) timemarg_reduce_kernel(const double2* __restrict__ spec, long n_full,
                                                             int marg_phi_ref, long w0, double* __restrict__ R) {
    __shared__ double sh[256];
    const double2* row = spec + (size_t)blockIdx.x * n_full;
    auto val = [&](long m) {
        const double2 z = row[m];
        return marg_phi_ref ? log_bessel_i0(hypot(z.x, z.y)) : z.x;
    };
    double mx = -INFINITY;
    for (long m = threadIdx.x; m < n_full; m += blockDim.x) mx = fmax(mx, val(m));
    sh[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    mx = sh[0];
    __syncthreads();
    double sum = 0.0;
    for (long m = threadIdx.x; m < n_full; m += blockDim.x) sum += exp(val(m) - mx);
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) R[w0 + blockIdx.x] = mx + log(sh[0]) - log((double)n_full);
}

// ---- logL from R and the Gram-moment <h|h> -----------------------------------------------------------------------
__global__ void timemarg_finish_kernel(DeviceConfig cfg, const double* __restrict__ gram, const double* __restrict__ coef,
                                       long stride, long n_walkers, const double* __restrict__ R,
                                       double* __restrict__ logl) {
    const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    const double amp = coef[C_AMP * stride + w];
    double g[kBasis];
    g[0] = 1.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) g[1 + j] = coef[(C_Q2 + j) * stride + w];
    g[7] = coef[C_Q6L * stride + w];
    double hh = 0.0;
    for (int i = 0; i < cfg.n_ifo; ++i) {
        const double* G = gram + (size_t)i * 3 * kGram;          // one cell, first t-set
        double q = 0.0;
        int idx = 0;
        for (int m = 0; m < kBasis; ++m) {
            double rowv = 0.0;
            for (int n = m; n < kBasis; ++n, ++idx) rowv = fma((n == m ? 1.0 : 2.0) * g[n], G[idx], rowv);
            q = fma(g[m], rowv, q);
        }
        const double cr = coef[(C_CR0 + i) * stride + w], ci = coef[(C_CI0 + i) * stride + w];
        hh += amp * amp * (cr * cr + ci * ci) * q;
    }
    double v = -0.5 * (hh + cfg.dd_sum) + R[w] - cfg.logZ_noise;      // log_like.py:181
    if (coef[C_VALID * stride + w] == 0.0) v = -INFINITY;
    logl[w] = v;
}

// ---- cuFFT, loaded lazily (the main path has no link-time dependency on it) -----------------------------------------
struct CufftApi {
    void* lib = nullptr;
    int (*PlanMany)(int*, int, int*, int*, int, int, int*, int, int, int, int) = nullptr;
    int (*ExecZ2Z)(int, double2*, double2*, int) = nullptr;
    int (*SetStream)(int, cudaStream_t) = nullptr;
    int (*Destroy)(int) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        PlanMany = (decltype(PlanMany))dlsym(lib, "cufftPlanMany");
        ExecZ2Z = (decltype(ExecZ2Z))dlsym(lib, "cufftExecZ2Z");
        SetStream = (decltype(SetStream))dlsym(lib, "cufftSetStream");
        Destroy = (decltype(Destroy))dlsym(lib, "cufftDestroy");
        return PlanMany && ExecZ2Z && SetStream && Destroy;
    }
};
constexpr int kCufftZ2Z = 0x69;      // CUFFT_Z2Z
constexpr int kCufftForward = -1;    // CUFFT_FORWARD: exp(-2 pi i k m / N), the numpy.fft.fft convention

}  // namespace bj
