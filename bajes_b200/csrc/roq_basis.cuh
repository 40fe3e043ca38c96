// ROQ basis CONSTRUCTION on the device: training set, reduced-basis greedy.
//
// The reference ships only the run-time side of ROQ (loader + weights, bajes/pipe/utils/roq.py:808-872); its own
// construction code is commented out (roq.py:16-662, "internal bajes implementation of an ROQ algorithm").  What is built
// here follows that code: training waveforms = h_plus of TaylorF2 on the in-band axis WITHOUT the centre shift (:471-473,
// waveform.py:390), weighted inner product (a, b) = 4 df sum_k w_k conj(a_k) b_k with w = 1/S (:654-662), greedy
// selection by projection error with the projections kept up to date incrementally (rb_greedy :537-612,
// update_projection :524-527).  One repair: the commented code keeps only the REAL part of (a, b) (:662); under a real
// inner product e and i e are orthogonal, the greedy admits both, and the complex DEIM / interpolant that follow see
// linearly dependent columns (singular to rounding: checked with the NumPy restatement).  The Hermitian product of the
// algorithm it cites (Antil et al., arXiv:1210.0577) is used instead.  DEIM (:614-652) is O(M m^2) on an [M x m] matrix and stays on the host
// (bajes_b200/roq_basis.py).
//
// Layout: the training set is ONE matrix R [n_train][n_f] complex128 in HBM (1 GB at 4096 x 16 385), turned in place into
// the residuals R_i = T_i - P(T_i).  One greedy step = one streaming pass: every row takes its coefficient
// c_i = (e, R_i) against the new basis vector e (first sweep over the row), then subtracts c_i e and accumulates the new
// sigma_i^2 = (R_i, R_i) (second sweep: the row is L2-hot) -- 2 reads + 1 write of the matrix per basis vector, HBM-bound:
// algorithmic bytes per step = 3 n_train n_f 16.  One CTA per row, FP64 throughout, fixed-order block reduction (picks are
// argmax decisions: they must not depend on atomics' order).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace bj {

// h_plus (quadratic == 0) or |h_plus|^2 (quadratic == 1) of walker w at frequency k: bank[w][k]
__global__ void wavebank_kernel(const double* __restrict__ freqs, long n_f, const double* __restrict__ coef, long stride,
                                long n_walkers, int quadratic, double2* __restrict__ bank) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long w = blockIdx.y;
    if (k >= n_f || w >= n_walkers) return;
    WalkerCoefReg r;
    load_coef(r, coef, stride, w);
    const double f = freqs[k];
    const double u = cbrt(f), L = log(f);
    const double w5 = 1.0 / (u * u * u * u * u);
    double turns, Q;
    eval_point(r, u, L, w5, turns, Q);
    const double cosi = coef[C_COSI * stride + w];
    const double amp = coef[C_AMP * stride + w] * Q * pow(f, -7.0 / 6.0) * (1.0 + cosi * cosi) * 0.5;
    double2 out;
    if (quadratic) {
        out.x = amp * amp;
        out.y = 0.0;
    } else {
        double c, sn;
        cis_neg_turns(turns, c, sn);
        out.x = amp * c;
        out.y = amp * sn;
    }
    if (coef[C_VALID * stride + w] == 0.0) out.x = out.y = nan("");
    bank[w * n_f + k] = out;
}

constexpr int kRbThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    // fixed-order tree: warp shuffles, then the eight warp sums in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kRbThreads / 32; ++i) t += sh[i];
    return t;
}

// sigma_i^2 = 4 df sum_k w_k |R_ik|^2
__global__ void __launch_bounds__(kRbThreads) rb_norm_kernel(const double2* __restrict__ R, long n_f,
                                                             const double* __restrict__ wgt, double four_df,
                                                             double* __restrict__ sig2) {
    __shared__ double sh[kRbThreads / 32];
    const double2* row = R + (size_t)blockIdx.x * n_f;
    double acc = 0.0;
    for (long k = threadIdx.x; k < n_f; k += kRbThreads) {
        const double2 v = row[k];
        acc = fma(wgt[k], fma(v.x, v.x, v.y * v.y), acc);
    }
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) sig2[blockIdx.x] = four_df * t;
}

// index and value of the largest sigma^2 (first index on ties, NaN rows never win) -> pick[0] = index, val[0] = sigma^2
__global__ void __launch_bounds__(1024) rb_argmax_kernel(const double* __restrict__ sig2, long n, int* __restrict__ pick,
                                                         double* __restrict__ val) {
    __shared__ double sv[1024];
    __shared__ long si[1024];
    double best = -1.0;
    long bi = 0;
    for (long i = threadIdx.x; i < n; i += 1024) {
        const double v = sig2[i];
        if (v > best) { best = v; bi = i; }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v = sv[threadIdx.x + s];
            const long j = si[threadIdx.x + s];
            if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && j < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = j; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { pick[0] = (int)si[0]; val[0] = sv[0]; }
}

// e = R[pick] / sigma
__global__ void rb_take_kernel(const double2* __restrict__ R, long n_f, const int* __restrict__ pick,
                               const double* __restrict__ val, double2* __restrict__ e) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_f) return;
    const double inv = 1.0 / sqrt(val[0]);
    const double2 v = R[(size_t)pick[0] * n_f + k];
    e[k] = make_double2(v.x * inv, v.y * inv);
}

// one greedy step on every row: c_i = 4 df sum_k w_k conj(e_k) R_ik (complex);  R_i -= c_i e;  sigma_i^2 = (R_i, R_i)
__global__ void __launch_bounds__(kRbThreads) rb_project_kernel(double2* __restrict__ R, long n_f,
                                                                const double* __restrict__ wgt, double four_df,
                                                                const double2* __restrict__ e, double* __restrict__ sig2,
                                                                double2* __restrict__ coeff) {
    __shared__ double sh[kRbThreads / 32];
    double2* row = R + (size_t)blockIdx.x * n_f;
    double ar = 0.0, ai = 0.0;
    for (long k = threadIdx.x; k < n_f; k += kRbThreads) {
        const double2 v = row[k], b = e[k];
        const double wk = wgt[k];
        ar = fma(wk, fma(b.x, v.x, b.y * v.y), ar);             // conj(e) R
        ai = fma(wk, fma(b.x, v.y, -b.y * v.x), ai);
    }
    const double cr = four_df * block_sum(ar, sh);
    const double ci = four_df * block_sum(ai, sh);
    double a2 = 0.0;
    for (long k = threadIdx.x; k < n_f; k += kRbThreads) {
        double2 v = row[k];
        const double2 b = e[k];
        v.x -= fma(cr, b.x, -ci * b.y);
        v.y -= fma(cr, b.y, ci * b.x);
        row[k] = v;
        a2 = fma(wgt[k], fma(v.x, v.x, v.y * v.y), a2);
    }
    const double t = block_sum(a2, sh);
    if (threadIdx.x == 0) {
        sig2[blockIdx.x] = four_df * t;
        if (coeff) coeff[blockIdx.x] = make_double2(cr, ci);
    }
}

}  // namespace bj
