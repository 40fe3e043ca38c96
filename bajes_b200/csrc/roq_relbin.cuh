// ROQ and relative-binning likelihood kernels (FP64 throughout, no tensor cores: neither is a dense
// contraction per walker -- the ROQ weights are a walker-dependent spline gather).
//
// ROQ     : Detector.compute_inner_products, roq branch (bajes/obs/gw/detector.py:534-537) with the
//           weights of bajes/pipe/utils/roq.py:808-872; one WARP per walker, lanes stride the nodes so
//           the four cubic-B-spline coefficient rows are read with coalesced 16-byte loads.
// relbin  : GWBinningLikelihood.log_like (bajes/pipe/utils/binning.py:248-306, with the repairs listed
//           in DESIGN.md); one THREAD per walker, bin-edge tables are warp-uniform loads.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <vector>

#include "kernels.cuh"

namespace bj {

// ================================================================================================
// ROQ
// ================================================================================================
struct RoqDevice {
    int n_ifo = 0;
    long n_join = 0, n_lin = 0, n_quad = 0, n_knots = 0, n_coef = 0;
    double tc_min = 0, tc_max = 0;
    double* node = nullptr;        // [n_join][4] : u, L, w5, g = f^(-7/6)
    int* mask_omega = nullptr;     // [n_lin]
    int* mask_psi = nullptr;       // [n_quad]
    double* knots = nullptr;       // [n_ifo][n_knots]
    double2* coef = nullptr;       // [n_ifo][n_coef][n_lin]
    double* psi = nullptr;         // [n_ifo][n_quad]

    cudaError_t upload(int nifo, const bj_roq_tables* t) {
        n_ifo = nifo;
        n_join = t->n_join; n_lin = t->n_lin; n_quad = t->n_quad; n_knots = t->n_knots; n_coef = t->n_coef;
        tc_min = t->tc_min; tc_max = t->tc_max;
        std::vector<double> nd((size_t)n_join * 4);
        for (long j = 0; j < n_join; ++j) {
            const double f = t->freqs_join[j];
            const double u = std::cbrt(f);
            nd[4 * j + 0] = u;
            nd[4 * j + 1] = std::log(f);
            nd[4 * j + 2] = 1.0 / (u * u * u * u * u);
            nd[4 * j + 3] = std::pow(f, -7.0 / 6.0);
        }
        cudaError_t e;
#define BJ_UP(dst, src, bytes)                                                      \
    if ((e = cudaMalloc((void**)&(dst), (bytes))) != cudaSuccess) return e;         \
    if ((e = cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        BJ_UP(node, nd.data(), nd.size() * sizeof(double));
        BJ_UP(mask_omega, t->mask_omega, (size_t)n_lin * sizeof(int));
        BJ_UP(mask_psi, t->mask_psi, (size_t)n_quad * sizeof(int));
        if ((e = cudaMalloc((void**)&knots, (size_t)nifo * n_knots * sizeof(double))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&coef, (size_t)nifo * n_coef * n_lin * sizeof(double2))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&psi, (size_t)nifo * n_quad * sizeof(double))) != cudaSuccess) return e;
        for (int i = 0; i < nifo; ++i) {
            if ((e = cudaMemcpy(knots + (size_t)i * n_knots, t->knots[i], (size_t)n_knots * sizeof(double),
                                cudaMemcpyHostToDevice)) != cudaSuccess) return e;
            if ((e = cudaMemcpy(coef + (size_t)i * n_coef * n_lin, t->coef_ri[i],
                                (size_t)n_coef * n_lin * sizeof(double2), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
            if ((e = cudaMemcpy(psi + (size_t)i * n_quad, t->psi_weights[i], (size_t)n_quad * sizeof(double),
                                cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        }
#undef BJ_UP
        return cudaSuccess;
    }
    void release() {
        cudaFree(node); cudaFree(mask_omega); cudaFree(mask_psi); cudaFree(knots); cudaFree(coef); cudaFree(psi);
        node = nullptr; mask_omega = nullptr; mask_psi = nullptr; knots = nullptr; coef = nullptr; psi = nullptr;
    }
};

// cubic B-spline: interval l with t[l] <= x < t[l+1] (3 <= l <= n_coef-1) and the four non-zero basis
// values N_{l-3..l}(x)  (Cox - de Boor, the recurrence scipy.interpolate.BSpline evaluates)
__device__ inline int bspline_basis3(const double* __restrict__ t, int n_coef, double x, double N[4]) {
    int lo = 3, hi = n_coef;      // invariant t[lo] <= x < t[hi]  (x == t[n_coef] maps to the last interval)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= t[mid]) lo = mid; else hi = mid;
    }
    const int l = lo;
    double left[4], right[4];
    N[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= 3; ++j) {
        left[j] = x - t[l + 1 - j];
        right[j] = t[l + j] - x;
        double saved = 0.0;
#pragma unroll
        for (int r = 0; r < j; ++r) {
            const double temp = N[r] / (right[r + 1] + left[j - r]);
            N[r] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        N[j] = saved;
    }
    return l;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kRoqWarps = 4;

__global__ void __launch_bounds__(kRoqWarps * 32) roq_kernel(DeviceConfig cfg, RoqDevice tab,
                                                              const double* __restrict__ coef, long stride,
                                                              long n_walkers, double* __restrict__ logl,
                                                              double* __restrict__ sums) {
    const long w = (long)blockIdx.x * kRoqWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n_walkers) return;
    const int nifo = cfg.n_ifo;
    WalkerCoefReg r;
    load_coef(r, coef, stride, w);
    const double amp = coef[C_AMP * stride + w];
    bool valid = coef[C_VALID * stride + w] != 0.0;

    // per detector: spline interval and basis at tc = T/2 + tau_i   (detector.py:536)
    int row0[BJ_MAX_IFO];
    double N[BJ_MAX_IFO][4];
    for (int i = 0; i < nifo; ++i) {
        const double tc = cfg.seglen / 2.0 + coef[(C_TAU0 + i) * stride + w];
        if (!(tc >= tab.tc_min && tc <= tab.tc_max)) {   // interp1d fill_value=-inf (roq.py:866)
            valid = false;
            row0[i] = 0;
            N[i][0] = N[i][1] = N[i][2] = N[i][3] = 0.0;
        } else {
            const int l = bspline_basis3(tab.knots + (size_t)i * tab.n_knots, (int)tab.n_coef, tc, N[i]);
            row0[i] = l - 3;
        }
    }

    // <d|h>: sum_j omega_ij(tc) h(F_L,j)     (h without centre shift / ramp: waveform.py:390, detector.py:418)
    double zr[BJ_MAX_IFO] = {0, 0, 0}, zi[BJ_MAX_IFO] = {0, 0, 0};
    for (long j = lane; j < tab.n_lin; j += 32) {
        const int idx = tab.mask_omega[j];
        const double4 nd = reinterpret_cast<const double4*>(tab.node)[idx];
        double turns, Q;
        eval_point(r, nd.x, nd.y, nd.z, turns, Q);
        double c, sn;
        cis_neg_turns(turns, c, sn);
        const double a = Q * nd.w;
        const double hr = a * c, hi = a * sn;
        for (int i = 0; i < nifo; ++i) {
            const double2* cp = tab.coef + ((size_t)i * tab.n_coef + row0[i]) * tab.n_lin + j;
            double wr = 0.0, wi = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const double2 cc = cp[(size_t)m * tab.n_lin];
                wr = fma(N[i][m], cc.x, wr);
                wi = fma(N[i][m], cc.y, wi);
            }
            zr[i] += wr * hr - wi * hi;
            zi[i] += wr * hi + wi * hr;
        }
    }
    // <h|h>: sum_j psi_ij |h(F_Q,j)|^2
    double hh[BJ_MAX_IFO] = {0, 0, 0};
    for (long j = lane; j < tab.n_quad; j += 32) {
        const int idx = tab.mask_psi[j];
        const double4 nd = reinterpret_cast<const double4*>(tab.node)[idx];
        double turns, Q;
        eval_point(r, nd.x, nd.y, nd.z, turns, Q);
        const double a = Q * nd.w;
        for (int i = 0; i < nifo; ++i) hh[i] = fma(tab.psi[(size_t)i * tab.n_quad + j], a * a, hh[i]);
    }

    double dh_re = 0.0, dh_im = 0.0, hh_tot = 0.0;
    for (int i = 0; i < nifo; ++i) {
        const double sr = warp_sum(zr[i]), si = warp_sum(zi[i]), sh = warp_sum(hh[i]);
        const double cr = coef[(C_CR0 + i) * stride + w], ci = coef[(C_CI0 + i) * stride + w];
        const double re = amp * (cr * sr - ci * si);
        const double im = amp * (cr * si + ci * sr);
        const double h = amp * amp * (cr * cr + ci * ci) * sh;
        if (sums && lane == 0) {
            sums[(w * nifo + i) * 3 + 0] = re;
            sums[(w * nifo + i) * 3 + 1] = im;
            sums[(w * nifo + i) * 3 + 2] = h;
        }
        dh_re += re; dh_im += im; hh_tot += h;
    }
    if (lane == 0) {
        const double R = cfg.marg_phi_ref ? log_bessel_i0(hypot(dh_re, dh_im)) : dh_re;
        double v = -0.5 * (hh_tot + cfg.dd_sum) + R - cfg.logZ_noise;
        if (!valid) v = -INFINITY;
        logl[w] = v;
    }
}

static inline int launch_roq(const DeviceConfig& cfg, const RoqDevice& tab, const double* coef, long stride,
                             long n_walkers, double* logl, double* sums, cudaStream_t st, int* info) {
    const unsigned nblk = (unsigned)((n_walkers + kRoqWarps - 1) / kRoqWarps);
    roq_kernel<<<nblk, kRoqWarps * 32, 0, st>>>(cfg, tab, coef, stride, n_walkers, logl, sums);
    info[0] = (int)nblk; info[1] = 1; info[2] = kRoqWarps * 32; info[3] = 0;
    info[4] = (int)tab.n_lin; info[5] = (int)tab.n_quad; info[6] = (int)tab.n_coef; info[7] = 100;
    return (int)cudaGetLastError();
}

// ================================================================================================
// relative binning
// ================================================================================================
struct RelbinDevice {
    int n_ifo = 0;
    long n_edges = 0;
    double seglen = 0;
    double* edge = nullptr;     // [n_edges][6]: u, L, w5, g, f, pad
    double2* inv_h0 = nullptr;  // [n_ifo][n_edges]   1 / h0 at the edges
    double* bin = nullptr;      // [n_ifo][n_bins][8]: A0re, A0im, A1re, A1im, B0, B1, 1/df, pad

    cudaError_t upload(int nifo, double T, long ne, const double* fbin, const double* const* h0_ri,
                       const double* sdat_ri) {
        n_ifo = nifo; n_edges = ne; seglen = T;
        const long nb = ne - 1;
        std::vector<double> ed((size_t)ne * 6, 0.0);
        for (long k = 0; k < ne; ++k) {
            const double f = fbin[k];
            const double u = std::cbrt(f);
            ed[6 * k + 0] = u;
            ed[6 * k + 1] = std::log(f);
            ed[6 * k + 2] = 1.0 / (u * u * u * u * u);
            ed[6 * k + 3] = std::pow(f, -7.0 / 6.0);
            ed[6 * k + 4] = f;
        }
        std::vector<double> ih((size_t)nifo * ne * 2);
        for (int i = 0; i < nifo; ++i)
            for (long k = 0; k < ne; ++k) {
                const double a = h0_ri[i][2 * k], b = h0_ri[i][2 * k + 1];
                const double d = a * a + b * b;
                ih[((size_t)i * ne + k) * 2 + 0] = a / d;
                ih[((size_t)i * ne + k) * 2 + 1] = -b / d;
            }
        // sdat_ri: [4][n_ifo][n_bins][2]
        std::vector<double> bn((size_t)nifo * nb * 8, 0.0);
        auto S = [&](int which, int i, long k, int c) { return sdat_ri[((((size_t)which * nifo + i) * nb) + k) * 2 + c]; };
        for (int i = 0; i < nifo; ++i)
            for (long k = 0; k < nb; ++k) {
                double* o = &bn[((size_t)i * nb + k) * 8];
                o[0] = S(0, i, k, 0); o[1] = S(0, i, k, 1);
                o[2] = S(1, i, k, 0); o[3] = S(1, i, k, 1);
                o[4] = S(2, i, k, 0); o[5] = S(3, i, k, 0);
                o[6] = 1.0 / (fbin[k + 1] - fbin[k]);
            }
        cudaError_t e;
        if ((e = cudaMalloc((void**)&edge, ed.size() * sizeof(double))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(edge, ed.data(), ed.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&inv_h0, ih.size() * sizeof(double))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(inv_h0, ih.data(), ih.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&bin, bn.size() * sizeof(double))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(bin, bn.data(), bn.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    void release() {
        cudaFree(edge); cudaFree(inv_h0); cudaFree(bin);
        edge = nullptr; inv_h0 = nullptr; bin = nullptr;
    }
};

__global__ void __launch_bounds__(128) relbin_kernel(DeviceConfig cfg, RelbinDevice tab,
                                                      const double* __restrict__ coef, long stride, long n_walkers,
                                                      double* __restrict__ logl, double* __restrict__ sums) {
    const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    const int nifo = cfg.n_ifo;
    const long ne = tab.n_edges, nb = ne - 1;
    WalkerCoefReg r;
    load_coef(r, coef, stride, w);
    const double amp = coef[C_AMP * stride + w];
    const bool valid = coef[C_VALID * stride + w] != 0.0;
    double tau[BJ_MAX_IFO], cr[BJ_MAX_IFO], ci[BJ_MAX_IFO];
    for (int i = 0; i < nifo; ++i) {
        tau[i] = coef[(C_TAU0 + i) * stride + w];
        cr[i] = coef[(C_CR0 + i) * stride + w];
        ci[i] = coef[(C_CI0 + i) * stride + w];
    }
    double pr[BJ_MAX_IFO], pi_[BJ_MAX_IFO];   // r(f) = h / h0 at the previous edge
    double dre[BJ_MAX_IFO] = {0, 0, 0}, dim[BJ_MAX_IFO] = {0, 0, 0}, hh[BJ_MAX_IFO] = {0, 0, 0};
    for (long k = 0; k < ne; ++k) {
        const double* ed = tab.edge + 6 * k;
        const double u = ed[0], L = ed[1], w5 = ed[2], g = ed[3], f = ed[4];
        double turns, Q;
        eval_point(r, u, L, w5, turns, Q);
        const double a = amp * Q * g;
        for (int i = 0; i < nifo; ++i) {
            // h_det = C_i a exp(-2 pi i [turns - f T / 2 + f tau_i])
            double c, sn;
            cis_neg_turns(turns + f * tau[i] - 0.5 * f * tab.seglen, c, sn);
            const double hr = a * (cr[i] * c - ci[i] * sn);
            const double hi = a * (cr[i] * sn + ci[i] * c);
            const double2 ih = tab.inv_h0[(size_t)i * ne + k];
            const double rr = hr * ih.x - hi * ih.y;
            const double ri = hr * ih.y + hi * ih.x;
            if (k > 0) {
                const double* bn = tab.bin + ((size_t)i * nb + (k - 1)) * 8;
                // binning.py:301-304
                const double r0r = 0.5 * (pr[i] + rr), r0i = 0.5 * (pi_[i] + ri);
                const double r1r = (rr - pr[i]) * bn[6], r1i = (ri - pi_[i]) * bn[6];
                // binning.py:281-282 (per-bin terms; summed)
                hh[i] += bn[4] * (r0r * r0r + r0i * r0i) + bn[5] * 2.0 * (r0r * r1r + r0i * r1i);
                // A0 conj(r0) + A1 conj(r1)
                dre[i] += (bn[0] * r0r + bn[1] * r0i) + (bn[2] * r1r + bn[3] * r1i);
                dim[i] += (bn[1] * r0r - bn[0] * r0i) + (bn[3] * r1r - bn[2] * r1i);
            }
            pr[i] = rr;
            pi_[i] = ri;
        }
    }
    double dh_re = 0.0, dh_im = 0.0, hh_tot = 0.0;
    for (int i = 0; i < nifo; ++i) {
        if (sums) {
            sums[(w * nifo + i) * 3 + 0] = dre[i];
            sums[(w * nifo + i) * 3 + 1] = dim[i];
            sums[(w * nifo + i) * 3 + 2] = hh[i];
        }
        dh_re += dre[i]; dh_im += dim[i]; hh_tot += hh[i];
    }
    const double R = cfg.marg_phi_ref ? log_bessel_i0(hypot(dh_re, dh_im)) : dh_re;
    double v = R - 0.5 * hh_tot;      // binning.py:269
    if (!valid || !isfinite(v)) v = -INFINITY;
    logl[w] = v;
}

static inline int launch_relbin(const DeviceConfig& cfg, const RelbinDevice& tab, const double* coef, long stride,
                                long n_walkers, double* logl, double* sums, cudaStream_t st, int* info) {
    const unsigned nblk = (unsigned)((n_walkers + 127) / 128);
    relbin_kernel<<<nblk, 128, 0, st>>>(cfg, tab, coef, stride, n_walkers, logl, sums);
    info[0] = (int)nblk; info[1] = 1; info[2] = 128; info[3] = 0;
    info[4] = (int)tab.n_edges; info[5] = 0; info[6] = 0; info[7] = 200;
    return (int)cudaGetLastError();
}

}  // namespace bj
