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

// Two phases per walker (one warp):
//   1. the waveform at the linear nodes, FP64 (generic Horner + sincospi), into shared memory -- compute only;
//   2. per detector, the four cubic-B-spline coefficient rows are STREAMED against it: the loop body is four independent
//      16-byte loads per node and lane, unrolled, with nothing but FMAs behind them, so that a warp keeps ~8 KB in
//      flight (the first version evaluated the waveform inside this loop: every load waited behind ~100 dependent FP64
//      instructions and the kernel sat at 17 % of the DRAM bandwidth, profiles/r02_roq_wide_before_*.txt).
// Shared memory: kRoqWarps x n_lin complex doubles (26 KB at n_lin = 400).
template <int MODEL>
__global__ void __launch_bounds__(kRoqWarps * 32) roq_kernel(DeviceConfig cfg, RoqDevice tab,
                                                              const double* __restrict__ coef, long stride,
                                                              long n_walkers, double* __restrict__ logl,
                                                              double* __restrict__ sums) {
    extern __shared__ __align__(16) unsigned char roq_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* hs = reinterpret_cast<double2*>(roq_smem) + (size_t)warp * tab.n_lin;
    const long w = (long)blockIdx.x * kRoqWarps + warp;
    if (w >= n_walkers) return;
    const int nifo = cfg.n_ifo;
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

    // ---- phase 1: h(F_L,j) (no centre shift / ramp: waveform.py:390, detector.py:418) and <h|h> ----
    double hh[BJ_MAX_IFO] = {0, 0, 0};
    {
        WalkerCoefReg r;
        load_coef(r, coef, stride, w);
        for (long j = lane; j < tab.n_lin; j += 32) {
            const double4 nd = reinterpret_cast<const double4*>(tab.node)[tab.mask_omega[j]];
            double turns, Q;
            eval_point_m<MODEL>(r, nd.x, nd.y, nd.z, turns, Q);
            double c, sn;
            cis_neg_turns_fast(turns, c, sn);
            const double a = Q * nd.w;
            hs[j] = make_double2(a * c, a * sn);
        }
        // <h|h>: sum_j psi_ij |h(F_Q,j)|^2
        for (long j = lane; j < tab.n_quad; j += 32) {
            const double4 nd = reinterpret_cast<const double4*>(tab.node)[tab.mask_psi[j]];
            double turns, Q;
            eval_point_m<MODEL>(r, nd.x, nd.y, nd.z, turns, Q);
            const double a = Q * nd.w;
            for (int i = 0; i < nifo; ++i) hh[i] = fma(tab.psi[(size_t)i * tab.n_quad + j], a * a, hh[i]);
        }
    }
    __syncwarp();

    // ---- phase 2: <d|h>_i = sum_j omega_ij(tc) h_j, omega = sum_m N_m c_{row0 + m, j} ----
    double zr[BJ_MAX_IFO] = {0, 0, 0}, zi[BJ_MAX_IFO] = {0, 0, 0};
    for (int i = 0; i < nifo; ++i) {
        const double2* p0 = tab.coef + ((size_t)i * tab.n_coef + row0[i]) * tab.n_lin;
        const double2* p1 = p0 + tab.n_lin;
        const double2* p2 = p1 + tab.n_lin;
        const double2* p3 = p2 + tab.n_lin;
        const double n0 = N[i][0], n1 = N[i][1], n2 = N[i][2], n3 = N[i][3];
        double ar = 0.0, ai = 0.0;
#pragma unroll 4
        for (long j = lane; j < tab.n_lin; j += 32) {
            const double2 c0 = __ldg(p0 + j), c1 = __ldg(p1 + j), c2 = __ldg(p2 + j), c3 = __ldg(p3 + j);
            const double2 h = hs[j];
            const double wr = fma(n3, c3.x, fma(n2, c2.x, fma(n1, c1.x, n0 * c0.x)));
            const double wi = fma(n3, c3.y, fma(n2, c2.y, fma(n1, c1.y, n0 * c0.y)));
            ar += wr * h.x - wi * h.y;
            ai += wr * h.y + wi * h.x;
        }
        zr[i] = ar;
        zi[i] = ai;
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

template <int MODEL>
static inline int launch_roq_m(const DeviceConfig& cfg, const RoqDevice& tab, const double* coef, long stride,
                               long n_walkers, double* logl, double* sums, cudaStream_t st, int* info) {
    const unsigned nblk = (unsigned)((n_walkers + kRoqWarps - 1) / kRoqWarps);
    const size_t smem = (size_t)kRoqWarps * tab.n_lin * sizeof(double2);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(roq_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    roq_kernel<MODEL><<<nblk, kRoqWarps * 32, smem, st>>>(cfg, tab, coef, stride, n_walkers, logl, sums);
    info[0] = (int)nblk; info[1] = 1; info[2] = kRoqWarps * 32; info[3] = (int)smem;
    info[4] = (int)tab.n_lin; info[5] = (int)tab.n_quad; info[6] = (int)tab.n_coef; info[7] = 100;
    return (int)cudaGetLastError();
}

static inline int launch_roq(const DeviceConfig& cfg, const RoqDevice& tab, const double* coef, long stride,
                             long n_walkers, double* logl, double* sums, cudaStream_t st, int* info) {
    if (cfg.approx == BJ_APPROX_TAYLORF2_35PN)
        return launch_roq_m<MODEL_35>(cfg, tab, coef, stride, n_walkers, logl, sums, st, info);
    return launch_roq_m<MODEL_GENERIC>(cfg, tab, coef, stride, n_walkers, logl, sums, st, info);
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

// One thread per walker, loop over the bin edges.  NIFO is a template parameter so that the per-detector state stays in
// registers (the first version indexed it with a run-time detector count: it lived in local memory), and the edge / bin
// tables -- read by every thread at the same index -- sit in shared memory when they fit (STAGE).
template <int NIFO, int MODEL, bool STAGE>
__global__ void __launch_bounds__(128) relbin_kernel(DeviceConfig cfg, RelbinDevice tab,
                                                      const double* __restrict__ coef, long stride, long n_walkers,
                                                      double* __restrict__ logl, double* __restrict__ sums) {
    extern __shared__ __align__(16) unsigned char relbin_smem[];
    const long ne = tab.n_edges, nb = ne - 1;
    const double* edge = tab.edge;
    const double2* inv_h0 = tab.inv_h0;
    const double* bin = tab.bin;
    if constexpr (STAGE) {
        double* s_edge = reinterpret_cast<double*>(relbin_smem);
        double2* s_ih = reinterpret_cast<double2*>(s_edge + 6 * ne);
        double* s_bin = reinterpret_cast<double*>(s_ih + (size_t)NIFO * ne);
        for (long k = threadIdx.x; k < 6 * ne; k += blockDim.x) s_edge[k] = tab.edge[k];
        for (long k = threadIdx.x; k < (long)NIFO * ne; k += blockDim.x) s_ih[k] = tab.inv_h0[k];
        for (long k = threadIdx.x; k < (long)NIFO * nb * 8; k += blockDim.x) s_bin[k] = tab.bin[k];
        __syncthreads();
        edge = s_edge;
        inv_h0 = s_ih;
        bin = s_bin;
    }
    const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    WalkerCoefReg r;
    load_coef(r, coef, stride, w);
    const double amp = coef[C_AMP * stride + w];
    const bool valid = coef[C_VALID * stride + w] != 0.0;
    double tau[NIFO], cr[NIFO], ci[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) {
        tau[i] = coef[(C_TAU0 + i) * stride + w];
        cr[i] = coef[(C_CR0 + i) * stride + w];
        ci[i] = coef[(C_CI0 + i) * stride + w];
    }
    double pr[NIFO], pi_[NIFO];   // r(f) = h / h0 at the previous edge
    double dre[NIFO], dim[NIFO], hh[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) { pr[i] = pi_[i] = dre[i] = dim[i] = hh[i] = 0.0; }
#pragma unroll 1
    for (long k = 0; k < ne; ++k) {
        const double* ed = edge + 6 * k;
        const double u = ed[0], L = ed[1], w5 = ed[2], g = ed[3], f = ed[4];
        double turns, Q;
        eval_point_m<MODEL>(r, u, L, w5, turns, Q);
        const double a = amp * Q * g;
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            // h_det = C_i a exp(-2 pi i [turns - f T / 2 + f tau_i])
            double c, sn;
            cis_neg_turns_fast(turns + f * tau[i] - 0.5 * f * tab.seglen, c, sn);
            const double hr = a * (cr[i] * c - ci[i] * sn);
            const double hi = a * (cr[i] * sn + ci[i] * c);
            const double2 ih = inv_h0[(size_t)i * ne + k];
            const double rr = hr * ih.x - hi * ih.y;
            const double ri = hr * ih.y + hi * ih.x;
            if (k > 0) {
                const double* bn = bin + ((size_t)i * nb + (k - 1)) * 8;
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
#pragma unroll
    for (int i = 0; i < NIFO; ++i) {
        if (sums) {
            sums[(w * NIFO + i) * 3 + 0] = dre[i];
            sums[(w * NIFO + i) * 3 + 1] = dim[i];
            sums[(w * NIFO + i) * 3 + 2] = hh[i];
        }
        dh_re += dre[i]; dh_im += dim[i]; hh_tot += hh[i];
    }
    const double R = cfg.marg_phi_ref ? log_bessel_i0(hypot(dh_re, dh_im)) : dh_re;
    double v = R - 0.5 * hh_tot;      // binning.py:269
    if (!valid || !isfinite(v)) v = -INFINITY;
    logl[w] = v;
}

template <int NIFO, int MODEL>
static inline int launch_relbin_n(const DeviceConfig& cfg, const RelbinDevice& tab, const double* coef, long stride,
                                  long n_walkers, double* logl, double* sums, cudaStream_t st, int* info) {
    const unsigned nblk = (unsigned)((n_walkers + 127) / 128);
    const long ne = tab.n_edges, nb = ne - 1;
    const size_t smem = (size_t)(6 * ne + 2 * NIFO * ne + 8 * NIFO * nb) * sizeof(double);
    if (smem <= 100 * 1024) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(relbin_kernel<NIFO, MODEL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        relbin_kernel<NIFO, MODEL, true><<<nblk, 128, smem, st>>>(cfg, tab, coef, stride, n_walkers, logl, sums);
        info[3] = (int)smem;
    } else {
        relbin_kernel<NIFO, MODEL, false><<<nblk, 128, 0, st>>>(cfg, tab, coef, stride, n_walkers, logl, sums);
        info[3] = 0;
    }
    info[0] = (int)nblk; info[1] = 1; info[2] = 128;
    info[4] = (int)tab.n_edges; info[5] = 0; info[6] = 0; info[7] = 200;
    return (int)cudaGetLastError();
}

template <int NIFO>
static inline int launch_relbin_i(const DeviceConfig& cfg, const RelbinDevice& tab, const double* coef, long stride,
                                  long n_walkers, double* logl, double* sums, cudaStream_t st, int* info) {
    if (cfg.approx == BJ_APPROX_TAYLORF2_35PN)
        return launch_relbin_n<NIFO, MODEL_35>(cfg, tab, coef, stride, n_walkers, logl, sums, st, info);
    return launch_relbin_n<NIFO, MODEL_GENERIC>(cfg, tab, coef, stride, n_walkers, logl, sums, st, info);
}

static inline int launch_relbin(const DeviceConfig& cfg, const RelbinDevice& tab, const double* coef, long stride,
                                long n_walkers, double* logl, double* sums, cudaStream_t st, int* info) {
    switch (cfg.n_ifo) {
        case 1: return launch_relbin_i<1>(cfg, tab, coef, stride, n_walkers, logl, sums, st, info);
        case 2: return launch_relbin_i<2>(cfg, tab, coef, stride, n_walkers, logl, sums, st, info);
        default: return launch_relbin_i<3>(cfg, tab, coef, stride, n_walkers, logl, sums, st, info);
    }
}

}  // namespace bj
