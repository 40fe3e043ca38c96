// Device kernels of the batched GW likelihood (sm_100a).
//
//   prologue_kernel    one thread per walker  -> coefficient vector (walker.cuh)
//   fullgrid_kernel    fused waveform + projection + <d|h>,<h|h> partial sums per
//                      (walker tile x frequency chunk); static per-bin records stream through a
//                      shared-memory ring filled by TMA bulk copies; nothing per-bin touches HBM
//   epilogue_kernel    fixed-order sum over frequency chunks, C_i / amplitude scaling, logL
//   waveform_kernel    materialises h_det,i(f) for ONE walker (FP64 sin/cos) -- per-point API
//
// Reference behaviour: bajes/pipe/log_like.py:107-183, bajes/obs/gw/detector.py:394-544.
#pragma once
#include <cuda_runtime.h>
#include "walker.cuh"

namespace bj {

// ----------------------------------------------------------------------------------------------
// per-bin record layout (doubles):  [u, L, w5, f, (dr_i, di_i) x NIFO, s_i x NIFO, pad]
//   u = f^(1/3), L = ln f, w5 = f^(-5/3)
//   dr + i di = (4/T) conj(d_i)/S_i * f^(-7/6) * exp(+i pi f T)      (centre shift folded in)
//   s         = (4/T)/S_i * f^(-7/3)
// ----------------------------------------------------------------------------------------------
template <int NIFO> struct Rec {
    static constexpr int kDoubles = ((4 + 3 * NIFO) + 1) & ~1;   // even => 16-byte aligned records
    static constexpr int kBytes = kDoubles * 8;
};

constexpr int kThreads = 128;        // walkers per CTA
constexpr int kTileBins = 64;        // bins per smem stage
constexpr int kStages = 4;
constexpr double kMagic = 1572864.0; // 1.5 * 2^20: ulp(t + kMagic) == 2^-32 turns for |t| < 2^19
constexpr double kMaxTurns = 450000.0;

enum : int { MODEL_35 = 0, MODEL_GENERIC = 1 };

// ---- mbarrier / TMA bulk-copy helpers (PTX) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- sin/cos of a 32-bit binary angle (turns * 2^32, two's complement) ------------------------
template <int SC>
__device__ __forceinline__ void sincos_turns32(int32_t b, double& s, double& c) {
    if constexpr (SC == BJ_SINCOS_MUFU) {
        const float x = __int2float_rn(b) * 1.4629180792671596e-9f;   // 2 pi / 2^32  -> [-pi, pi]
        float sf, cf;
        __sincosf(x, &sf, &cf);
        s = (double)sf;
        c = (double)cf;
    } else if constexpr (SC == BJ_SINCOS_POLY32) {
        // nearest quarter turn q, remainder in [-pi/4, pi/4]
        const int32_t q = (b + (1 << 29)) >> 30;
        const int32_t r = b - (q << 30);
        const float x = __int2float_rn(r) * 1.4629180792671596e-9f;
        const float x2 = x * x;
        float ps = fmaf(-1.9515295891e-4f, x2, 8.3321608736e-3f);
        ps = fmaf(ps, x2, -1.6666654611e-1f);
        const float sf = fmaf(x * x2, ps, x);
        float pc = fmaf(2.443315711809948e-5f, x2, -1.388731625493765e-3f);
        pc = fmaf(pc, x2, 4.166664568298827e-2f);
        const float cf = fmaf(x2 * x2, pc, fmaf(-0.5f, x2, 1.0f));
        const float s0 = (q & 1) ? cf : sf;
        const float c0 = (q & 1) ? sf : cf;
        // q mod 4: 0 (c,s)  1 (-s,c)  2 (-c,-s)  3 (s,-c)
        const bool neg_s = (q & 2) != 0;
        const bool neg_c = ((q + 1) & 2) != 0;
        s = (double)(neg_s ? -s0 : s0);
        c = (double)(neg_c ? -c0 : c0);
    } else {
        const double r = (double)b * 2.3283064365386963e-10;   // 2^-32 -> turns in [-0.5, 0.5)
        sincospi(2.0 * r, &s, &c);
    }
}

// ---- prologue ---------------------------------------------------------------------------------
__global__ void prologue_kernel(DeviceConfig cfg, ParamMapDev pm, const double* __restrict__ x, int ndim,
                                long n_walkers, double* __restrict__ coef, long stride) {
    const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    double out[C_NUM];
    walker_coefficients(cfg, pm, x + w * ndim, out);
#pragma unroll 1
    for (int j = 0; j < C_NUM; ++j) coef[j * stride + w] = out[j];
}

// ---- the fused full-grid kernel -----------------------------------------------------------------
template <int NIFO, int MODEL, int SC>
__global__ void __launch_bounds__(kThreads) fullgrid_kernel(const double* __restrict__ rec, long n_bins,
                                                              int bins_per_chunk, const double* __restrict__ coef,
                                                              long stride, long n_walkers,
                                                              double* __restrict__ partial) {
    constexpr int RS = Rec<NIFO>::kDoubles;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kTileBins * RS * 8);

    const long w = (long)blockIdx.y * kThreads + threadIdx.x;
    const long wl = w < n_walkers ? w : n_walkers - 1;

    // coefficients -> registers
    constexpr int NA = (MODEL == MODEL_35) ? 8 : 16;
    double a[NA];
    double a10 = 0.0, a12 = 0.0;
#pragma unroll
    for (int j = 0; j < NA; ++j) a[j] = coef[(C_A0 + j) * stride + wl];
    if constexpr (MODEL == MODEL_35) {
        a10 = coef[(C_A0 + 10) * stride + wl];
        a12 = coef[(C_A0 + 12) * stride + wl];
    }
    constexpr int NB = (MODEL == MODEL_35) ? 2 : 7;
    double bq[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) bq[j] = coef[(C_B0 + j) * stride + wl];
    double c8 = 0.0;
    if constexpr (MODEL != MODEL_35) c8 = coef[C_C8 * stride + wl];
    const double k0m = coef[C_K0 * stride + wl] + kMagic;
    double q[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) q[j] = coef[(C_Q2 + j) * stride + wl];
    const double q6l = coef[C_Q6L * stride + wl];
    double tau[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) tau[i] = coef[(C_TAU0 + i) * stride + wl];

    double zr[NIFO], zi[NIFO], hh[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) { zr[i] = 0.0; zi[i] = 0.0; hh[i] = 0.0; }

    const long bin0 = (long)blockIdx.x * bins_per_chunk;
    long remaining = n_bins - bin0;
    const int nb_total = remaining < bins_per_chunk ? (int)remaining : bins_per_chunk;
    const int n_tiles = (nb_total + kTileBins - 1) / kTileBins;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int slot) {
        const int nb = min(kTileBins, nb_total - tile * kTileBins);
        const uint32_t bytes = (uint32_t)nb * RS * 8;
        mbar_expect_tx(&full[slot], bytes);
        tma_bulk_g2s(stages + (size_t)slot * kTileBins * RS, rec + (bin0 + (long)tile * kTileBins) * RS, bytes,
                     &full[slot]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages && s < n_tiles; ++s) issue(s, s);
    }

    for (int t = 0; t < n_tiles; ++t) {
        const int slot = t % kStages;
        mbar_wait(&full[slot], (uint32_t)((t / kStages) & 1));
        const double* tile = stages + (size_t)slot * kTileBins * RS;
        const int nb = min(kTileBins, nb_total - t * kTileBins);
#pragma unroll 2
        for (int b = 0; b < nb; ++b) {
            const double2* r = reinterpret_cast<const double2*>(tile + b * RS);
            const double2 h0 = r[0];   // u, L
            const double2 h1 = r[1];   // w5, f
            const double u = h0.x, L = h0.y, w5 = h1.x, f = h1.y;
            const double u2 = u * u;

            // ---- phase in turns (+ magic) ------------------------------------------------------
            double ph;
            if constexpr (MODEL == MODEL_35) {
                const double u3 = u2 * u;
                double p = fma(a12, u2, a10);          // tidal pair enters at u^7 * u^3
                p = fma(p, u3, a[7]);
                p = fma(p, u, a[6]);
                p = fma(p, u, a[5]);
                p = fma(p, u, a[4]);
                p = fma(p, u, a[3]);
                p = fma(p, u, a[2]);
                p = fma(p, u2, a[0]);
                const double bl = fma(bq[1], u, bq[0]);
                ph = fma(w5, p, fma(L, bl, k0m));
            } else {
                double p = a[15];
#pragma unroll
                for (int j = 14; j >= 0; --j) p = fma(p, u, a[j]);
                double bl = bq[6];
#pragma unroll
                for (int j = 5; j >= 0; --j) bl = fma(bl, u, bq[j]);
                const double u3 = u2 * u;
                bl = fma(L, c8 * u3, bl);
                ph = fma(w5, p, fma(L, bl, k0m));
            }

            // ---- amplitude series ----------------------------------------------------------------
            double qq = fma(q6l, L, q[4]);
            qq = fma(q[5], u, qq);
            qq = fma(qq, u, q[3]);
            qq = fma(qq, u, q[2]);
            qq = fma(qq, u, q[1]);
            qq = fma(qq, u, q[0]);
            const double Q = fma(u2, qq, 1.0);
            const double Q2 = Q * Q;

            // ---- per detector ----------------------------------------------------------------------
#pragma unroll
            for (int i = 0; i < NIFO; ++i) {
                const double tt = fma(f, tau[i], ph);
                double s, c;
                sincos_turns32<SC>(__double2loint(tt), s, c);
                const double dr = tile[b * RS + 4 + 2 * i];
                const double di = tile[b * RS + 5 + 2 * i];
                const double sw = tile[b * RS + 4 + 2 * NIFO + i];
                // (dr + i di) * (c - i s)
                const double pr = fma(dr, c, di * s);
                const double pi_ = fma(di, c, -(dr * s));
                zr[i] = fma(Q, pr, zr[i]);
                zi[i] = fma(Q, pi_, zi[i]);
                hh[i] = fma(Q2, sw, hh[i]);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && t + kStages < n_tiles) {
            fence_proxy_async();
            issue(t + kStages, slot);
        }
    }

    if (w < n_walkers) {
        double* out = partial + (size_t)blockIdx.x * (3 * NIFO) * stride + w;
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            out[(3 * i + 0) * stride] = zr[i];
            out[(3 * i + 1) * stride] = zi[i];
            out[(3 * i + 2) * stride] = hh[i];
        }
    }
}

// log(I0(x)) = log(i0e(x)) + x for x >= 0   (log_like.py:176-177)
__device__ inline double log_bessel_i0(double x) {
    if (x < 500.0) return log(cyl_bessel_i0(x));
    // Hankel asymptotic expansion of i0e
    const double r = 1.0 / (8.0 * x);
    double sum = 1.0, term = 1.0;
#pragma unroll 1
    for (int k = 1; k <= 12; ++k) {
        const double odd = (double)(2 * k - 1);
        term *= odd * odd * r / (double)k;
        sum += term;
    }
    return x + log(sum) - 0.5 * log(kTwoPi * x);
}

// ---- epilogue: fixed-order reduction over chunks + logL ---------------------------------------------
// sums[W][n_ifo][3] (optional) receives Re<d|h>, Im<d|h>, <h|h> per detector.
__global__ void epilogue_kernel(DeviceConfig cfg, const double* __restrict__ partial, int n_chunks,
                                const double* __restrict__ coef, long stride, long n_walkers,
                                double* __restrict__ logl, double* __restrict__ sums) {
    const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    const int nifo = cfg.n_ifo;
    const double amp = coef[C_AMP * stride + w];
    const bool valid = coef[C_VALID * stride + w] != 0.0;
    double dh_re = 0.0, dh_im = 0.0, hh_tot = 0.0;
    for (int i = 0; i < nifo; ++i) {
        double zr = 0.0, zi = 0.0, h = 0.0;
        for (int ch = 0; ch < n_chunks; ++ch) {
            const double* p = partial + ((size_t)ch * (3 * nifo) + 3 * i) * stride + w;
            zr += p[0];
            zi += p[stride];
            h += p[2 * stride];
        }
        const double cr = coef[(C_CR0 + i) * stride + w];
        const double ci = coef[(C_CI0 + i) * stride + w];
        const double re = amp * (cr * zr - ci * zi);
        const double im = amp * (cr * zi + ci * zr);
        const double hi = amp * amp * (cr * cr + ci * ci) * h;
        if (sums) {
            sums[(w * nifo + i) * 3 + 0] = re;
            sums[(w * nifo + i) * 3 + 1] = im;
            sums[(w * nifo + i) * 3 + 2] = hi;
        }
        dh_re += re;
        dh_im += im;
        hh_tot += hi;
    }
    double R;
    if (cfg.marg_phi_ref) R = log_bessel_i0(hypot(dh_re, dh_im));
    else                  R = dh_re;
    // log_like.py:181 with _psd_fact = 0
    double v = -0.5 * (hh_tot + cfg.dd_sum) + R - cfg.logZ_noise;
    if (!valid) v = -INFINITY;
    logl[w] = v;
}

// ---- FP64 generic evaluation at one frequency (used by waveform / ROQ / relbin kernels) ---------
struct WalkerCoefReg {
    double a[16], b[7], c8, k0, q[6], q6l;
};
__device__ __forceinline__ void load_coef(WalkerCoefReg& r, const double* __restrict__ coef, long stride, long w) {
#pragma unroll
    for (int j = 0; j < 16; ++j) r.a[j] = coef[(C_A0 + j) * stride + w];
#pragma unroll
    for (int j = 0; j < 7; ++j) r.b[j] = coef[(C_B0 + j) * stride + w];
    r.c8 = coef[C_C8 * stride + w];
    r.k0 = coef[C_K0 * stride + w];
#pragma unroll
    for (int j = 0; j < 6; ++j) r.q[j] = coef[(C_Q2 + j) * stride + w];
    r.q6l = coef[C_Q6L * stride + w];
}
// returns phase in turns and the amplitude series Q(u)
__device__ __forceinline__ void eval_point(const WalkerCoefReg& r, double u, double L, double w5, double& turns,
                                           double& Q) {
    double p = r.a[15];
#pragma unroll
    for (int j = 14; j >= 0; --j) p = fma(p, u, r.a[j]);
    double bl = r.b[6];
#pragma unroll
    for (int j = 5; j >= 0; --j) bl = fma(bl, u, r.b[j]);
    const double u2 = u * u;
    bl = fma(L, r.c8 * (u2 * u), bl);
    turns = fma(w5, p, fma(L, bl, r.k0));
    double qq = fma(r.q6l, L, r.q[4]);
    qq = fma(r.q[5], u, qq);
    qq = fma(qq, u, r.q[3]);
    qq = fma(qq, u, r.q[2]);
    qq = fma(qq, u, r.q[1]);
    qq = fma(qq, u, r.q[0]);
    Q = fma(u2, qq, 1.0);
}
// e^{-2 pi i turns}
__device__ __forceinline__ void cis_neg_turns(double turns, double& c, double& s_neg) {
    const double r = turns - rint(turns);
    double s;
    sincospi(2.0 * r, &s, &c);
    s_neg = -s;
}

// ---- waveform of ONE walker (walker index 0 of coef) on every detector ---------------------------
// out[i][k] = projected strain (re, im) -- Waveform.compute_hphc + Detector.project_fdwave
__global__ void waveform_kernel(DeviceConfig cfg, const double* __restrict__ freqs, long n_bins,
                                const double* __restrict__ coef, long stride, double2* __restrict__ out) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_bins) return;
    WalkerCoefReg r;
    load_coef(r, coef, stride, 0);
    const double f = freqs[k];
    const double u = cbrt(f), L = log(f);
    const double w5 = 1.0 / (u * u * u * u * u);
    double turns, Q;
    eval_point(r, u, L, w5, turns, Q);
    const double amp = coef[C_AMP * stride] * Q * pow(f, -7.0 / 6.0);
    const bool valid = coef[C_VALID * stride] != 0.0;
    for (int i = 0; i < cfg.n_ifo; ++i) {
        const double tau = coef[(C_TAU0 + i) * stride];
        // centre shift exp(+i pi f T) and detector ramp exp(-2 pi i f tau)
        const double tt = turns + f * tau - 0.5 * f * cfg.seglen;
        double c, sn;
        cis_neg_turns(tt, c, sn);
        const double cr = coef[(C_CR0 + i) * stride], ci = coef[(C_CI0 + i) * stride];
        double2 h;
        h.x = amp * (cr * c - ci * sn);
        h.y = amp * (cr * sn + ci * c);
        if (!valid) { h.x = nan(""); h.y = nan(""); }
        out[(size_t)i * n_bins + k] = h;
    }
}

// ---- h_plus, h_cross of ONE walker on an arbitrary axis ---------------------------------------------
__global__ void polarizations_kernel(double seglen, int centre_shift, const double* __restrict__ freqs, long n,
                                     const double* __restrict__ coef, long stride, double2* __restrict__ hp,
                                     double2* __restrict__ hc) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    WalkerCoefReg r;
    load_coef(r, coef, stride, 0);
    const double f = freqs[k];
    const double u = cbrt(f), L = log(f);
    const double w5 = 1.0 / (u * u * u * u * u);
    double turns, Q;
    eval_point(r, u, L, w5, turns, Q);
    if (centre_shift) turns -= 0.5 * f * seglen;
    const double amp = coef[C_AMP * stride] * Q * pow(f, -7.0 / 6.0);
    const double cosi = coef[C_COSI * stride];
    double c, sn;
    cis_neg_turns(turns, c, sn);
    const double ip = (1.0 + cosi * cosi) * 0.5;
    const bool valid = coef[C_VALID * stride] != 0.0;
    double2 p, x;
    p.x = amp * ip * c;
    p.y = amp * ip * sn;
    // hc = h * (-i) * cos(iota)
    x.x = amp * cosi * sn;
    x.y = -amp * cosi * c;
    if (!valid) { p.x = p.y = x.x = x.y = nan(""); }
    hp[k] = p;
    hc[k] = x;
}

}  // namespace bj
