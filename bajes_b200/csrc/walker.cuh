// Per-walker prologue: turns one parameter point into the coefficient vector the bin kernels
// consume.  All FP64.  Formulas restate bajes/obs/gw/approx/taylorf2.py (phase :158-297, tides
// :9-119, QM :121-156, amplitude :300-345), bajes/obs/gw/waveform.py:334-377 (parameter
// derivations) and bajes/obs/gw/detector.py:246-357 (GMST, antenna pattern, geocentric delay).
//
// Algebraic form (SURVEY.md section 8a "closed form"), with v = c*u, c = (pi M gt)^(1/3),
// u = f^(1/3), L = ln f, w5 = f^(-5/3):
//   Phi/(2 pi) = w5 * sum_j A_j u^j  +  L * sum_m B_m u^m  +  L^2 * C8 * u^3  +  K0
//   |h|        = amp * f^(-7/6) * ( 1 + u^2 ( q2 + q3 u + q4 u^2 + q5 u^3 + (q6 + q6L L) u^4 + q7 u^5 ) )
//   h_det,i    = C_i * |h| * exp(-i [Phi - pi f T + 2 pi f tau_i])
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/bajes_b200.h"

namespace bj {

// literals of taylorf2.py:6-7,379,266-267
__device__ constexpr double kGT = 4.92549094830932e-6;
__device__ constexpr double kEulerGamma = 0.57721566490153286060;
__device__ constexpr double kPre = 3.6686934875530996e-19;
__device__ constexpr double kLog2 = 0.69314718055994528623;
__device__ constexpr double kLog3 = 1.0986122886681097821;
__device__ constexpr double kPi = 3.141592653589793238462643383279502884;
__device__ constexpr double kTwoPi = 6.283185307179586476925286766559005768;
__device__ constexpr double kCLight = 299792458.0;

// coefficient slots (SoA in global memory: coef[slot * stride + walker])
enum : int {
    C_A0 = 0,        // A_0 .. A_15
    C_B0 = 16,       // B_0 .. B_6
    C_C8 = 23,
    C_K0 = 24,       // phi_ref / 2 pi
    C_Q2 = 25,       // q2 .. q7
    C_Q6L = 31,
    C_TAU0 = 32,     // tau_i = geocentric delay + time_shift   [BJ_MAX_IFO]
    C_AMP = 35,
    C_CR0 = 36,      // Re C_i   [BJ_MAX_IFO]
    C_CI0 = 39,      // Im C_i   [BJ_MAX_IFO]
    C_VALID = 42,    // 0 = return -inf (log_like.py:121-129), 1 = evaluate, 2 = evaluate with the wide-phase path
    C_FP0 = 43,      // F+_i     [BJ_MAX_IFO]   (diagnostics / polarisation dump)
    C_FC0 = 46,      // Fx_i
    C_COSI = 49,
    // detector-to-detector rotations of the fused kernel (uniform grids; kernels.cuh: tile_step ROT), detector r + 1
    // against detector 0, d = tau_{r+1} - tau_0, eps = 2 pi df d:  C_ROT0 + 10 r + {0..3} = cos(a eps) - 1,
    // {4..7} = sin(a eps) for a = 3.5, 4.5, 11.5, 12.5 bins; {8, 9} = cos, -sin of 32 eps (advance per 32-bin step)
    C_ROT0 = 50,
    // 1 = the walker's phase slope exceeds the design bound of the compressed frequency axis (kernels.cuh: steep_kernel):
    // it is evaluated on the full grid instead
    C_STEEP = C_ROT0 + 10 * (BJ_MAX_IFO - 1),
    C_NUM
};
constexpr int kRotSlots = 10;

struct DeviceConfig {
    int approx;
    int n_ifo;
    int marg_phi_ref;
    double seglen;
    double dd_sum;
    double logZ_noise;
    double dh_fix_re, dh_fix_im;   // <d|h> correction: table scale / calibrated phasor bias (1 + 0i in fp64 mode)
    double f_lo, f_hi;             // in-band frequency range (0, 0 when the context has no full grid)
    double grid_df;                // bin spacing of a uniform in-band grid, else 0
    bj_detector det[BJ_MAX_IFO];
};

struct ParamMapDev {
    int column[BJ_NUM_PARAMS];
    double value[BJ_NUM_PARAMS];
};

// calibration envelopes / PSD weights (detector.py:17-23,517-531): static dimensions + per-call parameter map
struct ExtrasDims {
    int nspcal;                        // 0 = off
    int nweights;                      // 0 = off
    int len_weights[BJ_MAX_WEIGHTS];   // bins per band
};
typedef bj_extra_map ExtraMapDev;

// extras coefficient block (SoA, coef_ex[slot * stride + walker]):
//   [(i * nspcal + j) * 2 + {0,1}]                      Re / Im of (1 + amp_j) exp(i phi_j), detector i, node j
//   [2 n_ifo nspcal + i * nweights + b]                 1 / w_b of detector i
//   [2 n_ifo nspcal + n_ifo nweights]                   sum_i sum_b len_b log w_b   ("_psd_fact", log_like.py:181)
__host__ __device__ inline int ex_cal_slot(const ExtrasDims& d, int i, int j, int c) { return (i * d.nspcal + j) * 2 + c; }
__host__ __device__ inline int ex_invw_slot(const ExtrasDims& d, int n_ifo, int i, int b) {
    return 2 * n_ifo * d.nspcal + i * d.nweights + b;
}
__host__ __device__ inline int ex_psdfact_slot(const ExtrasDims& d, int n_ifo) {
    return 2 * n_ifo * d.nspcal + n_ifo * d.nweights;
}
__host__ __device__ inline int ex_num_slots(const ExtrasDims& d, int n_ifo) { return ex_psdfact_slot(d, n_ifo) + 1; }

// returns false when a weight is not positive (log w is NaN in the reference)
__device__ inline bool walker_extras(const DeviceConfig& cfg, const ExtrasDims& d, const ExtraMapDev& em,
                                     const double* __restrict__ x, double* __restrict__ coef_ex, long stride, long w) {
    bool ok = true;
    for (int i = 0; i < cfg.n_ifo; ++i) {
        for (int j = 0; j < d.nspcal; ++j) {
            const int ca = em.spcal_amp_col[i][j], cp = em.spcal_phi_col[i][j];
            const double amp = ca >= 0 ? x[ca] : em.spcal_amp_val[i][j];
            const double phi = cp >= 0 ? x[cp] : em.spcal_phi_val[i][j];
            double sn, cs;
            sincos(phi, &sn, &cs);
            coef_ex[ex_cal_slot(d, i, j, 0) * stride + w] = (1.0 + amp) * cs;     // detector.py:17-20
            coef_ex[ex_cal_slot(d, i, j, 1) * stride + w] = (1.0 + amp) * sn;
            ok = ok && isfinite(amp) && isfinite(phi);
        }
    }
    double fact = 0.0;
    for (int i = 0; i < cfg.n_ifo; ++i) {
        for (int b = 0; b < d.nweights; ++b) {
            const int cw = em.weight_col[i][b];
            const double wt = cw >= 0 ? x[cw] : em.weight_val[i][b];
            coef_ex[ex_invw_slot(d, cfg.n_ifo, i, b) * stride + w] = 1.0 / wt;
            fact += (double)d.len_weights[b] * log(wt);                            // detector.py:528
            ok = ok && (wt > 0.0) && isfinite(wt);
        }
    }
    coef_ex[ex_psdfact_slot(d, cfg.n_ifo) * stride + w] = fact;
    return ok;
}

__device__ __forceinline__ bool has(const ParamMapDev& m, int slot) { return m.column[slot] != BJ_COL_ABSENT; }
__device__ __forceinline__ double get(const ParamMapDev& m, const double* __restrict__ x, int slot) {
    int c = m.column[slot];
    return c >= 0 ? x[c] : m.value[slot];
}

// python-style modulo (result has the sign of the divisor), detector.py:252
__device__ __forceinline__ double pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

__device__ __forceinline__ double gmst_estimate(const bj_detector& d, double t) {
    double dphase = (t - d.reference_time) / d.sday * kTwoPi;
    return pymod(d.gmst_reference + dphase, kTwoPi);
}

// detector.py:318-357
__device__ __forceinline__ double geocentric_delay(const bj_detector& d, double ra, double dec, double t) {
    double ang = gmst_estimate(d, t) - ra;
    double sa, ca, sd, cd;
    sincos(ang, &sa, &ca);
    sincos(dec, &sd, &cd);
    double e0 = cd * ca, e1 = cd * -sa, e2 = sd;
    double s = (0.0 - d.location[0]) * e0 + (0.0 - d.location[1]) * e1 + (0.0 - d.location[2]) * e2;
    return s / kCLight;
}

// detector.py:268-316
__device__ __forceinline__ void antenna_pattern(const bj_detector& d, double ra, double dec, double psi, double t,
                                                double& fplus, double& fcross) {
    t += geocentric_delay(d, ra, dec, t);
    double gha = gmst_estimate(d, t) - ra;
    double sg, cg, sd, cd, sp, cp;
    sincos(gha, &sg, &cg);
    sincos(dec, &sd, &cd);
    sincos(psi, &sp, &cp);
    double x[3] = {-cp * sg - sp * cg * sd, -cp * cg + sp * sg * sd, sp * cd};
    double y[3] = {sp * sg - cp * cg * sd, sp * cg + cp * sg * sd, cp * cd};
    double dx[3], dy[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        dx[r] = d.response[3 * r] * x[0] + d.response[3 * r + 1] * x[1] + d.response[3 * r + 2] * x[2];
        dy[r] = d.response[3 * r] * y[0] + d.response[3 * r + 1] * y[1] + d.response[3 * r + 2] * y[2];
    }
    fplus = (x[0] * dx[0] - y[0] * dy[0]) + (x[1] * dx[1] - y[1] * dy[1]) + (x[2] * dx[2] - y[2] * dy[2]);
    fcross = (x[0] * dy[0] + y[0] * dx[0]) + (x[1] * dy[1] + y[1] * dx[1]) + (x[2] * dy[2] + y[2] * dx[2]);
}

// utils/__init__.py:127-160 with m1 = xa, m2 = xb (mass fractions)
__device__ __forceinline__ void tidal_combinations(double xa, double xb, double l1, double l2, double& lt, double& dl) {
    double mt = xa + xb;
    double a4 = xa * xa * xa * xa, b4 = xb * xb * xb * xb;
    double mt4 = mt * mt * mt * mt;
    lt = (16.0 / 13.0) * ((xa + 12.0 * xb) * a4 * l1 + (xb + 12.0 * xa) * b4 * l2) / (mt4 * mt);
    double q = xa / xb;
    double eta = q / ((1.0 + q) * (1.0 + q));
    double X = sqrt(1.0 - 4.0 * eta);
    dl = (1690.0 * eta / 1319.0 - 4843.0 / 1319.0) * (a4 * l1 - b4 * l2) / mt4 +
         (6162.0 * X / 1319.0) * (a4 * l1 + b4 * l2) / mt4;
}

// utils/__init__.py:22-27
__device__ __forceinline__ double quadrupole_yy(double lam) {
    if (lam <= 0.0) return 1.0;
    double ll = log(lam);
    double ll2 = ll * ll;
    return exp(0.194 + 0.0936 * ll + 0.0474 * ll2 - 4.21e-3 * ll2 * ll + 1.23e-4 * ll2 * ll2);
}

// The prologue in three parts, so that one walker can be spread over a few threads (kernels.cuh: prologue_kernel):
//   walker_intrinsic   phase / amplitude coefficients of the source               -> out[0 .. C_TAU0), C_AMP, C_COSI
//   walker_detector    antenna pattern, delay, rotation constants of ONE detector -> out[C_TAU0 + i], C_CR0/CI0/FP0/FC0 + i,
//                                                                                    C_ROT0 + 10 (i - 1) ..
//   walker_validity    log_like.py:121-129 and the wide-phase flag                -> out[C_VALID]
// `x` is the walker's row.  Returns whether every intrinsic quantity is finite and the waveform is not identically zero.
__device__ inline bool walker_intrinsic(const DeviceConfig& cfg, const ParamMapDev& pm,
                                        const double* __restrict__ x, double* out) {

    // ---- parameter derivations (waveform.py:334-377) --------------------------------------
    const double q = get(pm, x, BJ_P_Q);
    const double nu = q / pow(1.0 + q, 2.0);
    double mtot;
    if (has(pm, BJ_P_MCHIRP)) mtot = get(pm, x, BJ_P_MCHIRP) / pow(nu, 0.6);
    else                      mtot = get(pm, x, BJ_P_MTOT);

    double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
    if (has(pm, BJ_P_S1X)) s1[0] = get(pm, x, BJ_P_S1X);
    if (has(pm, BJ_P_S1Y)) s1[1] = get(pm, x, BJ_P_S1Y);
    if (has(pm, BJ_P_S1Z)) s1[2] = get(pm, x, BJ_P_S1Z);
    if (has(pm, BJ_P_S2X)) s2[0] = get(pm, x, BJ_P_S2X);
    if (has(pm, BJ_P_S2Y)) s2[1] = get(pm, x, BJ_P_S2Y);
    if (has(pm, BJ_P_S2Z)) s2[2] = get(pm, x, BJ_P_S2Z);
    if (has(pm, BJ_P_S1) && has(pm, BJ_P_TILT1) && has(pm, BJ_P_PHI_1L)) {
        double r = get(pm, x, BJ_P_S1), th = get(pm, x, BJ_P_TILT1), ph = get(pm, x, BJ_P_PHI_1L);
        s1[0] = r * sin(th) * cos(ph); s1[1] = r * sin(th) * sin(ph); s1[2] = r * cos(th);
    }
    if (has(pm, BJ_P_S2) && has(pm, BJ_P_TILT2) && has(pm, BJ_P_PHI_2L)) {
        double r = get(pm, x, BJ_P_S2), th = get(pm, x, BJ_P_TILT2), ph = get(pm, x, BJ_P_PHI_2L);
        s2[0] = r * sin(th) * cos(ph); s2[1] = r * sin(th) * sin(ph); s2[2] = r * cos(th);
    }
    const double lam1 = has(pm, BJ_P_LAMBDA1) ? get(pm, x, BJ_P_LAMBDA1) : 0.0;
    const double lam2 = has(pm, BJ_P_LAMBDA2) ? get(pm, x, BJ_P_LAMBDA2) : 0.0;
    const double dist = get(pm, x, BJ_P_DISTANCE);
    double iota;
    if (has(pm, BJ_P_COS_IOTA)) iota = acos(get(pm, x, BJ_P_COS_IOTA));
    else                        iota = get(pm, x, BJ_P_IOTA);
    const double cosi = cos(iota);
    const double phiref = get(pm, x, BJ_P_PHI_REF);

    // ---- mass / spin combinations (taylorf2.py:176-210,380) -------------------------------
    const double eta = q / ((1.0 + q) * (1.0 + q));
    const double e2 = eta * eta, e3 = e2 * eta;
    const double dlt = sqrt(1.0 - 4.0 * eta);
    const double xa = 0.5 * (1.0 + dlt), xb = 0.5 * (1.0 - dlt);
    const double c1l = s1[2], c2l = s2[2];
    const double c1sq = s1[0] * s1[0] + s1[1] * s1[1] + s1[2] * s1[2];
    const double c2sq = s2[0] * s2[0] + s2[1] * s2[1] + s2[2] * s2[2];
    const double c12 = s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2];
    const double sl = xa * xa * c1l + xb * xb * c2l;
    const double dsl = dlt * (xb * c2l - xa * c1l);

    double sig = eta * (721.0 / 48.0 * c1l * c2l - 247.0 / 48.0 * c12);
    sig += 719.0 / 96.0 * (xa * xa * c1l * c1l + xb * xb * c2l * c2l);
    sig -= 233.0 / 96.0 * (xa * xa * c1sq + xb * xb * c2sq);
    const double so15 = 188.0 * sl / 3.0 + 25.0 * dsl;
    const double gam = (554345.0 / 1134.0 + 110.0 * eta / 9.0) * sl + (13915.0 / 84.0 - 10.0 * eta / 3.0) * dsl;
    double ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * c1l * c2l;
    ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * xa - 120.0 * xa * xa) +
            (-4108.25 / 6.72 - 108.5 / 1.2 * xa + 125.5 / 3.6 * xa * xa)) * xa * xa * c1sq;
    ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * xb - 120.0 * xb * xb) +
            (-4108.25 / 6.72 - 108.5 / 1.2 * xb + 125.5 / 3.6 * xb * xb)) * xb * xb * c2sq;
    const double so3 = kPi * (3760.0 * sl + 1490.0 * dsl) / 3.0 + ss3;
    const double so35 = (-8980424995.0 / 762048.0 + 6586595.0 * eta / 756.0 - 305.0 * e2 / 36.0) * sl -
                        (170978035.0 / 48384.0 - 2876425.0 * eta / 672.0 - 4735.0 * e2 / 144.0) * dsl;

    // ---- v = c * f^(1/3) -------------------------------------------------------------------
    const double c = cbrt(fabs(kPi * mtot * kGT));
    const double lc = log(c);
    double cp[16];
    cp[0] = 1.0;
#pragma unroll
    for (int j = 1; j < 16; ++j) cp[j] = cp[j - 1] * c;
    const double inv2pi = 1.0 / kTwoPi;
    const double K = 3.0 / 128.0 / eta / cp[5] * inv2pi;   // LO = 3/(128 eta v^5), in turns

    // series coefficients s_j of the bracket, L-coefficients l_j, L^2 coefficient ll8
    double s[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { s[j] = 0.0; l[j] = 0.0; }
    double ll8 = 0.0;

    // 3.5PN point mass + spin (taylorf2.py:212-220)
    const double pi2 = kPi * kPi;
    s[0] = 1.0;
    s[2] = 20.0 / 9.0 * (743.0 / 336.0 + 11.0 / 4.0 * eta);
    s[3] = so15 - 16.0 * kPi;
    s[4] = 10.0 * (3058673.0 / 1016064.0 + 5429.0 / 1008.0 * eta + 617.0 / 144.0 * e2 - sig);
    {
        const double g5 = 38645.0 / 756.0 * kPi - 65.0 / 9.0 * eta * kPi - gam;
        // (1 + 3 ln(v / v_lso)) with v_lso = 6^(-1/2):  ln v = lc + L/3
        s[5] = g5 * (1.0 + 3.0 * lc + 1.5 * log(6.0));
        l[5] = g5;
        const double h6 = 11583231236531.0 / 4694215680.0 - 640.0 / 3.0 * pi2 - 6848.0 / 21.0 * kEulerGamma +
                          (-15737765635.0 / 3048192.0 + 2255.0 * pi2 / 12.0) * eta + 76055.0 / 1728.0 * e2 -
                          127825.0 / 1296.0 * e3 + so3;
        s[6] = h6 - 6848.0 / 21.0 * (log(4.0) + lc);
        l[6] = -6848.0 / 63.0;
    }
    s[7] = kPi * (77096675.0 / 254016.0 + 378515.0 / 1512.0 * eta - 74045.0 / 756.0 * e2) + so35;

    const bool order55 = cfg.approx != BJ_APPROX_TAYLORF2_35PN;
    if (order55) {
        // taylorf2.py:269-297 with c_21_3PN = c_22_4PN = c_22_5PN = a_6_c = 0
        const double e4 = e2 * e2;
        const double c8 = -36946947827.5 / 1601901100.8 * e4 + 51004148102.5 / 1310646355.2 * e3 +
                          (30060067316599.7 / 57668439628.8 - 39954.5 / 2721.6 * pi2) * e2 +
                          (-567987228950352.7 / 128152088064.0 - 532292.8 / 396.9 * kEulerGamma +
                           930221.5 / 5443.2 * pi2 - 142068.8 / 44.1 * kLog2 + 2632.5 / 4.9 * kLog3) * eta -
                          9049.0 / 56.7 * pi2 - 3681.2 / 18.9 * kEulerGamma +
                          255071384399888515.3 / 83042553065472.0 - 2632.5 / 19.6 * kLog3 - 101102.0 / 396.9 * kLog2;
        const double c8l = -3.0 * c8;
        const double c8ll = 9.0 * (266146.4 / 1190.7 * eta + 1840.6 / 56.7);
        const double c9 = kPi * (1032375.5 / 19958.4 * e3 + 4529333.5 / 12700.8 * e2 +
                                 (2255.0 / 6.0 * pi2 - 149291726073.5 / 13412044.8) * eta - 640.0 / 3.0 * pi2 -
                                 1369.6 / 2.1 * kEulerGamma + 10534427947316.3 / 1877686272.0 - 2739.2 / 2.1 * kLog2);
        const double c9l = -3.0 * 1369.6 / 6.3 * kPi;
        const double inv13 = 1.0 / (1.0 - 3.0 * eta);
        const double c10 = inv13 * (
            (242506658510205297979.7 / 85723270481616768.0) * e4 * e2 -
            (1272143474037195162.1 / 67631771583129.6) * e4 * eta +
            (1116081080066315514991.3 / 27213736660830720.0 - 943479.7 / 1881.6 * pi2) * e4 +
            (-85710407655931086054085.1 / 3428930819264670720.0 - 614779314.2 / 152806.5 * kEulerGamma -
             46051.9 / 153.6 * pi2 - 4311179766.8 / 152806.5 * kLog2 + 127939.5 / 9.8 * kLog3) * e3 +
            (-1873639936380505730110521.7 / 36575262072156487680.0 - (9923919211.9 / 458419.5) * kEulerGamma +
             41579551.7 / 90316.8 * pi2 - 11734037971.3 / 458419.5 * kLog2 - 5833093.5 / 548.8 * kLog3) * e2 +
            (56993518125966874478111.3 / 1083711468804636672.0 + 6378740752.7 / 916839.0 * kEulerGamma -
             545142954.7 / 812851.2 * pi2 + 15994339707.7 / 1833678.0 * kLog2 + (892417.5 / 313.6) * kLog3) * eta +
            (57822311.5 / 304819.2) * pi2 + (647058264.7 / 2750517.0) * kEulerGamma -
            143300652329540712655.9 / 12630669799587840.0 - 551245.5 / 2195.2 * kLog3 +
            5399283943.1 / 5501034.0 * kLog2);
        const double c10l = 3.0 * inv13 * (1286378036.2 / 458419.5 * e3 + 1384949312.9 / 1375258.5 * e2 -
                                           2427943164.1 / 2750517.0 * eta + 647058264.7 / 8251551.0);
        const double c11 = kPi * (65762707344.5 / 14417109907.2 * e4 - 108059782847.5 / 2621292710.4 * e3 +
                                  512031495514639.7 / 62911025049.6 * e2 +
                                  (-1064790.5 / 3628.8 * e2 + 4501578.5 / 14515.2 * eta - 9439.0 / 56.7) * pi2 +
                                  (-134666.2 / 56.7 * kEulerGamma - 43038370739839704.7 / 3460106377728.0 +
                                   2632.5 / 4.9 * kLog3 - 2100962.6 / 396.9 * kLog2) * eta -
                                  355801.1 / 793.8 * kEulerGamma + 185754140723659441.1 / 27680851021824.0 -
                                  2632.5 / 19.6 * kLog3 - 86254.9 / 113.4 * kLog2);
        const double c11l = -3.0 * kPi * (134666.2 / 170.1 * eta + 355801.1 / 2381.4);
        // ln v = lc + L/3
        s[8] = c8 + c8l * lc + c8ll * lc * lc;
        l[8] = (c8l + 2.0 * c8ll * lc) / 3.0;
        ll8 = c8ll / 9.0;
        s[9] = c9 + c9l * lc;    l[9] = c9l / 3.0;
        s[10] = c10 + c10l * lc; l[10] = c10l / 3.0;
        s[11] = c11 + c11l * lc; l[11] = c11l / 3.0;
    }

    // direct (not LO-scaled) additions to the turn polynomial, index = power of u after * u^5
    double add[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) add[j] = 0.0;

    if (lam1 != 0.0 || lam2 != 0.0) {      // taylorf2.py:394
        const bool tides75 = cfg.approx == BJ_APPROX_TAYLORF2_55PN_75PNTIDES ||
                             cfg.approx == BJ_APPROX_TAYLORF2_55PN_35PNQM_75PNTIDES ||
                             cfg.approx == BJ_APPROX_TAYLORF2_55PN_75PNTIDES2020;
        if (!tides75) {
            // 6PN tides, taylorf2.py:9-36
            double lt, dl;
            tidal_combinations(xa, xb, lam1, lam2, lt, dl);
            s[10] += -39.0 / 2.0 * lt;
            s[12] += -3115.0 / 64.0 * lt + 6595.0 / 364.0 * dl;
        } else {
            // 7.5PN tides, taylorf2.py:38-119
            const double xa2 = xa * xa, xa3 = xa2 * xa, xa4 = xa3 * xa, xa5 = xa4 * xa;
            const double xb2 = xb * xb, xb3 = xb2 * xb, xb4 = xb3 * xb, xb5 = xb4 * xb;
            const double ka = 3.0 * lam1 * xa4 * xb, kb = 3.0 * lam2 * xb4 * xa;
            const double na = -3.0 / (16.0 * eta) * (12.0 + xa / xb);
            const double nb = -3.0 / (16.0 * eta) * (12.0 + xb / xa);
            const double da = 12.0 - 11.0 * xa, db = 12.0 - 11.0 * xb;
            const double p1a = 5.0 * (3179.0 - 919.0 * xa - 2286.0 * xa2 + 260.0 * xa3) / (672.0 * da);
            const double p1b = 5.0 * (3179.0 - 919.0 * xb - 2286.0 * xb2 + 260.0 * xb3) / (672.0 * db);
            double p3a, p3b, p4a;
            const double p4b = -kPi * (27719.0 - 22127.0 * xb + 7022.0 * xb2 - 10232.0 * xb3) / (672.0 * db);
            if (cfg.approx == BJ_APPROX_TAYLORF2_55PN_75PNTIDES2020) {
                p3a = -5.0 * (-387973870.0 + 43246839.0 * xa + 174965616.0 * xa2 + 158378220.0 * xa3 -
                              20427120.0 * xa4 + 4572288.0 * xa5) / 27433728.0 / da;
                p3b = -5.0 * (-387973870.0 + 43246839.0 * xb + 174965616.0 * xb2 + 158378220.0 * xb3 -
                              20427120.0 * xb4 + 4572288.0 * xb5) / 27433728.0 / db;
                // quirk kept (taylorf2.py:115-116): body A and body B use different polynomials
                p4a = -kPi * (27719.0 - 22415.0 * xa + 7598.0 * xa2 - 10520.0 * xa3) / (672.0 * da);
            } else {
                p3a = (39927845.0 / 508032.0 - 480043345.0 / 9144576.0 * xa + 9860575.0 / 127008.0 * xa2 -
                       421821905.0 / 2286144.0 * xa3 + 4359700.0 / 35721.0 * xa4 - 10578445.0 / 285768.0 * xa5) / da;
                p3b = (39927845.0 / 508032.0 - 480043345.0 / 9144576.0 * xb + 9860575.0 / 127008.0 * xb2 -
                       421821905.0 / 2286144.0 * xb3 + 4359700.0 / 35721.0 * xb4 - 10578445.0 / 285768.0 * xb5) / db;
                p4a = -kPi * (27719.0 - 22127.0 * xa + 7022.0 * xa2 - 10232.0 * xa3) / (672.0 * da);
            }
            const double wa = ka * na, wb = kb * nb;
            // Phi_T = v^5 (k0 + k2 v^2 + k3 v^3 + k4 v^4 + k5 v^5)
            add[10] += (wa + wb) * cp[5] * inv2pi;
            add[12] += (wa * p1a + wb * p1b) * cp[7] * inv2pi;
            add[13] += (-kPi) * (wa + wb) * cp[8] * inv2pi;
            add[14] += (wa * p3a + wb * p3b) * cp[9] * inv2pi;
            add[15] += (wa * p4a + wb * p4b) * cp[10] * inv2pi;
        }
        if (cfg.approx == BJ_APPROX_TAYLORF2_55PN_35PNQM_75PNTIDES) {
            // taylorf2.py:121-156: Phi_QM = m/v + p v + t v^2
            const double at1 = xa * c1l, at2 = xb * c2l;
            const double q1 = quadrupole_yy(lam1) - 1.0, q2 = quadrupole_yy(lam2) - 1.0;
            const double plus = at1 * at1 * q1 + at2 * at2 * q2;
            const double minus = at1 * at1 * q1 - at2 * at2 * q2;
            const double m = -75.0 / (64.0 * eta) * plus;
            const double p = ((45.0 / 16.0 * eta + 15635.0 / 896.0) * plus + 2215.0 / 512 * dlt * minus) / eta;
            const double t = -75.0 / (8.0 * eta) * plus * kPi;
            add[4] += m / c * inv2pi;
            add[6] += p * c * inv2pi;
            add[7] += t * cp[2] * inv2pi;
        }
    }

#pragma unroll
    for (int j = 0; j < 16; ++j) out[C_A0 + j] = K * s[j] * cp[j] + add[j];
#pragma unroll
    for (int j = 5; j < 12; ++j) out[C_B0 + (j - 5)] = K * l[j] * cp[j];
    out[C_C8] = K * ll8 * cp[8];
    out[C_K0] = phiref * inv2pi;

    // ---- amplitude (taylorf2.py:300-345) -----------------------------------------------------
    {
        const double cs = 0.5 * (c1l + c2l), ca = 0.5 * (c1l - c2l);
        const double be = 113.0 / 12.0 * (cs + dlt * ca - 76.0 / 113.0 * eta * cs);
        const double sg = ca * ca * (81.0 / 16.0 - 20.0 * eta) + 81.0 / 8.0 * ca * cs * dlt +
                          cs * cs * (81.0 / 16.0 - eta / 4.0);
        const double ep = dlt * ca * (502429.0 / 16128.0 - 907.0 / 192.0 * eta) +
                          cs * (5.0 / 48.0 * e2 - 73921.0 / 2016.0 * eta + 502429.0 / 16128.0);
        const double al2 = 11.0 / 8.0 * eta + 743.0 / 672.0;
        const double al3 = be / 2.0 - 2.0 * kPi;
        const double al4 = 1379.0 / 1152.0 * e2 + 18913.0 / 16128.0 * eta + 7266251.0 / 8128512.0 - sg / 2.0;
        const double al5 = 57.0 / 16.0 * kPi * eta - 4757.0 * kPi / 1344.0 + ep;
        const double al6 = 856.0 / 105.0 * kEulerGamma + 67999.0 / 82944.0 * e3 - 1041557.0 / 258048.0 * e2 -
                           451.0 / 96.0 * pi2 * eta + 10.0 * pi2 / 3.0 + 3526813753.0 / 27869184.0 * eta -
                           29342493702821.0 / 500716339200.0;
        const double al6l = 856.0 / 105.0;
        const double al7 = (-1349.0 / 24192.0 * e2 - 72221.0 / 24192.0 * eta - 5111593.0 / 2709504.0) * kPi;
        out[C_Q2 + 0] = al2 * cp[2];
        out[C_Q2 + 1] = al3 * cp[3];
        out[C_Q2 + 2] = al4 * cp[4];
        out[C_Q2 + 3] = al5 * cp[5];
        out[C_Q2 + 4] = (al6 + al6l * (log(4.0) + lc)) * cp[6];
        out[C_Q2 + 5] = al7 * cp[7];
        out[C_Q6L] = al6l * cp[6] / 3.0;
        const double mc = mtot * pow(fabs(eta), 3.0 / 5.0);
        out[C_AMP] = kPre * (pow(fabs(mc), 5.0 / 6.0) / dist / pow(kPi, 2.0 / 3.0) * sqrt(5.0 / 24.0));
    }

    const double incl_plus = (1.0 + cosi * cosi) * 0.5;
    out[C_COSI] = cosi;
    bool ok = true;
    for (int j = 0; j < C_TAU0; ++j) ok = ok && isfinite(out[j]);
    ok = ok && isfinite(out[C_AMP]);
    const double hp_scale = out[C_AMP] * incl_plus;
    ok = ok && isfinite(hp_scale) && (hp_scale != 0.0) && isfinite(out[C_AMP] * cosi);
    // quirk (taylorf2.py:252): the 5.5PN phase takes (pi M f gt)^(1/3) WITHOUT abs -> NaN for M < 0
    if (order55 && !(mtot > 0.0)) ok = false;
    return ok;
}

// geocentric arrival-time offset of detector i: delay + time_shift
__device__ __forceinline__ double walker_tau(const DeviceConfig& cfg, const ParamMapDev& pm, const double* __restrict__ x,
                                             int i) {
    const double tshift = get(pm, x, BJ_P_TIME_SHIFT);
    return geocentric_delay(cfg.det[i], get(pm, x, BJ_P_RA), get(pm, x, BJ_P_DEC), get(pm, x, BJ_P_T_GPS) + tshift) + tshift;
}

// detector projection (detector.py:394-418, taylorf2.py:427-432) of detector i; o5 = {tau, Re C, Im C, F+, Fx},
// rot10 (i >= 1, uniform grids) = the rotation constants against detector 0.  Returns finiteness.
__device__ inline bool walker_detector(const DeviceConfig& cfg, const ParamMapDev& pm, const double* __restrict__ x, int i,
                                       double* o5, double* rot10) {
    double iota;
    if (has(pm, BJ_P_COS_IOTA)) iota = acos(get(pm, x, BJ_P_COS_IOTA));
    else                        iota = get(pm, x, BJ_P_IOTA);
    const double cosi = cos(iota);
    const double incl_plus = (1.0 + cosi * cosi) * 0.5;
    const double ra = get(pm, x, BJ_P_RA), dec = get(pm, x, BJ_P_DEC), psi = get(pm, x, BJ_P_PSI);
    const double t_arr = get(pm, x, BJ_P_T_GPS) + get(pm, x, BJ_P_TIME_SHIFT);
    double fp, fc;
    antenna_pattern(cfg.det[i], ra, dec, psi, t_arr, fp, fc);
    const double tau = walker_tau(cfg, pm, x, i);
    o5[0] = tau;
    o5[1] = fp * incl_plus;      // C_i = F+ (1+cos^2 i)/2 - i Fx cos i
    o5[2] = -fc * cosi;
    o5[3] = fp;
    o5[4] = fc;
    // rotation constants: the phase of detector i differs from detector 0's by 2 pi f (tau_i - tau_0) only
    if (i >= 1 && rot10 != nullptr) {
        if (cfg.grid_df > 0.0) {
            const double turns_per_bin = cfg.grid_df * (tau - walker_tau(cfg, pm, x, 0));
            const double offs[4] = {3.5, 4.5, 11.5, 12.5};
#pragma unroll 1
            for (int a = 0; a < 4; ++a) {
                double sh, ch, sn, cs;
                sincospi(offs[a] * turns_per_bin, &sh, &ch);          // half angle: cos x - 1 = -2 sin^2(x / 2)
                sincospi(2.0 * offs[a] * turns_per_bin, &sn, &cs);
                rot10[a] = -2.0 * sh * sh;
                rot10[4 + a] = sn;
            }
            double sn, cs;
            sincospi(64.0 * turns_per_bin, &sn, &cs);
            rot10[8] = cs;
            rot10[9] = -sn;
        } else {
#pragma unroll 1
            for (int a = 0; a < kRotSlots; ++a) rot10[a] = 0.0;
        }
    }
    return isfinite(tau) && isfinite(fp) && isfinite(fc);
}

// C_VALID: 0 = -inf (log_like.py:121-129: NaN/Inf anywhere in hp/hc, or hp identically zero), 1 = evaluate, 2 = evaluate
// with the wide-phase path.  The fast kernel reads the binary angle from the low word of (turns + 1.5 * 2^20), exact while
// |turns| < 2^19; every term of the phase is monotone in f, so its magnitude is bounded by the larger of the two band ends.
__device__ inline double walker_validity(const DeviceConfig& cfg, const double* out, bool ok, double tmax) {
    double wide = 0.0;
    if (ok && cfg.f_hi > 0.0) {
        double bound = 0.0;
        const double fe[2] = {cfg.f_lo, cfg.f_hi};
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
            const double u = cbrt(fe[e]), L = fabs(log(fe[e]));
            const double w5 = 1.0 / (u * u * u * u * u);
            double up = 1.0, sa = 0.0, sb = 0.0;
#pragma unroll 1
            for (int j = 0; j < 16; ++j) {
                sa += fabs(out[C_A0 + j]) * up;
                if (j < 7) sb += fabs(out[C_B0 + j]) * up;
                up *= u;
            }
            bound = fmax(bound, w5 * sa + L * sb + L * L * fabs(out[C_C8]) * u * u * u);
        }
        bound += fabs(out[C_K0]) + cfg.f_hi * tmax;
        if (!(bound < 500000.0)) wide = 1.0;
    }
    return ok ? 1.0 + wide : 0.0;
}

// one thread does it all (single-point entry points); fills out[C_NUM]
__device__ inline void walker_coefficients(const DeviceConfig& cfg, const ParamMapDev& pm,
                                           const double* __restrict__ x, double* out) {
#pragma unroll 1
    for (int i = 0; i < C_NUM; ++i) out[i] = 0.0;
    bool ok = walker_intrinsic(cfg, pm, x, out);
    double tmax = 0.0;
    for (int i = 0; i < cfg.n_ifo; ++i) {
        double o5[5];
        ok = walker_detector(cfg, pm, x, i, o5, i >= 1 ? out + C_ROT0 + kRotSlots * (i - 1) : nullptr) && ok;
        out[C_TAU0 + i] = o5[0];
        out[C_CR0 + i] = o5[1];
        out[C_CI0 + i] = o5[2];
        out[C_FP0 + i] = o5[3];
        out[C_FC0 + i] = o5[4];
        tmax = fmax(tmax, fabs(o5[0]));
    }
    out[C_VALID] = walker_validity(cfg, out, ok, tmax);
}

}  // namespace bj
