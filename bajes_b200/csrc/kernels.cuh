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
// per-bin record layout (doubles):  [u, L, w5, f, (dr_i, di_i) x NIFO]
//   u = f^(1/3), L = ln f, w5 = f^(-5/3)
//   dr + i di = (4/T) conj(d_i)/S_i * f^(-7/6) * exp(+i pi f T)      (centre shift folded in)
// <h|h> needs no per-bin work at all: |h|^2 = amp^2 |C_i|^2 f^(-7/3) Q^2 and Q is a polynomial in
// (u, L) with per-walker coefficients, so <h|h>_i = sum_{m<=n} g_m g_n G_i[m][n] with the Gram moments
// G_i[m][n] = sum_k phi_m phi_n (4/T) f^(-7/3) / S_i tabulated once at set-up (kGram entries per detector).
// ----------------------------------------------------------------------------------------------
template <int NIFO> struct Rec {
    static constexpr int kDoubles = 4 + 2 * NIFO;                 // even => 16-byte aligned records
    static constexpr int kBytes = kDoubles * 8;
};
constexpr int kBasis = 8;                         // phi = {1, u^2, u^3, u^4, u^5, u^6, u^7, L u^6}
constexpr int kGram = kBasis * (kBasis + 1) / 2;  // upper triangle, row-major

constexpr int kThreads = 128;        // walkers per CTA
constexpr int kTileBins = 64;        // bins per smem stage
constexpr int kStages = 4;
constexpr double kMagic = 1572864.0; // 1.5 * 2^20: ulp(t + kMagic) == 2^-32 turns for |t| < 2^19

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

// ---- packed FP32x2 helpers (Blackwell FFMA2 / FMUL2) -----------------------------------------------
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ unsigned long long neg2(unsigned long long a) {
    // folds into the operand-negate modifier of FFMA2
    unsigned long long d;
    asm("{.reg .f32 lo, hi; mov.b64 {lo, hi}, %1; neg.f32 lo, lo; neg.f32 hi, hi; mov.b64 %0, {lo, hi};}" : "=l"(d) : "l"(a));
    return d;
}

// ---- Q e^{-i theta} from a 32-bit binary angle (turns * 2^32, two's complement) ---------------------
// Returns cq = Q cos(theta), sq = Q sin(theta) as doubles.
//   POLY32 : theta = pi z + pi n with z = sext31(b) 2^-31 in [-1/2, 1/2); sin(pi z)/z and cos(pi z) are
//            degree-5 minimax polynomials in z^2 evaluated TOGETHER with packed FP32x2 FMAs; the (-1)^n
//            sign and the amplitude Q ride on one packed multiply (scripts/fit_sincospi.py: max abs error
//            1.9e-7, rms 4e-8, norm bias 2e-8).  No quadrant selects, no predicates.
//   MUFU   : MUFU.SIN / MUFU.COS (fastest per evaluation but biased: fails the parity tolerance on some configs)
//   FP64   : sincospi in double (validation)
template <int SC>
__device__ __forceinline__ void scaled_phasor(int32_t b, double Q, float Qf, double& cq, double& sq) {
    if constexpr (SC == BJ_SINCOS_POLY32) {
        int32_t rem;
        asm("bfe.s32 %0, %1, 0, 31;" : "=r"(rem) : "r"(b));            // SGXT: nearest-half-turn remainder
        const uint32_t ub = (uint32_t)b;
        const uint32_t flip = (ub ^ (ub + ub)) & 0x80000000u;           // parity of the half turn
        const float qs = __uint_as_float(__float_as_uint(Qf) ^ flip);
        const float z = __int2float_rn(rem) * 4.656612873077393e-10f;   // 2^-31
        const float z2 = z * z;
        const float zq = z * qs;
        const unsigned long long zz = pack2(z2, z2);
        // FP32-aware minimax coefficients (scripts/fit_sincospi.py); lanes = (sin(pi z)/z, cos(pi z))
        unsigned long long p = pack2(-6.772854831e-03f, -2.412191778e-02f);
        p = fma2(p, zz, pack2(8.189047128e-02f, 2.347589880e-01f));
        p = fma2(p, zz, pack2(-5.992177725e-01f, -1.335172534e+00f));
        p = fma2(p, zz, pack2(2.550160408e+00f, 4.058705807e+00f));
        p = fma2(p, zz, pack2(-5.167712688e+00f, -4.934802055e+00f));
        p = fma2(p, zz, pack2(-8.747612412e-08f, 1.000000000e+00f));  // sine: low word of pi ; cosine: 1
        p = mul2(p, pack2(zq, qs));
        float sf, cf;
        unpack2(p, sf, cf);
        sf = fmaf(zq, 3.141592741e+00f, sf);                            // + high word of pi
        sq = (double)sf;
        cq = (double)cf;
    } else if constexpr (SC == BJ_SINCOS_MUFU) {
        const float x = __int2float_rn(b) * 1.4629180792671596e-9f;     // 2 pi / 2^32  -> [-pi, pi]
        float sf, cf;
        __sincosf(x, &sf, &cf);
        sq = (double)(sf * Qf);
        cq = (double)(cf * Qf);
    } else {
        const double r = (double)b * 2.3283064365386963e-10;            // 2^-32 -> turns in [-0.5, 0.5)
        double s, c;
        sincospi(2.0 * r, &s, &c);
        sq = Q * s;
        cq = Q * c;
    }
}

// binary angle of an arbitrary turn count (no range limit): exact reduction t - rint(t), then round to 2^-32
__device__ __forceinline__ int32_t robust_angle(double turns) {
    const double r = turns - rint(turns);
    return (int32_t)__double2ll_rn(r * 4294967296.0);      // r = +-0.5 wraps to the same angle
}

// Same phasors, returned as FP32 (cq, sq) for the mixed-precision kernel (no FP64 conversion at all).
template <int SC>
__device__ __forceinline__ void scaled_phasor_f32(int32_t b, float Qf, float& cq, float& sq) {
    if constexpr (SC == BJ_SINCOS_POLY32) {
        int32_t rem;
        asm("bfe.s32 %0, %1, 0, 31;" : "=r"(rem) : "r"(b));            // SGXT: nearest-half-turn remainder
        const uint32_t ub = (uint32_t)b;
        const uint32_t flip = (ub ^ (ub + ub)) & 0x80000000u;           // parity of the half turn
        const float qs = __uint_as_float(__float_as_uint(Qf) ^ flip);
        const float z = __int2float_rn(rem) * 4.656612873077393e-10f;   // 2^-31
        const float z2 = z * z;
        const float zq = z * qs;
        const unsigned long long zz = pack2(z2, z2);
        // FP32-aware minimax coefficients (scripts/fit_sincospi.py); lanes = (sin(pi z)/z, cos(pi z))
        unsigned long long p = pack2(-6.772854831e-03f, -2.412191778e-02f);
        p = fma2(p, zz, pack2(8.189047128e-02f, 2.347589880e-01f));
        p = fma2(p, zz, pack2(-5.992177725e-01f, -1.335172534e+00f));
        p = fma2(p, zz, pack2(2.550160408e+00f, 4.058705807e+00f));
        p = fma2(p, zz, pack2(-5.167712688e+00f, -4.934802055e+00f));
        p = fma2(p, zz, pack2(-8.747612412e-08f, 1.000000000e+00f));  // sine: low word of pi ; cosine: 1
        p = mul2(p, pack2(zq, qs));
        float sf, cf;
        unpack2(p, sf, cf);
        sf = fmaf(zq, 3.141592741e+00f, sf);                            // + high word of pi
        sq = sf;
        cq = cf;
    } else {
#ifdef BJ_DIAG_EXACT_PHASOR
        // diagnostic build (scripts/diag_tail.py): FP64 sin/cos rounded once -- isolates the phasor error from the rest
        double sd, cd;
        sincospi(2.0 * ((double)b * 2.3283064365386963e-10), &sd, &cd);
        sq = (float)sd * Qf;
        cq = (float)cd * Qf;
#else
        const float x = __int2float_rn(b) * 1.4629180792671596e-9f;
        float sf, cf;
        __sincosf(x, &sf, &cf);
        sq = sf * Qf;
        cq = cf * Qf;
#endif
    }
}

// ---- calibration of the FP32 phasor bias -------------------------------------------------------------
// The hardware / polynomial phasor e^ = c^ - i s^ equals the true one times (1 + eps_a - i eps_p) on average
// over the circle (MUFU: eps_a = -7.34e-8 in EVERY octant; scripts/sincos_error.cu).  A constant factor on
// every term is a constant factor on <d|h>: it is measured once here and divided out in the epilogue, which
// leaves only zero-mean rounding noise.  out[0] += c^ c + s^ s - 1, out[1] += s^ c - c^ s over 2^22 angles.
template <int SC>
__global__ void calibrate_kernel(double* out) {
    double ea = 0.0, ep = 0.0;
    const int n = 1 << 22;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // golden-ratio sequence: equidistributed over the circle AND in the low bits (a fixed low-bit pattern
        // would bias the int->float rounding of the remainder and with it the measured phase)
        const int32_t b = (int32_t)((uint32_t)i * 2654435769u + 0x7F4A7C15u);
        float cq, sq;
        scaled_phasor_f32<SC>(b, 1.0f, cq, sq);
        double s, c;
        sincospi(2.0 * ((double)b * 2.3283064365386963e-10), &s, &c);
        ea += (double)cq * c + (double)sq * s - 1.0;
        ep += (double)sq * c - (double)cq * s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ea += __shfl_xor_sync(0xffffffffu, ea, o);
        ep += __shfl_xor_sync(0xffffffffu, ep, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], ea / n);
        atomicAdd(&out[1], ep / n);
    }
}

// ---- prologue ---------------------------------------------------------------------------------
// One walker = 1 + BJ_MAX_IFO threads: role 0 derives the source's phase / amplitude coefficients, role 1 + i the antenna
// pattern, delay and rotation constants of detector i (all independent given the walker's row), then role 0 combines the
// validity flags.  A call with few walkers is latency-bound in this kernel: the split shortens its dependent chain ~2x.
constexpr int kSteepChecks = 32;   // check frequencies of classify_kernel
constexpr int kMaxLevels = 2;      // compressed axes per context
constexpr int kProWalkers = 32;
__global__ void __launch_bounds__(kProWalkers * (1 + BJ_MAX_IFO))
prologue_kernel(DeviceConfig cfg, ParamMapDev pm, const double* __restrict__ x, int ndim, long n_walkers,
                double* __restrict__ coef, long stride, int* __restrict__ list_counts) {
    // work-list counters of classify_kernel (contexts with compressed axes), zeroed here to save a memset node
    if (list_counts && blockIdx.x == 0 && threadIdx.y == 0 && threadIdx.x <= kMaxLevels + 1) list_counts[threadIdx.x] = 0;
    __shared__ double s_tau[BJ_MAX_IFO][kProWalkers];
    __shared__ int s_ok[BJ_MAX_IFO][kProWalkers];
    const int lane = threadIdx.x, role = threadIdx.y;
    const long w = (long)blockIdx.x * kProWalkers + lane;
    const bool live = w < n_walkers;
    const double* xr = x + (live ? w : 0) * ndim;
    double out[C_ROT0];
    bool ok = true;
    if (role == 0) {
#pragma unroll 1
        for (int j = 0; j < C_ROT0; ++j) out[j] = 0.0;
        ok = walker_intrinsic(cfg, pm, xr, out);
    } else {
        const int i = role - 1;
        double o5[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, rot[kRotSlots];
#pragma unroll 1
        for (int a = 0; a < kRotSlots; ++a) rot[a] = 0.0;
        bool okd = true;
        if (i < cfg.n_ifo) okd = walker_detector(cfg, pm, xr, i, o5, i >= 1 ? rot : nullptr);
        s_tau[i][lane] = fabs(o5[0]);
        s_ok[i][lane] = okd ? 1 : 0;
        if (live) {
            coef[(C_TAU0 + i) * stride + w] = o5[0];
            coef[(C_CR0 + i) * stride + w] = o5[1];
            coef[(C_CI0 + i) * stride + w] = o5[2];
            coef[(C_FP0 + i) * stride + w] = o5[3];
            coef[(C_FC0 + i) * stride + w] = o5[4];
            if (i >= 1) {
#pragma unroll 1
                for (int a = 0; a < kRotSlots; ++a) coef[(size_t)(C_ROT0 + kRotSlots * (i - 1) + a) * stride + w] = rot[a];
            }
        }
    }
    __syncthreads();
    if (role == 0 && live) {
        double tmax = 0.0;
        for (int i = 0; i < cfg.n_ifo; ++i) {
            ok = ok && s_ok[i][lane] != 0;
            tmax = fmax(tmax, s_tau[i][lane]);
        }
        out[C_VALID] = walker_validity(cfg, out, ok, tmax);
#pragma unroll 1
        for (int j = 0; j < C_TAU0; ++j) coef[j * stride + w] = out[j];
        coef[C_AMP * stride + w] = out[C_AMP];
        coef[C_VALID * stride + w] = out[C_VALID];
        coef[C_COSI * stride + w] = out[C_COSI];
    }
}

// calibration-envelope node values and PSD weights of every walker (contexts created with extras)
__global__ void prologue_extras_kernel(DeviceConfig cfg, ExtrasDims dims, ExtraMapDev em, const double* __restrict__ x,
                                       int ndim, long n_walkers, double* __restrict__ coef,
                                       double* __restrict__ coef_ex, long stride) {
    const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    if (!walker_extras(cfg, dims, em, x + w * ndim, coef_ex, stride, w)) coef[C_VALID * stride + w] = 0.0;
}

// per-chunk description: records [start, start + len) of the table, all in one (weight band, calibration segment) cell
struct ChunkDesc {
    int start, len, band, seg;
};

// ---- compressed frequency axes: which table evaluates which walker ----------------------------------------------------
// A compressed axis (capi.cu: build_compressed_axis) replaces runs of bins by Chebyshev nodes whose spacing is chosen for
// a DESIGN bound on the phase slope  |d(Phi/2pi)/df| + |tau_i| <= budget(f) = margin t_N(f; mchirp_min) + tau_max.  A
// context holds up to kMaxLevels such axes with growing budgets, and the full grid.  classify_kernel -- one warp per
// walker, one lane per check frequency (log-spaced over the band, both ends included) -- evaluates the slope of the
// walker's own phase series  P = w5 sum_j A_j u^j + L sum_m B_m u^m + L^2 C8 u^3 :
//     dP/df = [ w5 sum_j A_j (j-5)/3 u^j + sum_m B_m u^m (1 + L m/3) + C8 u^3 (2 L + L^2) ] / f ,
// and appends the walker to the work list of the first level whose budget it respects everywhere, else (or when its phase
// is too wide for the binary-angle trick, C_VALID == 2, or anything is NaN) to the list of the full grid.  Every bin
// kernel then runs once per list, on that list's table (Sel below); walkers that return -inf anyway go to level 0.
// The order of a list varies from run to run (atomics), the results do not: no kernel mixes walkers.
struct SteepTable {
    int n, n_levels;
    double u[kSteepChecks], L[kSteepChecks], w5[kSteepChecks], inv_f[kSteepChecks];
    double budget[kMaxLevels][kSteepChecks];
    // focus table (capi.cu: bj_focus): budget on the slope RELATIVE to a fiducial point, whose coefficients and delays are
    // fid = {A_0..15, B_0..6, C8, tau_0..2}; focus == 0: none.  Its work list is number n_levels + 1.
    int focus;
    double budget_focus[kSteepChecks];
    double fid[24 + BJ_MAX_IFO];
};
// work list of one table: idx[0 .. *count) are walker indices (idx == nullptr: all walkers, in order)
struct Sel {
    const int* idx;
    const int* count;
};
// One launch over ALL tables of a context with compressed axes: CTA (x, y) serves chunk x of the table that owns walker
// tile y, tiles being numbered table after table (the lists' lengths are only known on the device).  The grid is sized
// for the worst case -- ceil(W / tile) + n - 1 tiles, the longest chunk list -- and surplus CTAs return at once.
struct TableRef {
    const unsigned char* basis;      // tensor-core tables (kernels.cuh: TileRec)
    const unsigned char* data;
    const unsigned char* rec_ul;     // Horner records (u, L of a bin: amplitude correction)
    const double* rec64;             // FP64 window records, offset so that chunk.start indexes them
    const ChunkDesc* chunks;
    int n_chunks, n_precise;
    double dh_fix_re, dh_fix_im;
};
struct MultiTable {
    int n;                           // 0: single-table launch, the explicit kernel arguments apply
    long list_stride;
    const int* lists;                // [n][list_stride]
    const int* counts;               // [n]
    TableRef t[kMaxLevels + 2];      // compressed levels, the full grid, the focus table
};
__device__ __forceinline__ bool resolve_tile(const MultiTable& mt, int tile, int by, int& lv, long& entry0, long& n_sel) {
    int t0 = 0;
    for (lv = 0; lv < mt.n; ++lv) {
        n_sel = (long)mt.counts[lv];
        const int nt = (int)((n_sel + tile - 1) / tile);
        if (by < t0 + nt) {
            entry0 = (long)(by - t0) * tile;
            return true;
        }
        t0 += nt;
    }
    return false;
}
__global__ void __launch_bounds__(128) classify_kernel(SteepTable tab, int n_ifo, double* __restrict__ coef, long stride,
                                                       long n_walkers, int* __restrict__ lists, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_walkers) return;
    const int c = lane < tab.n ? lane : 0;
    const double u = tab.u[c], L = tab.L[c];
    // all loads first (independent, warp-uniform addresses), then the arithmetic
    double a[16], b[7], tau[BJ_MAX_IFO];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = coef[(size_t)(C_A0 + j) * stride + w];
#pragma unroll
    for (int j = 0; j < 7; ++j) b[j] = coef[(size_t)(C_B0 + j) * stride + w];
    const double c8 = coef[(size_t)C_C8 * stride + w];
#pragma unroll
    for (int i = 0; i < BJ_MAX_IFO; ++i) tau[i] = i < n_ifo ? coef[(size_t)(C_TAU0 + i) * stride + w] : 0.0;
    const double valid = coef[(size_t)C_VALID * stride + w];
    double up = 1.0, sa = 0.0, sb = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        sa = fma(a[j] * ((double)(j - 5) * (1.0 / 3.0)), up, sa);
        if (j < 7) sb = fma(b[j] * fma(L, (double)j * (1.0 / 3.0), 1.0), up, sb);
        up *= u;
    }
    const double slope = tab.inv_f[c] * (tab.w5[c] * sa + sb + c8 * u * u * u * (2.0 * L + L * L));
    double tmax = 0.0;
#pragma unroll
    for (int i = 0; i < BJ_MAX_IFO; ++i) tmax = fmax(tmax, fabs(tau[i]));
    const double need = fabs(slope) + tmax;
    bool in_focus = false;
    if (tab.focus) {
        double upf = 1.0, saf = 0.0, sbf = 0.0, dtau = 0.0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            saf = fma((a[j] - tab.fid[j]) * ((double)(j - 5) * (1.0 / 3.0)), upf, saf);
            if (j < 7) sbf = fma((b[j] - tab.fid[16 + j]) * fma(L, (double)j * (1.0 / 3.0), 1.0), upf, sbf);
            upf *= u;
        }
        const double rel = tab.inv_f[c] * (tab.w5[c] * saf + sbf + (c8 - tab.fid[23]) * u * u * u * (2.0 * L + L * L));
#pragma unroll
        for (int i = 0; i < BJ_MAX_IFO; ++i)
            if (i < n_ifo) dtau = fmax(dtau, fabs(tau[i] - tab.fid[24 + i]));
        const bool badf = lane < tab.n && !(fabs(rel) + dtau <= tab.budget_focus[c]);
        in_focus = !__any_sync(0xffffffffu, badf);
    }
    int level = tab.n_levels;                      // n_levels = the full grid
    if (valid == 0.0) level = 0;
    else if (valid == 1.0 && in_focus) level = tab.n_levels + 1;
    else if (valid == 1.0) {
        for (int lv = tab.n_levels - 1; lv >= 0; --lv) {
            const bool bad = lane < tab.n && !(need <= tab.budget[lv][c]);
            if (!__any_sync(0xffffffffu, bad)) level = lv;
        }
    }
    if (lane == 0) {
        coef[(size_t)C_STEEP * stride + w] = (double)level;
        const int pos = atomicAdd(&counts[level], 1);
        lists[(size_t)level * stride + pos] = (int)w;
    }
}


// ---- the fused full-grid kernel -----------------------------------------------------------------
// One CTA = kThreads walkers (one per thread, coefficients in registers) x one frequency chunk.
// Per bin and walker: phase (FP64 Horner, + magic constant so that the low word of the double IS the binary
// angle), amplitude series, and per detector one FP64 FMA for the time-delay ramp, one scaled phasor, and a
// 4-DFMA complex multiply-accumulate with the static data record.  NB bins are processed in lockstep so the
// scheduler always has independent DFMA chains to interleave.
template <int NIFO, int MODEL>
struct WalkerRegs {
    static constexpr int NA = (MODEL == MODEL_35) ? 8 : 16;
    static constexpr int NBQ = (MODEL == MODEL_35) ? 2 : 7;
    double a[NA], a10, a12, bq[NBQ], c8, k0m, q[6], q6l, tau[NIFO];
};

// calibration envelope of the chunk's segment, FP64: cal(f) = c0 + dc * clamp((f - f0) * inv_width, 0, 1)
template <int NIFO>
struct Spcal64 {
    double c0r[NIFO], c0i[NIFO], dcr[NIFO], dci[NIFO], f0, inv_width;
};

template <int NIFO, int MODEL, int SC, int NB, bool SPCAL>
__device__ __forceinline__ void process_bins(const double* __restrict__ rec, const WalkerRegs<NIFO, MODEL>& w,
                                             const Spcal64<NIFO>& sp, double (&zr)[NIFO], double (&zi)[NIFO]) {
    constexpr int RS = Rec<NIFO>::kDoubles;
    double u[NB], L[NB], w5[NB], f[NB], u2[NB], ph[NB], Q[NB];
    float Qf[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const double2* r = reinterpret_cast<const double2*>(rec + j * RS);
        const double2 h0 = r[0], h1 = r[1];
        u[j] = h0.x; L[j] = h0.y; w5[j] = h1.x; f[j] = h1.y;
        u2[j] = u[j] * u[j];
    }
    // ---- phase in turns (+ magic) ----
    if constexpr (MODEL == MODEL_35) {
        double p[NB], bl[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) p[j] = fma(w.a12, u2[j], w.a10);       // tidal pair enters at u^7 * u^3
#pragma unroll
        for (int j = 0; j < NB; ++j) p[j] = fma(p[j], u2[j] * u[j], w.a[7]);
#pragma unroll
        for (int k = 6; k >= 2; --k) {
#pragma unroll
            for (int j = 0; j < NB; ++j) p[j] = fma(p[j], u[j], w.a[k]);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            p[j] = fma(p[j], u2[j], w.a[0]);
            bl[j] = fma(w.bq[1], u[j], w.bq[0]);
            ph[j] = fma(w5[j], p[j], fma(L[j], bl[j], w.k0m));
        }
    } else {
        double p[NB], bl[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) { p[j] = w.a[15]; bl[j] = w.bq[6]; }
#pragma unroll
        for (int k = 14; k >= 0; --k) {
#pragma unroll
            for (int j = 0; j < NB; ++j) p[j] = fma(p[j], u[j], w.a[k]);
        }
#pragma unroll
        for (int k = 5; k >= 0; --k) {
#pragma unroll
            for (int j = 0; j < NB; ++j) bl[j] = fma(bl[j], u[j], w.bq[k]);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            bl[j] = fma(L[j], w.c8 * (u2[j] * u[j]), bl[j]);
            ph[j] = fma(w5[j], p[j], fma(L[j], bl[j], w.k0m));
        }
    }
    // ---- amplitude series ----
    {
        double qq[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) qq[j] = fma(w.q6l, L[j], w.q[4]);
#pragma unroll
        for (int j = 0; j < NB; ++j) qq[j] = fma(w.q[5], u[j], qq[j]);
#pragma unroll
        for (int k = 3; k >= 0; --k) {
#pragma unroll
            for (int j = 0; j < NB; ++j) qq[j] = fma(qq[j], u[j], w.q[k]);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            Q[j] = fma(u2[j], qq[j], 1.0);
            Qf[j] = (float)Q[j];
        }
    }
    // ---- per detector ----
#pragma unroll
    for (int i = 0; i < NIFO; ++i) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const double tt = fma(f[j], w.tau[i], ph[j]);
            double cq, sq;
            if constexpr (SC == BJ_SINCOS_FP64) {
                // validation mode: no binary-angle trick at all (k0m carries no magic constant here)
                const double r = tt - rint(tt);
                double sn, cs;
                sincospi(2.0 * r, &sn, &cs);
                sq = Q[j] * sn;
                cq = Q[j] * cs;
            } else {
                scaled_phasor<SC>(__double2loint(tt), Q[j], Qf[j], cq, sq);
            }
            if constexpr (SPCAL) {
                // h *= cal  <=>  (cq - i sq) *= (calr + i cali)      (detector.py:517-519)
                const double t = fmin(fmax((f[j] - sp.f0) * sp.inv_width, 0.0), 1.0);
                const double calr = fma(sp.dcr[i], t, sp.c0r[i]), cali = fma(sp.dci[i], t, sp.c0i[i]);
                const double c2 = cq * calr + sq * cali, s2 = sq * calr - cq * cali;
                cq = c2;
                sq = s2;
            }
            const double2 d = *reinterpret_cast<const double2*>(rec + j * RS + 4 + 2 * i);
            // (dr + i di) * Q (c - i s)
            zr[i] = fma(d.x, cq, zr[i]);
            zr[i] = fma(d.y, sq, zr[i]);
            zi[i] = fma(d.y, cq, zi[i]);
            zi[i] = fma(-d.x, sq, zi[i]);
        }
    }
}

template <int NIFO, int MODEL, int SC, bool SPCAL>
__global__ void __launch_bounds__(kThreads, 4) fullgrid_kernel(const double* __restrict__ rec,
                                                                 const ChunkDesc* __restrict__ chunks,
                                                                 const double* __restrict__ coef, long stride,
                                                                 long n_walkers, double* __restrict__ partial,
                                                                 const double* __restrict__ coef_ex, ExtrasDims dims,
                                                                 const double* __restrict__ spcal_freqs, int chunk0,
                                                                 Sel sel, const double* __restrict__ fid) {
    constexpr int RS = Rec<NIFO>::kDoubles;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kTileBins * RS * 8);

    // work list (Sel): entry -> walker; a CTA beyond the end of the list has nothing to do
    const long n_sel = sel.idx ? (long)*sel.count : n_walkers;
    if ((long)blockIdx.y * kThreads >= n_sel) return;
    const long entry = (long)blockIdx.y * kThreads + threadIdx.x;
    const bool live = entry < n_sel;
    const long wl = sel.idx ? (long)sel.idx[live ? entry : n_sel - 1] : (live ? entry : n_sel - 1);
    const long wi = wl;

    WalkerRegs<NIFO, MODEL> w;
#pragma unroll
    for (int j = 0; j < WalkerRegs<NIFO, MODEL>::NA; ++j) w.a[j] = coef[(C_A0 + j) * stride + wl];
    w.a10 = coef[(C_A0 + 10) * stride + wl];
    w.a12 = coef[(C_A0 + 12) * stride + wl];
#pragma unroll
    for (int j = 0; j < WalkerRegs<NIFO, MODEL>::NBQ; ++j) w.bq[j] = coef[(C_B0 + j) * stride + wl];
    w.c8 = coef[C_C8 * stride + wl];
    w.k0m = coef[C_K0 * stride + wl] + (SC == BJ_SINCOS_FP64 ? 0.0 : kMagic);
#pragma unroll
    for (int j = 0; j < 6; ++j) w.q[j] = coef[(C_Q2 + j) * stride + wl];
    w.q6l = coef[C_Q6L * stride + wl];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) w.tau[i] = coef[(C_TAU0 + i) * stride + wl];
    if (fid) {
        // focus table (capi.cu: bj_focus): the fiducial phase and delays live in the data weights -- the walker is evaluated
        // relative to them.  fid = {A_0..15, B_0..6, C8, tau_0..}
#pragma unroll
        for (int j = 0; j < WalkerRegs<NIFO, MODEL>::NA; ++j) w.a[j] -= fid[j];
        w.a10 -= fid[10];
        w.a12 -= fid[12];
#pragma unroll
        for (int j = 0; j < WalkerRegs<NIFO, MODEL>::NBQ; ++j) w.bq[j] -= fid[16 + j];
        w.c8 -= fid[23];
#pragma unroll
        for (int i = 0; i < NIFO; ++i) w.tau[i] -= fid[24 + i];
    }

    double zr[NIFO], zi[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) { zr[i] = 0.0; zi[i] = 0.0; }

    const int chunk_id = chunk0 + (int)blockIdx.x;
    const ChunkDesc chunk = chunks[chunk_id];
    const long bin0 = chunk.start;
    const int nb_total = chunk.len;
    const int n_tiles = (nb_total + kTileBins - 1) / kTileBins;
    Spcal64<NIFO> sp;
    if constexpr (SPCAL) {
        sp.f0 = spcal_freqs[chunk.seg];
        sp.inv_width = 1.0 / (spcal_freqs[chunk.seg + 1] - sp.f0);
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            sp.c0r[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg, 0) * stride + wl];
            sp.c0i[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg, 1) * stride + wl];
            sp.dcr[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 0) * stride + wl] - sp.c0r[i];
            sp.dci[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 1) * stride + wl] - sp.c0i[i];
        }
    }

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
        int b = 0;
#pragma unroll 1
        for (; b + 2 <= nb; b += 2) process_bins<NIFO, MODEL, SC, 2, SPCAL>(tile + b * RS, w, sp, zr, zi);
        if (b < nb) process_bins<NIFO, MODEL, SC, 1, SPCAL>(tile + b * RS, w, sp, zr, zi);
        __syncthreads();
        if (threadIdx.x == 0 && t + kStages < n_tiles) {
            fence_proxy_async();
            issue(t + kStages, slot);
        }
    }

    if (live) {
        double* out = partial + (size_t)chunk_id * (2 * NIFO) * stride + wi;
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            out[(2 * i + 0) * stride] = zr[i];
            out[(2 * i + 1) * stride] = zi[i];
        }
    }
}

// ---- the FP64 window kernel ----------------------------------------------------------------------------------------
// The production path evaluates the bins that carry most of the signal power -- where, at high SNR, the terms of <d|h>
// are coherent and FP32 rounding does not average down -- in FP64 (capi.cu: build_layout, the "precise window": a few
// per cent of the table, cut into chunks of 64 records = one TMA tile).  Same structure as fullgrid_kernel (one thread = one walker,
// TMA ring of FP64 records), but built for speed on a uniform grid:
//   * ONE sin/cos per bin (detector 0), from the exactly reduced angle with FP64 polynomials on [-pi/4, pi/4];
//   * the other detectors by FP64 rotation e^{-2 pi i f (tau_i - tau_0)}, advanced bin to bin by a complex multiplication
//     (re-anchored exactly at every 64-bin tile: the recurrence never runs longer than that).
// sin(x), cos(x) for |x| <= pi/4: Taylor polynomials of degree 13 / 12 (truncation error < 1e-14)
__device__ __forceinline__ void sincos_turns_f64(double turns, double& s, double& c) {
    const double r = turns - rint(turns);              // [-1/2, 1/2]
    const double q = rint(4.0 * r);
    const double x = fma(q, -0.25, r) * kTwoPi;        // |x| <= pi/4
    const double x2 = x * x;
    double sp = 1.6059043836821613e-10;                // 1/13!
    sp = fma(sp, x2, -2.5052108385441720e-08);
    sp = fma(sp, x2, 2.7557319223985893e-06);
    sp = fma(sp, x2, -1.9841269841269841e-04);
    sp = fma(sp, x2, 8.3333333333333332e-03);
    sp = fma(sp, x2, -1.6666666666666666e-01);
    sp = fma(sp * x2, x, x);
    double cp = 2.0876756987868100e-09;                // 1/12!
    cp = fma(cp, x2, -2.7557319223985888e-07);
    cp = fma(cp, x2, 2.4801587301587302e-05);
    cp = fma(cp, x2, -1.3888888888888889e-03);
    cp = fma(cp, x2, 4.1666666666666664e-02);
    cp = fma(cp, x2, -0.5);
    cp = fma(cp, x2, 1.0);
    const int qi = (int)q & 3;
    const double s0 = (qi & 1) ? cp : sp, c0 = (qi & 1) ? sp : cp;
    s = (qi & 2) ? -s0 : s0;
    c = (qi == 1 || qi == 2) ? -c0 : c0;
}

template <int NIFO, int MODEL, bool SPCAL>
__global__ void __launch_bounds__(kThreads, 2) fullgrid_window_kernel(const double* __restrict__ rec,
                                                                      const ChunkDesc* __restrict__ chunks,
                                                                      const double* __restrict__ coef, long stride,
                                                                      long n_walkers, double* __restrict__ partial,
                                                                      const double* __restrict__ coef_ex, ExtrasDims dims,
                                                                      const double* __restrict__ spcal_freqs, int chunk0,
                                                                      double df, Sel sel,
                                                                      const __grid_constant__ MultiTable mt) {
    constexpr int RS = Rec<NIFO>::kDoubles;
    constexpr int NR = NIFO > 1 ? NIFO - 1 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kTileBins * RS * 8);

    long n_sel, entry0 = (long)blockIdx.y * kThreads;
    if (mt.n > 0) {
        int lv;
        if (!resolve_tile(mt, kThreads, (int)blockIdx.y, lv, entry0, n_sel)) return;
        const TableRef& T = mt.t[lv];
        if ((int)blockIdx.x >= T.n_precise) return;
        rec = T.rec64;
        chunks = T.chunks;
        chunk0 = 0;
        sel.idx = mt.lists + (size_t)lv * mt.list_stride;
    } else {
        n_sel = sel.idx ? (long)*sel.count : n_walkers;
        if (entry0 >= n_sel) return;
    }
    const long entry = entry0 + threadIdx.x;
    const bool live = entry < n_sel;
    const long wl = sel.idx ? (long)sel.idx[live ? entry : n_sel - 1] : (live ? entry : n_sel - 1);
    const long wi = wl;

    WalkerRegs<NIFO, MODEL> w;
#pragma unroll
    for (int j = 0; j < WalkerRegs<NIFO, MODEL>::NA; ++j) w.a[j] = coef[(C_A0 + j) * stride + wl];
    w.a10 = coef[(C_A0 + 10) * stride + wl];
    w.a12 = coef[(C_A0 + 12) * stride + wl];
#pragma unroll
    for (int j = 0; j < WalkerRegs<NIFO, MODEL>::NBQ; ++j) w.bq[j] = coef[(C_B0 + j) * stride + wl];
    w.c8 = coef[C_C8 * stride + wl];
    w.k0m = coef[C_K0 * stride + wl];
#pragma unroll
    for (int j = 0; j < 6; ++j) w.q[j] = coef[(C_Q2 + j) * stride + wl];
    w.q6l = coef[C_Q6L * stride + wl];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) w.tau[i] = coef[(C_TAU0 + i) * stride + wl];
    // one-bin rotation of detector r + 1 against detector 0: e^{-2 pi i df (tau_i - tau_0)}
    double dtau[NR], r1c[NR], r1s[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        dtau[r] = NIFO > 1 ? w.tau[(r + 1) % NIFO] - w.tau[0] : 0.0;
        const double t = df * dtau[r];
        sincos_turns_f64(t, r1s[r], r1c[r]);
    }

    double zr[NIFO], zi[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) { zr[i] = 0.0; zi[i] = 0.0; }

    const int chunk_id = chunk0 + (int)blockIdx.x;
    const ChunkDesc chunk = chunks[chunk_id];
    const long bin0 = chunk.start;
    const int nb_total = chunk.len;
    const int n_tiles = (nb_total + kTileBins - 1) / kTileBins;
    Spcal64<NIFO> sp;
    if constexpr (SPCAL) {
        sp.f0 = spcal_freqs[chunk.seg];
        sp.inv_width = 1.0 / (spcal_freqs[chunk.seg + 1] - sp.f0);
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            sp.c0r[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg, 0) * stride + wl];
            sp.c0i[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg, 1) * stride + wl];
            sp.dcr[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 0) * stride + wl] - sp.c0r[i];
            sp.dci[i] = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 1) * stride + wl] - sp.c0i[i];
        }
    }

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
        // exact anchors of the rotations at the tile's first bin
        double rc[NR], rs[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double sn, cs;
            sincos_turns_f64(tile[3] * dtau[r], sn, cs);
            rc[r] = cs;
            rs[r] = sn;                                  // R = rc - i rs
        }
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            const double* rb = tile + (size_t)b * RS;
            const double2 h0 = *reinterpret_cast<const double2*>(rb);
            const double2 h1 = *reinterpret_cast<const double2*>(rb + 2);
            const double u = h0.x, L = h0.y, w5 = h1.x, f = h1.y, u2 = u * u;
            double ph;
            if constexpr (MODEL == MODEL_35) {
                double p = fma(w.a12, u2, w.a10);
                p = fma(p, u2 * u, w.a[7]);
#pragma unroll
                for (int k = 6; k >= 2; --k) p = fma(p, u, w.a[k]);
                p = fma(p, u2, w.a[0]);
                const double bl = fma(w.bq[1], u, w.bq[0]);
                ph = fma(w5, p, fma(L, bl, w.k0m));
            } else {
                double p = w.a[15], bl = w.bq[6];
#pragma unroll
                for (int k = 14; k >= 0; --k) p = fma(p, u, w.a[k]);
#pragma unroll
                for (int k = 5; k >= 0; --k) bl = fma(bl, u, w.bq[k]);
                bl = fma(L, w.c8 * (u2 * u), bl);
                ph = fma(w5, p, fma(L, bl, w.k0m));
            }
            double qq = fma(w.q6l, L, w.q[4]);
            qq = fma(w.q[5], u, qq);
#pragma unroll
            for (int k = 3; k >= 0; --k) qq = fma(qq, u, w.q[k]);
            const double Q = fma(u2, qq, 1.0);
            double sn, cs;
            sincos_turns_f64(fma(f, w.tau[0], ph), sn, cs);
            const double pc = Q * cs, ps = Q * sn;       // Q e^{-i theta_0} = pc - i ps
#pragma unroll
            for (int i = 0; i < NIFO; ++i) {
                double cq = pc, sq = ps;
                if (i > 0) {
                    // (pc - i ps)(rc - i rs)
                    cq = pc * rc[i - 1] - ps * rs[i - 1];
                    sq = ps * rc[i - 1] + pc * rs[i - 1];
                }
                if constexpr (SPCAL) {
                    const double tt = fmin(fmax((f - sp.f0) * sp.inv_width, 0.0), 1.0);
                    const double calr = fma(sp.dcr[i], tt, sp.c0r[i]), cali = fma(sp.dci[i], tt, sp.c0i[i]);
                    const double c2 = cq * calr + sq * cali, s2 = sq * calr - cq * cali;
                    cq = c2;
                    sq = s2;
                }
                const double2 d = *reinterpret_cast<const double2*>(rb + 4 + 2 * i);
                zr[i] = fma(d.x, cq, zr[i]);
                zr[i] = fma(d.y, sq, zr[i]);
                zi[i] = fma(d.y, cq, zi[i]);
                zi[i] = fma(-d.x, sq, zi[i]);
            }
            // advance the rotations by one bin
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double nc = rc[r] * r1c[r] - rs[r] * r1s[r];
                const double ns = rs[r] * r1c[r] + rc[r] * r1s[r];
                rc[r] = nc;
                rs[r] = ns;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && t + kStages < n_tiles) {
            fence_proxy_async();
            issue(t + kStages, slot);
        }
    }

    if (live) {
        double* out = partial + (size_t)chunk_id * (2 * NIFO) * stride + wi;
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            out[(2 * i + 0) * stride] = zr[i];
            out[(2 * i + 1) * stride] = zi[i];
        }
    }
}

// ---- the production kernel: FP64 phase, FP32 amplitude / phasor / multiply-accumulate ---------------
// Cost model measured on B200 (scripts/ubench.cu): what limits this kernel is operand-read bandwidth, not a
// math pipe -- a DFMA with three distinct register pairs issues every 3 cycles, F2F.F64.F32 runs at XU rate,
// FP32 and FP64 FMAs do not overlap for free.  Hence: everything that does not NEED 53 bits is FP32:
//   * phase in turns: FP64 Horner + magic constant (exact binary angle in the low word)      [must be FP64]
//   * amplitude series Q: FP32 Horner, two bins per packed FFMA2
//   * phasor: calibrated MUFU or packed polynomial, scaled by Q in FP32
//   * <d|h> terms: FP32 records (dr, di, di, -dr), two packed FFMA2 per detector into an FP32 block
//     accumulator that is flushed into the FP64 accumulators every kBlockBins bins (rounding noise of the
//     short FP32 partial sums is below the phasor noise; analysis in DESIGN.md section 5)
// Record (bytes): [u, L, w5, f : f64 x4][pair slot : f32 x2, pad 8][(dr, di, di, -dr) : f32 x4] x NIFO
// Bins are consumed in (even, odd) pairs; the pair slot of the even record holds (uf_even, uf_odd) and that of
// the odd record (Lf_even, Lf_odd), so each LDS.64 delivers a ready-packed FFMA2 operand.
template <int NIFO> struct RecM {
    static constexpr int kBytes = 48 + 16 * NIFO;
    static constexpr int kDoubles = kBytes / 8;
};
#ifndef BJ_PAIR_UNROLL
#define BJ_PAIR_UNROLL 4
#endif
#ifndef BJ_MIN_CTAS
#define BJ_MIN_CTAS 3
#endif
#ifndef BJ_BLOCK_BINS
#define BJ_BLOCK_BINS 8
#endif
constexpr int kPairUnroll = BJ_PAIR_UNROLL;
constexpr int kBlockBins = BJ_BLOCK_BINS;   // FP32 partial sums are flushed to FP64 every kBlockBins bins

template <int NIFO, int MODEL>
struct WalkerRegsM {
    static constexpr int NA = (MODEL == MODEL_35) ? 8 : 16;
    static constexpr int NBQ = (MODEL == MODEL_35) ? 2 : 7;
    double a[NA], a10, a12, bq[NBQ], c8, k0m, tau[NIFO];
    float q[6], q6l;
};

// amplitude series Q(u, L) in FP64 with the exact per-walker coefficients
__device__ __forceinline__ double amp_series_f64(double u, double L, const double* __restrict__ coef, long stride,
                                                 long w) {
    double t = fma(coef[C_Q6L * stride + w], L, coef[(C_Q2 + 4) * stride + w]);
    t = fma(coef[(C_Q2 + 5) * stride + w], u, t);
    t = fma(t, u, coef[(C_Q2 + 3) * stride + w]);
    t = fma(t, u, coef[(C_Q2 + 2) * stride + w]);
    t = fma(t, u, coef[(C_Q2 + 1) * stride + w]);
    t = fma(t, u, coef[(C_Q2 + 0) * stride + w]);
    return fma(u * u, t, 1.0);
}

// Systematic factor of the FP32 amplitude series at (u, L): the series with the exact coefficients over the series with
// the float-rounded coefficients q[], BOTH evaluated in FP64 (no rounding noise enters).  It is ~1 + 3e-8 x (condition
// number of the series) and varies smoothly with frequency, so it rides on the FP32 -> FP64 flush of the bins around
// (u, L).  Next to a zero of the series -- the 3.5PN amplitude changes sign near v ~ 1, i.e. at the top of the band
// for heavy binaries, where bajes applies no ISCO cut -- the ratio is meaningless (and the bins carry no weight):
// it is then dropped rather than applied to the neighbouring bins.
__device__ __forceinline__ double amp_series_correction(const float (&q)[6], float q6l, double u, double L,
                                                        const double* __restrict__ coef, long stride, long w) {
    const double q_exact = amp_series_f64(u, L, coef, stride, w);
    double t = fma((double)q6l, L, (double)q[4]);
    t = fma((double)q[5], u, t);
    t = fma(t, u, (double)q[3]);
    t = fma(t, u, (double)q[2]);
    t = fma(t, u, (double)q[1]);
    t = fma(t, u, (double)q[0]);
    // 1 + (exact - rounded) / rounded: the difference is ~1e-7, so an FP32 reciprocal is plenty for the quotient
    const double q_rounded = fma(u * u, t, 1.0);
    const double c = (double)((float)(q_exact - q_rounded) * __frcp_rn((float)q_rounded));
    return fabs(c) < 1e-4 ? 1.0 + c : 1.0;
}

// calibration envelope of the chunk's segment, FP32 packed (re, im): cal(t) = c0 + dc * t, t in the record
template <int NIFO>
struct Spcal32 {
    unsigned long long c0[NIFO], dc[NIFO];
};

template <int NIFO, int MODEL, int SC, bool ROBUST, bool SPCAL>
__device__ __forceinline__ void process_pair_mixed(const unsigned char* __restrict__ rec,
                                                   const WalkerRegsM<NIFO, MODEL>& w, const Spcal32<NIFO>& sp,
                                                   unsigned long long (&acc)[NIFO]) {
    constexpr int RB = RecM<NIFO>::kBytes;
    double u[2], L[2], w5[2], f[2], u2[2], ph[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double2* r = reinterpret_cast<const double2*>(rec + j * RB);
        const double2 h0 = r[0], h1 = r[1];
        u[j] = h0.x; L[j] = h0.y; w5[j] = h1.x; f[j] = h1.y;
        u2[j] = u[j] * u[j];
    }
    const unsigned long long up = *reinterpret_cast<const unsigned long long*>(rec + 32);        // (uf0, uf1)
    const unsigned long long lp = *reinterpret_cast<const unsigned long long*>(rec + RB + 32);   // (Lf0, Lf1)
    float tcal[2] = {0.0f, 0.0f};
    if constexpr (SPCAL) {
        const float2 tt2 = *reinterpret_cast<const float2*>(rec + 40);                           // (t0, t1)
        tcal[0] = tt2.x;
        tcal[1] = tt2.y;
    }
    // ---- phase in turns (+ magic), FP64 ----
    if constexpr (MODEL == MODEL_35) {
        double p[2], bl[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) p[j] = fma(w.a12, u2[j], w.a10);
#pragma unroll
        for (int j = 0; j < 2; ++j) p[j] = fma(p[j], u2[j] * u[j], w.a[7]);
#pragma unroll
        for (int k = 6; k >= 2; --k) {
#pragma unroll
            for (int j = 0; j < 2; ++j) p[j] = fma(p[j], u[j], w.a[k]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            p[j] = fma(p[j], u2[j], w.a[0]);
            bl[j] = fma(w.bq[1], u[j], w.bq[0]);
            ph[j] = fma(w5[j], p[j], fma(L[j], bl[j], w.k0m));
        }
    } else {
        double p[2], bl[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) { p[j] = w.a[15]; bl[j] = w.bq[6]; }
#pragma unroll
        for (int k = 14; k >= 0; --k) {
#pragma unroll
            for (int j = 0; j < 2; ++j) p[j] = fma(p[j], u[j], w.a[k]);
        }
#pragma unroll
        for (int k = 5; k >= 0; --k) {
#pragma unroll
            for (int j = 0; j < 2; ++j) bl[j] = fma(bl[j], u[j], w.bq[k]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            bl[j] = fma(L[j], w.c8 * (u2[j] * u[j]), bl[j]);
            ph[j] = fma(w5[j], p[j], fma(L[j], bl[j], w.k0m));
        }
    }
    // ---- amplitude series, FP32, both bins in one packed Horner ----
    float Qf[2];
    {
        unsigned long long t = fma2(pack2(w.q6l, w.q6l), lp, pack2(w.q[4], w.q[4]));
        t = fma2(pack2(w.q[5], w.q[5]), up, t);
        t = fma2(t, up, pack2(w.q[3], w.q[3]));
        t = fma2(t, up, pack2(w.q[2], w.q[2]));
        t = fma2(t, up, pack2(w.q[1], w.q[1]));
        t = fma2(t, up, pack2(w.q[0], w.q[0]));
        t = fma2(mul2(up, up), t, pack2(1.0f, 1.0f));
        unpack2(t, Qf[0], Qf[1]);
    }
    // ---- per detector: ramp, phasor, packed complex MAC ----
#pragma unroll
    for (int i = 0; i < NIFO; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double tt = fma(f[j], w.tau[i], ph[j]);
            float cq, sq;
            scaled_phasor_f32<SC>(ROBUST ? robust_angle(tt) : __double2loint(tt), Qf[j], cq, sq);
            if constexpr (SPCAL) {
                // h *= cal  <=>  (cq - i sq) *= (calr + i cali)      (detector.py:517-519)
                float calr, cali;
                unpack2(fma2(sp.dc[i], pack2(tcal[j], tcal[j]), sp.c0[i]), calr, cali);
                const float c2 = fmaf(sq, cali, cq * calr), s2 = fmaf(-cq, cali, sq * calr);
                cq = c2;
                sq = s2;
            }
            const ulonglong2 d = *reinterpret_cast<const ulonglong2*>(rec + j * RB + 48 + 16 * i);
            // (dr + i di) Q (c - i s) = (dr, di) cq + (di, -dr) sq
            acc[i] = fma2(d.x, pack2(cq, cq), acc[i]);
            acc[i] = fma2(d.y, pack2(sq, sq), acc[i]);
        }
    }
}

// ROBUST = false: the production kernel.  ROBUST = true: the fix-up pass, launched right after it over the same grid;
// a CTA exits at once unless the prologue flagged one of its walkers as "wide phase" (|turns| may exceed 2^19, the
// range of the binary-angle trick), otherwise it recomputes the tile with the exact reduction t - rint(t) and
// overwrites the tile's partial sums.
template <int NIFO, int MODEL, int SC, bool ROBUST, bool SPCAL>
__global__ void __launch_bounds__(kThreads, BJ_MIN_CTAS) fullgrid_mixed_kernel(const unsigned char* __restrict__ rec,
                                                                       const ChunkDesc* __restrict__ chunks,
                                                                       const double* __restrict__ coef, long stride,
                                                                       long n_walkers, double* __restrict__ partial,
                                                                       const double* __restrict__ coef_ex,
                                                                       ExtrasDims dims, int chunk0) {
    constexpr int RB = RecM<NIFO>::kBytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stages = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kTileBins * RB);

    const long wi = (long)blockIdx.y * kThreads + threadIdx.x;
    const long wl = wi < n_walkers ? wi : n_walkers - 1;
    // ROBUST: the walkers this pass replaces -- those the prologue flagged wide-phase
    bool selected = true;
    if constexpr (ROBUST) {
        selected = coef[C_VALID * stride + wl] == 2.0;
        if (!__syncthreads_or(selected)) return;
    }

    WalkerRegsM<NIFO, MODEL> w;
#pragma unroll
    for (int j = 0; j < WalkerRegsM<NIFO, MODEL>::NA; ++j) w.a[j] = coef[(C_A0 + j) * stride + wl];
    w.a10 = coef[(C_A0 + 10) * stride + wl];
    w.a12 = coef[(C_A0 + 12) * stride + wl];
#pragma unroll
    for (int j = 0; j < WalkerRegsM<NIFO, MODEL>::NBQ; ++j) w.bq[j] = coef[(C_B0 + j) * stride + wl];
    w.c8 = coef[C_C8 * stride + wl];
    w.k0m = coef[C_K0 * stride + wl] + (ROBUST ? 0.0 : kMagic);
#pragma unroll
    for (int j = 0; j < 6; ++j) w.q[j] = (float)coef[(C_Q2 + j) * stride + wl];
    w.q6l = (float)coef[C_Q6L * stride + wl];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) w.tau[i] = coef[(C_TAU0 + i) * stride + wl];

    double zr[NIFO], zi[NIFO];
    unsigned long long acc[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) { zr[i] = 0.0; zi[i] = 0.0; acc[i] = 0ull; }

    // every chunk starts on an even record and holds an even number of records (cells are padded with zero-weight
    // records), so bins can be consumed in (even, odd) pairs
    const int chunk_id = chunk0 + (int)blockIdx.x;
    const ChunkDesc chunk = chunks[chunk_id];
    const long bin0 = chunk.start;
    const int nb_total = chunk.len;
    const int n_tiles = (nb_total + kTileBins - 1) / kTileBins;

    // calibration envelope of this chunk's segment: FP32 node value + slope per detector, and the exact/rounded
    // ratio at the chunk's first bin (noise-free, both sides in FP64) that rides on the flush
    Spcal32<NIFO> sp;
    double ccr[NIFO], cci[NIFO];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) { ccr[i] = 1.0; cci[i] = 0.0; sp.c0[i] = 0ull; sp.dc[i] = 0ull; }
    if constexpr (SPCAL) {
        const double t0 = (double)*reinterpret_cast<const float*>(rec + bin0 * RB + 40);
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            const double c0r = coef_ex[ex_cal_slot(dims, i, chunk.seg, 0) * stride + wl];
            const double c0i = coef_ex[ex_cal_slot(dims, i, chunk.seg, 1) * stride + wl];
            const double dr = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 0) * stride + wl] - c0r;
            const double di = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 1) * stride + wl] - c0i;
            const float c0rf = (float)c0r, c0if = (float)c0i, drf = (float)dr, dif = (float)di;
            sp.c0[i] = pack2(c0rf, c0if);
            sp.dc[i] = pack2(drf, dif);
            const double er = fma(dr, t0, c0r), ei = fma(di, t0, c0i);                       // exact
            const double rr = fma((double)drf, t0, (double)c0rf), ri = fma((double)dif, t0, (double)c0if);
            const double den = rr * rr + ri * ri;
            ccr[i] = (er * rr + ei * ri) / den;                                               // exact / rounded
            cci[i] = (ei * rr - er * ri) / den;
        }
    }

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int slot) {
        const int nb = min(kTileBins, nb_total - tile * kTileBins);
        const uint32_t bytes = (uint32_t)nb * RB;
        mbar_expect_tx(&full[slot], bytes);
        tma_bulk_g2s(stages + (size_t)slot * kTileBins * RB, rec + (bin0 + (long)tile * kTileBins) * RB, bytes,
                     &full[slot]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages && s < n_tiles; ++s) issue(s, s);
    }

    // The FP32 amplitude series carries a SYSTEMATIC relative error of a few 1e-8 (its per-walker coefficients
    // are rounded once, for all bins), and a systematic factor on every term is a factor on <d|h> ~ 1e4.  It
    // varies smoothly with frequency, so once per tile the exact FP64 series is evaluated at the tile's first bin
    // and the ratio rides on the flush of the FP32 block sums (the flush add becomes an FMA: free).
    double corr = 1.0;
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            float re, im;
            unpack2(acc[i], re, im);
            if constexpr (SPCAL) {
                const double fr = corr * ccr[i], fi = corr * cci[i];
                zr[i] = fma((double)re, fr, fma(-(double)im, fi, zr[i]));
                zi[i] = fma((double)re, fi, fma((double)im, fr, zi[i]));
            } else {
                zr[i] = fma((double)re, corr, zr[i]);
                zi[i] = fma((double)im, corr, zi[i]);
            }
            acc[i] = 0ull;
        }
    };

    for (int t = 0; t < n_tiles; ++t) {
        const int slot = t % kStages;
        mbar_wait(&full[slot], (uint32_t)((t / kStages) & 1));
        const unsigned char* tile = stages + (size_t)slot * kTileBins * RB;
        const int nb = min(kTileBins, nb_total - t * kTileBins);
        {
            const double2 h0 = *reinterpret_cast<const double2*>(tile);          // the tile's first bin
            corr = amp_series_correction(w.q, w.q6l, h0.x, h0.y, coef, stride, wl);
        }
        const int n_blocks = nb / kBlockBins;
#pragma unroll 1
        for (int blk = 0; blk < n_blocks; ++blk) {
            const unsigned char* base = tile + (size_t)blk * kBlockBins * RB;
#pragma unroll (ROBUST ? 1 : kPairUnroll)
            for (int b = 0; b < kBlockBins; b += 2) process_pair_mixed<NIFO, MODEL, SC, ROBUST, SPCAL>(base + b * RB, w, sp, acc);
            flush();
        }
        {
            const int done = n_blocks * kBlockBins;
#pragma unroll 1
            for (int b = done; b + 2 <= nb; b += 2) process_pair_mixed<NIFO, MODEL, SC, ROBUST, SPCAL>(tile + (size_t)b * RB, w, sp, acc);
            if (done < nb) flush();
        }
        __syncthreads();
        if (threadIdx.x == 0 && t + kStages < n_tiles) {
            fence_proxy_async();
            issue(t + kStages, slot);
        }
    }

    // the fix-up pass replaces the sums of the flagged walkers only: every walker is written by exactly one path, so
    // its bits do not depend on which other walkers share its tile
    if (wi < n_walkers && (!ROBUST || selected)) {
        double* out = partial + (size_t)chunk_id * (2 * NIFO) * stride + wi;
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            out[(2 * i + 0) * stride] = zr[i];
            out[(2 * i + 1) * stride] = zi[i];
        }
    }
}

// ---- the tensor-core kernel: the phase as an FP64 GEMM (the default production path) ------------------------------
// The Horner kernels above are bound neither by a pipe nor by memory but by instruction issue / register-operand
// bandwidth: B200 delivers one 64-bit register operand per lane and cycle (scripts/ubench_operands.cu: DFMA and FFMA2
// with three distinct operands issue every 3 cycles, with one operand from the reuse cache every 2), so the ~56
// instructions per walker and bin cost ~97 cycles per warp (profiles/r01_ablation.txt: removing every MUFU gains 7 %,
// removing the phase gains 25 %).  The phase is LINEAR in the per-walker coefficients:
//     Phi_c[w, k] = sum_n A[w, n] B[n, k],   B[n, k] = w5 u^j, L u^j, L^2 u^3, 1   (static per bin)
// i.e. a [walkers x K] x [K x bins] GEMM with K = 12 (3.5PN) or 28 (5.5PN + tides), which the FP64 tensor cores
// (DMMA.8x8x4, mma.sync.m8n8k4.f64) evaluate at the FLOP rate of DFMA with 1/8 of the instructions
// (scripts/ubench_dmma.cu; tcgen05 has no FP64 kind, mma.sync IS the FP64 tensor path on sm_100a).
//
// Layout: a warp owns 8 walkers x 32 bins per step.  In the m8n8k4 accumulator fragment lane (g = lane / 4,
// t = lane % 4) holds walker g and bins {8m + 2t, 8m + 2t + 1}, m = 0..3: ONE walker per lane (coefficients and
// accumulators as in the kernels above), 8 bins per step, an (even, odd) pair per n-tile for the packed FP32
// amplitude.  The four lanes of a walker are reduced by two shuffles at the end of the kernel.  A CTA = 8 warps = 64
// walkers streams the chunk through a 3-stage TMA ring of 256 bins.
// Two static tables (SoA, so that both access patterns are free of shared-memory bank conflicts):
//   basis[bin][K]  f64      stride 96 B (K = 12) / 224 B (K = 28): the B-fragment load of a half warp (4 bins x 4
//                           doubles) covers all 32 banks exactly once;
//   data[bin]      48 B     [f : f64][pair slot : f32 x2][(dr, di) : f32 x2] x 3 [t : f32] 4 B padding: the four bins
//                           a quarter warp reads (2 records = 24 words apart) fall into disjoint banks.
// Every 128-bit load serves 8 walkers here instead of 32, and shared-memory wavefronts are what this kernel runs out
// of: with the 16-byte (dr, di, di, -dr) entries of the Horner kernels it sat at 87 % of the LSU's cycles (5.04 ms),
// hence the 8-byte entries and the two scalar FFMAs with a negated operand in the MAC.  (A hybrid that keeps one
// walker per thread and transposes the accumulator fragments through shared memory was measured too: 5.49 ms,
// against 5.59 ms for the Horner kernel; 4096 walkers x 521 729 bins x 3 detectors.)
#ifndef BJ_TILE_BINS
#define BJ_TILE_BINS 256
#endif
#ifndef BJ_TILE_STAGES
#define BJ_TILE_STAGES 3
#endif
constexpr int kTileBinsT = BJ_TILE_BINS;      // bins per TMA stage (fewer CTA-wide barriers than 128: -5 %)
constexpr int kStagesT = BJ_TILE_STAGES;
template <int MODEL> struct TileRec {
    static constexpr int K = (MODEL == MODEL_35) ? 12 : 28;
    static constexpr int kSteps = K / 4;
    static constexpr int kBasisBytes = K * 8;
    static constexpr int kDataBytes = 48;
    static constexpr int kF = 0;
    static constexpr int kPair = 8;
    static constexpr int kData = 16;
    static constexpr int kT = 40;            // f32: calibration interpolation fraction t(f) of the bin (0 without spcal)
    // bins per TMA stage: as many as two resident CTAs allow (fewer CTA-wide barriers: 256 instead of 128 is -5 %)
    static constexpr int kTileBins = (MODEL == MODEL_35) ? kTileBinsT : kTileBinsT / 2;
};
static_assert(TileRec<MODEL_35>::kData + 8 * BJ_MAX_IFO <= TileRec<MODEL_35>::kT, "record too small");

#ifndef BJ_TILE_THREADS
#define BJ_TILE_THREADS 256
#endif
constexpr int kTileThreads = BJ_TILE_THREADS;
constexpr int kTileWalkers = kTileThreads / 4;     // 8 per warp
constexpr int kStepBins = 32;

// coefficient slot of basis function n (-1: zero padding); the host builds B in the same order (capi.cu)
template <int MODEL>
__host__ __device__ constexpr int tile_basis_slot(int n) {
    if (MODEL == MODEL_35) {
        // w5 u^{0,2,3,4,5,6,7,10,12}, L, L u, 1
        return n == 0 ? C_A0 : n <= 6 ? C_A0 + n + 1 : n == 7 ? C_A0 + 10 : n == 8 ? C_A0 + 12
             : n == 9 ? C_B0 : n == 10 ? C_B0 + 1 : C_K0;
    }
    // w5 u^0..15, L u^0..6, L^2 u^3, 1, 3 x padding
    return n < 16 ? C_A0 + n : n < 23 ? C_B0 + (n - 16) : n == 23 ? C_C8 : n == 24 ? C_K0 : -1;
}

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int NIFO>
struct TileWalker {
    double tau[NIFO];
    float q[6], q6l;
    uint32_t d1[NIFO], d7[NIFO];   // uniform grids: one-bin step of the time-delay ramp, df tau_i, in 2^-32 turns (and x 7); binary angles wrap mod 2^32
};

// one step: 8 walkers x 32 bins per warp.  `limit` = number of valid bins of the step (32 except at the ragged end
// of a chunk; always even); GUARD = false compiles the test away.
//
// IRAMP (uniform grids, i.e. every FFT grid bajes produces): the time-delay ramp f_k tau_i is advanced in 32-bit binary
// angles from one FP64 anchor per step and detector, taken at the CENTRE of the step: the rounding of the increment
// (<= 2^-33 turns per bin, <= 1.2e-8 rad at the step's ends) is antisymmetric about the anchor and leaves no phase
// bias.  This replaces one DFMA (which competes with the DMMAs for the FP64 pipe) and the load of f per bin and
// detector by two integer adds.
//
// SPCAL (calibration envelopes, detector.py:517-519): inside a chunk cal_i(f) = c0_i + (c1_i - c0_i) t(f) with constant
// node values, so  sum_k cal_i X_k = c0_i sum_k X_k + (c1_i - c0_i) sum_k t_k X_k : a second accumulator fed with
// Q t instead of Q (one FMUL per bin + one FFMA2 per detector and bin); the node values enter once per chunk, in
// FP64, when the kernel stores its partial sums.
template <int NIFO, int MODEL, int SC, bool GUARD, bool IRAMP, bool SPCAL>
__device__ __forceinline__ void tile_step(const unsigned char* __restrict__ basis, const unsigned char* __restrict__ base,
                                          const double (&afrag)[TileRec<MODEL>::kSteps], const TileWalker<NIFO>& w,
                                          int g, int t, int limit, double df16, unsigned long long (&acc)[NIFO],
                                          unsigned long long (&acc1)[NIFO]) {
    using R = TileRec<MODEL>;
    constexpr int RB = R::kDataBytes;
    // ---- phase: D[m][h] = turns (+ magic) of walker g at bin 8m + 2t + h ----
    double D[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) { D[m][0] = 0.0; D[m][1] = 0.0; }
#pragma unroll
    for (int s = 0; s < R::kSteps; ++s) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const double b = *reinterpret_cast<const double*>(basis + (size_t)(8 * m + g) * R::kBasisBytes + (4 * s + t) * 8);
            dmma884(D[m], afrag[s], b);
        }
    }
    // ---- amplitude: packed FP32 Horner over the lane's four (even, odd) pairs ----
    float Qf[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const unsigned char* re = base + (size_t)(8 * m + 2 * t) * RB;
        const unsigned long long up = *reinterpret_cast<const unsigned long long*>(re + R::kPair);        // (uf0, uf1)
        const unsigned long long lp = *reinterpret_cast<const unsigned long long*>(re + RB + R::kPair);   // (Lf0, Lf1)
        unsigned long long q = fma2(pack2(w.q6l, w.q6l), lp, pack2(w.q[4], w.q[4]));
        q = fma2(pack2(w.q[5], w.q[5]), up, q);
        q = fma2(q, up, pack2(w.q[3], w.q[3]));
        q = fma2(q, up, pack2(w.q[2], w.q[2]));
        q = fma2(q, up, pack2(w.q[1], w.q[1]));
        q = fma2(q, up, pack2(w.q[0], w.q[0]));
        q = fma2(mul2(up, up), q, pack2(1.0f, 1.0f));
        unpack2(q, Qf[m][0], Qf[m][1]);
    }
    // ---- per bin and detector: ramp, phasor, packed complex MAC ----
    uint32_t e[NIFO];
    if constexpr (IRAMP) {
        const double fc = *reinterpret_cast<const double*>(base + R::kF) + df16;     // bin 16 of the step
#pragma unroll
        for (int i = 0; i < NIFO; ++i) e[i] = (uint32_t)__double2loint(fma(fc, w.tau[i], kMagic)) + (uint32_t)(2 * t - 16) * w.d1[i];
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (GUARD && 8 * m + 2 * t >= limit) continue;        // ring memory beyond the chunk is never used
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned char* rb = base + (size_t)(8 * m + 2 * t + h) * RB;
            double f = 0.0;
            if constexpr (!IRAMP) f = *reinterpret_cast<const double*>(rb + R::kF);
            float qt = 0.0f;
            if constexpr (SPCAL) qt = Qf[m][h] * *reinterpret_cast<const float*>(rb + R::kT);
            unsigned long long dd[NIFO];
            if constexpr (NIFO >= 2) {
                const ulonglong2 d01 = *reinterpret_cast<const ulonglong2*>(rb + R::kData);
                dd[0] = d01.x;
                dd[1] = d01.y;
                if constexpr (NIFO == 3) dd[2] = *reinterpret_cast<const unsigned long long*>(rb + R::kData + 16);
            } else {
                dd[0] = *reinterpret_cast<const unsigned long long*>(rb + R::kData);
            }
#pragma unroll
            for (int i = 0; i < NIFO; ++i) {
                float cq, sq;
                int32_t b;
                if constexpr (IRAMP) {
                    b = (int32_t)((uint32_t)__double2loint(D[m][h]) + e[i]);
                    e[i] += h == 0 ? w.d1[i] : w.d7[i];           // next bin of this lane: + 1, then + 7
                } else {
                    b = __double2loint(fma(f, w.tau[i], D[m][h]));
                }
                float dr, di, ar, ai;
                unpack2(dd[i], dr, di);
                if constexpr (SC == BJ_SINCOS_MUFU) {
                    // acc += Q [(dr, di) c + (di, -dr) s]: Q enters once, on the accumulate
                    scaled_phasor_f32<SC>(b, 1.0f, cq, sq);
                    unpack2(mul2(dd[i], pack2(cq, cq)), ar, ai);
                    const unsigned long long term = pack2(fmaf(di, sq, ar), fmaf(-dr, sq, ai));
                    acc[i] = fma2(term, pack2(Qf[m][h], Qf[m][h]), acc[i]);
                    if constexpr (SPCAL) acc1[i] = fma2(term, pack2(qt, qt), acc1[i]);
                } else if constexpr (SPCAL) {
                    scaled_phasor_f32<SC>(b, 1.0f, cq, sq);          // (the half-turn sign rides on the unit amplitude)
                    unpack2(mul2(dd[i], pack2(cq, cq)), ar, ai);
                    const unsigned long long term = pack2(fmaf(di, sq, ar), fmaf(-dr, sq, ai));
                    acc[i] = fma2(term, pack2(Qf[m][h], Qf[m][h]), acc[i]);
                    acc1[i] = fma2(term, pack2(qt, qt), acc1[i]);
                } else {
                    // (dr + i di) Q (c - i s) = (dr, di) cq + (di, -dr) sq
                    scaled_phasor_f32<SC>(b, Qf[m][h], cq, sq);
                    unpack2(fma2(dd[i], pack2(cq, cq), acc[i]), ar, ai);
                    acc[i] = pack2(fmaf(di, sq, ar), fmaf(-dr, sq, ai));
                }
            }
        }
    }
}

// ---- ROT: one phasor per bin, the other detectors by rotation ------------------------------------------------
// The phase of detector i differs from detector 0's by 2 pi f (tau_i - tau_0) only, and on a uniform grid the bins of a
// lane sit at fixed offsets j - 12.5, j in {0, 1, 8, 9, 16, 17, 24, 25}, from the lane's centre f_c:
//     e^{-i theta_i,k} = e^{-i theta_0,k} . e^{-i (j - 12.5) eps_i} . E_i,      eps_i = 2 pi df (tau_i - tau_0),
//     E_i = e^{-2 pi i f_c (tau_i - tau_0)}.
// The offsets come in +- pairs {3.5, 4.5, 11.5, 12.5}: four (cos - 1, sin) constants per detector and walker (prologue,
// FP64, rounded once; cos enters as 1 + rho so that its rounding error is ~1e-12).  Per bin, ONE phasor P = Q e^{-i theta_0}
// (MUFU or polynomial) is rotated to the other detectors with exact-angle FP32 rotations (no recurrence: every term
// sees two roundings, relative to itself), and E_i -- advanced per step in FP64 by e^{-i 32 eps_i} from an FP64 anchor
// per chunk -- multiplies the 8-bin FP32 sums on their way into the FP64 accumulators.  2 instead of 2 NIFO MUFU per bin;
// with three detectors the two rotated ones are packed into FFMA2 lanes (data table: (dr1, dr2), (di1, di2)).
template <int NIFO>
struct TileRot {
    static constexpr int NR = NIFO > 1 ? NIFO - 1 : 1;
    unsigned long long rho[4], sn[4];      // NR == 2: (detector 1, detector 2) packed; NR == 1: low half
    double er[NR], ei[NR];                 // E_i of the current step
    double rr[NR], ri[NR];                 // e^{-i 32 eps_i}
};

// accumulators: acc0 = (re, im) of detector 0; NR == 2: accR = (re1, re2), accI = (im1, im2); NR == 1: accR = (re1, im1).
// The *_t twins take the calibration-weighted terms (SPCAL).
template <int NIFO, int MODEL, int SC, bool GUARD, bool SPCAL>
__device__ __forceinline__ void tile_step_rot(const unsigned char* __restrict__ basis, const unsigned char* __restrict__ base,
                                              const double (&afrag)[TileRec<MODEL>::kSteps], const TileWalker<NIFO>& w,
                                              const TileRot<NIFO>& rot, int g, int t, int limit, double df16,
                                              unsigned long long& acc0, unsigned long long& accR, unsigned long long& accI,
                                              unsigned long long& acc0_t, unsigned long long& accR_t,
                                              unsigned long long& accI_t) {
    using R = TileRec<MODEL>;
    constexpr int RB = R::kDataBytes;
    constexpr int NR = NIFO - 1;
    static_assert(NIFO >= 2 && NIFO <= 3, "rotation path: two or three detectors");
    double D[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) { D[m][0] = 0.0; D[m][1] = 0.0; }
#pragma unroll
    for (int s = 0; s < R::kSteps; ++s) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const double b = *reinterpret_cast<const double*>(basis + (size_t)(8 * m + g) * R::kBasisBytes + (4 * s + t) * 8);
            dmma884(D[m], afrag[s], b);
        }
    }
    float Qf[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const unsigned char* re = base + (size_t)(8 * m + 2 * t) * RB;
        const unsigned long long up = *reinterpret_cast<const unsigned long long*>(re + R::kPair);
        const unsigned long long lp = *reinterpret_cast<const unsigned long long*>(re + RB + R::kPair);
        unsigned long long q = fma2(pack2(w.q6l, w.q6l), lp, pack2(w.q[4], w.q[4]));
        q = fma2(pack2(w.q[5], w.q[5]), up, q);
        q = fma2(q, up, pack2(w.q[3], w.q[3]));
        q = fma2(q, up, pack2(w.q[2], w.q[2]));
        q = fma2(q, up, pack2(w.q[1], w.q[1]));
        q = fma2(q, up, pack2(w.q[0], w.q[0]));
        q = fma2(mul2(up, up), q, pack2(1.0f, 1.0f));
        unpack2(q, Qf[m][0], Qf[m][1]);
    }
    // detector 0's time-delay ramp in binary angles, anchored at the step's centre (see tile_step)
    const double fc = *reinterpret_cast<const double*>(base + R::kF) + df16;
    uint32_t e0 = (uint32_t)__double2loint(fma(fc, w.tau[0], kMagic)) + (uint32_t)(2 * t - 16) * w.d1[0];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (GUARD && 8 * m + 2 * t >= limit) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned char* rb = base + (size_t)(8 * m + 2 * t + h) * RB;
            // offset of this bin from the lane's centre: j - 12.5 with j = 8 m + h
            constexpr int kIdx[4][2] = {{3, 2}, {1, 0}, {0, 1}, {2, 3}};
            const int idx = kIdx[m][h];
            const bool pos = m >= 2;                             // sign of the offset
            const int32_t b = (int32_t)((uint32_t)__double2loint(D[m][h]) + e0);
            e0 += h == 0 ? w.d1[0] : w.d7[0];
            float qt = 0.0f;
            if constexpr (SPCAL) qt = Qf[m][h] * *reinterpret_cast<const float*>(rb + R::kT);
            // the one phasor of this bin: (pc, ps) = Q (cos, sin) theta_0   (unit amplitude with SPCAL: Q and Q t enter on
            // the accumulate)
            float pc, ps;
            scaled_phasor_f32<SC>(b, SPCAL ? 1.0f : Qf[m][h], pc, ps);
            // ---- detector 0 ----
            {
                const unsigned long long d0 = *reinterpret_cast<const unsigned long long*>(rb + R::kData);
                float dr, di, ar, ai;
                unpack2(d0, dr, di);
                if constexpr (SPCAL) {
                    unpack2(mul2(d0, pack2(pc, pc)), ar, ai);
                    const unsigned long long term = pack2(fmaf(di, ps, ar), fmaf(-dr, ps, ai));
                    acc0 = fma2(term, pack2(Qf[m][h], Qf[m][h]), acc0);
                    acc0_t = fma2(term, pack2(qt, qt), acc0_t);
                } else {
                    unpack2(fma2(d0, pack2(pc, pc), acc0), ar, ai);
                    acc0 = pack2(fmaf(di, ps, ar), fmaf(-dr, ps, ai));
                }
            }
            // ---- rotated detectors ----
            if constexpr (NR == 2) {
                // (c1, c2) = pc (1 + rho) -+ ps s ;  (s1, s2) = ps (1 + rho) +- pc s
                unsigned long long cv = fma2(pack2(pc, pc), rot.rho[idx], pack2(pc, pc));
                unsigned long long sv = fma2(pack2(ps, ps), rot.rho[idx], pack2(ps, ps));
                if (pos) {
                    cv = fma2(pack2(-ps, -ps), rot.sn[idx], cv);
                    sv = fma2(pack2(pc, pc), rot.sn[idx], sv);
                } else {
                    cv = fma2(pack2(ps, ps), rot.sn[idx], cv);
                    sv = fma2(pack2(-pc, -pc), rot.sn[idx], sv);
                }
                const unsigned long long drv = *reinterpret_cast<const unsigned long long*>(rb + R::kData + 8);   // (dr1, dr2)
                const unsigned long long div = *reinterpret_cast<const unsigned long long*>(rb + R::kData + 16);  // (di1, di2)
                if constexpr (SPCAL) {
                    const unsigned long long tr = fma2(div, sv, mul2(drv, cv));
                    const unsigned long long ti = fma2(neg2(drv), sv, mul2(div, cv));
                    accR = fma2(tr, pack2(Qf[m][h], Qf[m][h]), accR);
                    accI = fma2(ti, pack2(Qf[m][h], Qf[m][h]), accI);
                    accR_t = fma2(tr, pack2(qt, qt), accR_t);
                    accI_t = fma2(ti, pack2(qt, qt), accI_t);
                } else {
                    accR = fma2(drv, cv, accR);
                    accR = fma2(div, sv, accR);
                    accI = fma2(div, cv, accI);
                    accI = fma2(neg2(drv), sv, accI);
                }
            } else {
                float rho, sn, dum;
                unpack2(rot.rho[idx], rho, dum);
                unpack2(rot.sn[idx], sn, dum);
                float c1 = fmaf(pc, rho, pc), s1 = fmaf(ps, rho, ps);
                if (pos) { c1 = fmaf(-ps, sn, c1); s1 = fmaf(pc, sn, s1); }
                else     { c1 = fmaf(ps, sn, c1);  s1 = fmaf(-pc, sn, s1); }
                const unsigned long long d1 = *reinterpret_cast<const unsigned long long*>(rb + R::kData + 8);
                float dr, di, ar, ai;
                unpack2(d1, dr, di);
                if constexpr (SPCAL) {
                    unpack2(mul2(d1, pack2(c1, c1)), ar, ai);
                    const unsigned long long term = pack2(fmaf(di, s1, ar), fmaf(-dr, s1, ai));
                    accR = fma2(term, pack2(Qf[m][h], Qf[m][h]), accR);
                    accR_t = fma2(term, pack2(qt, qt), accR_t);
                } else {
                    unpack2(fma2(d1, pack2(c1, c1), accR), ar, ai);
                    accR = pack2(fmaf(di, s1, ar), fmaf(-dr, s1, ai));
                }
            }
        }
    }
}

template <int NIFO, int MODEL, int SC, bool IRAMP, bool SPCAL, bool ROT>
__global__ void __launch_bounds__(kTileThreads, 512 / kTileThreads) fullgrid_tile_kernel(const unsigned char* __restrict__ tab_basis,
                                                                  const unsigned char* __restrict__ tab_data,
                                                                  const unsigned char* __restrict__ rec_ul,
                                                                  int rec_ul_stride,
                                                                  const ChunkDesc* __restrict__ chunks,
                                                                  const double* __restrict__ coef, long stride,
                                                                  long n_walkers, double* __restrict__ partial,
                                                                  int chunk0, double df,
                                                                  const double* __restrict__ coef_ex, ExtrasDims dims,
                                                                  Sel sel, const __grid_constant__ MultiTable mt) {
    using R = TileRec<MODEL>;
    constexpr int RB = R::kDataBytes, BB = R::kBasisBytes;
    constexpr int kStageBytes = R::kTileBins * (BB + RB);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stages = smem_raw;                    // per stage: [R::kTileBins x BB basis][R::kTileBins x RB data]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStagesT * kStageBytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // work list (Sel): entry -> walker; a CTA beyond the end of the list has nothing to do
    long n_sel, entry0 = (long)blockIdx.y * kTileWalkers;
    if (mt.n > 0) {
        int lv;
        if (!resolve_tile(mt, kTileWalkers, (int)blockIdx.y, lv, entry0, n_sel)) return;
        const TableRef& T = mt.t[lv];
        if ((int)blockIdx.x >= T.n_chunks - T.n_precise) return;
        tab_basis = T.basis;
        tab_data = T.data;
        rec_ul = T.rec_ul;
        chunks = T.chunks;
        chunk0 = T.n_precise;
        sel.idx = mt.lists + (size_t)lv * mt.list_stride;
    } else {
        n_sel = sel.idx ? (long)*sel.count : n_walkers;
        if (entry0 >= n_sel) return;
    }
    const long entry = entry0 + warp * 8 + g;
    const bool live = entry < n_sel;
    const long wl = sel.idx ? (long)sel.idx[live ? entry : n_sel - 1] : (live ? entry : n_sel - 1);
    const long wi = wl;

    // A fragments: row g (this lane's walker), columns 4 s + t; the magic constant rides on the constant term
    double afrag[R::kSteps];
#pragma unroll
    for (int s = 0; s < R::kSteps; ++s) {
        // t is a run-time value: pick among the four compile-time slots of this k-step
        const int s0 = tile_basis_slot<MODEL>(4 * s + 0), s1 = tile_basis_slot<MODEL>(4 * s + 1);
        const int s2 = tile_basis_slot<MODEL>(4 * s + 2), s3 = tile_basis_slot<MODEL>(4 * s + 3);
        const int slot = t == 0 ? s0 : t == 1 ? s1 : t == 2 ? s2 : s3;
        double a = slot >= 0 ? coef[(size_t)slot * stride + wl] : 0.0;
        if (slot == C_K0) a += kMagic;
        afrag[s] = a;
    }
    TileWalker<NIFO> w;
#pragma unroll
    for (int j = 0; j < 6; ++j) w.q[j] = (float)coef[(C_Q2 + j) * stride + wl];
    w.q6l = (float)coef[C_Q6L * stride + wl];
#pragma unroll
    for (int i = 0; i < NIFO; ++i) {
        w.tau[i] = coef[(C_TAU0 + i) * stride + wl];
        w.d1[i] = (uint32_t)__double2loint(fma(df, w.tau[i], kMagic));
        w.d7[i] = 7u * w.d1[i];
    }
    const double df16 = 16.0 * df;

    double zr[NIFO], zi[NIFO], yr[NIFO], yi[NIFO];            // y: the t-weighted sums (SPCAL)
    unsigned long long acc[NIFO], acc1[NIFO];                 // ROT: acc[0] = detector 0, acc[1] = accR, acc[2] = accI
#pragma unroll
    for (int i = 0; i < NIFO; ++i) { zr[i] = 0.0; zi[i] = 0.0; yr[i] = 0.0; yi[i] = 0.0; acc[i] = 0ull; acc1[i] = 0ull; }

    const int chunk_id = chunk0 + (int)blockIdx.x;
    const ChunkDesc chunk = chunks[chunk_id];
    const long bin0 = chunk.start;
    const int nb_total = chunk.len;
    const int n_tiles = (nb_total + R::kTileBins - 1) / R::kTileBins;

    TileRot<NIFO> rot;
    if constexpr (ROT) {
        constexpr int NR = NIFO - 1;
        float rf[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                rf[r][j] = r < NR ? (float)coef[(size_t)(C_ROT0 + kRotSlots * r + j) * stride + wl] : 0.0f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            rot.rho[a] = pack2(rf[0][a], rf[1][a]);
            rot.sn[a] = pack2(rf[0][4 + a], rf[1][4 + a]);
        }
        // anchor of the chunk's first step at this lane's centre bin (2 t + 12.5 bins after the chunk's first bin)
        const double f_first = *reinterpret_cast<const double*>(tab_data + (size_t)bin0 * RB + R::kF);
        const double f_c = f_first + (2.0 * t + 12.5) * df;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            rot.rr[r] = coef[(size_t)(C_ROT0 + kRotSlots * r + 8) * stride + wl];
            rot.ri[r] = coef[(size_t)(C_ROT0 + kRotSlots * r + 9) * stride + wl];
            const double turns = f_c * (w.tau[r + 1] - w.tau[0]);
            double sn, cs;
            sincospi(2.0 * (turns - rint(turns)), &sn, &cs);
            rot.er[r] = cs;
            rot.ei[r] = -sn;
        }
    }

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStagesT; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int slot) {
        const int nb = min(R::kTileBins, nb_total - tile * R::kTileBins);
        const long first = bin0 + (long)tile * R::kTileBins;
        unsigned char* dst = stages + (size_t)slot * kStageBytes;
        mbar_expect_tx(&full[slot], (uint32_t)nb * (BB + RB));
        tma_bulk_g2s(dst, tab_basis + first * BB, (uint32_t)nb * BB, &full[slot]);
        tma_bulk_g2s(dst + (size_t)R::kTileBins * BB, tab_data + first * RB, (uint32_t)nb * RB, &full[slot]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesT && s < n_tiles; ++s) issue(s, s);
    }

    double corr = 1.0;
    auto flush = [&]() {
        if constexpr (ROT) {
            constexpr int NR = NIFO - 1;
            float re, im;
            unpack2(acc[0], re, im);
            zr[0] = fma((double)re, corr, zr[0]);
            zi[0] = fma((double)im, corr, zi[0]);
            if constexpr (SPCAL) {
                unpack2(acc1[0], re, im);
                yr[0] = fma((double)re, corr, yr[0]);
                yi[0] = fma((double)im, corr, yi[0]);
            }
            float sr[2], si[2], tr[2] = {0.0f, 0.0f}, ti[2] = {0.0f, 0.0f};
            if constexpr (NR == 2) {
                unpack2(acc[1], sr[0], sr[1]);
                unpack2(acc[2], si[0], si[1]);
                if constexpr (SPCAL) { unpack2(acc1[1], tr[0], tr[1]); unpack2(acc1[2], ti[0], ti[1]); }
            } else {
                unpack2(acc[1], sr[0], si[0]);
                sr[1] = si[1] = 0.0f;
                if constexpr (SPCAL) unpack2(acc1[1], tr[0], ti[0]);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                // z += corr E_r (sr + i si), then E_r advances by one step
                const double fr = corr * rot.er[r], fi = corr * rot.ei[r];
                zr[r + 1] = fma((double)sr[r], fr, fma(-(double)si[r], fi, zr[r + 1]));
                zi[r + 1] = fma((double)sr[r], fi, fma((double)si[r], fr, zi[r + 1]));
                if constexpr (SPCAL) {
                    yr[r + 1] = fma((double)tr[r], fr, fma(-(double)ti[r], fi, yr[r + 1]));
                    yi[r + 1] = fma((double)tr[r], fi, fma((double)ti[r], fr, yi[r + 1]));
                }
                const double nr = fma(rot.er[r], rot.rr[r], -rot.ei[r] * rot.ri[r]);
                const double ni = fma(rot.er[r], rot.ri[r], rot.ei[r] * rot.rr[r]);
                rot.er[r] = nr;
                rot.ei[r] = ni;
            }
#pragma unroll
            for (int i = 0; i < NIFO; ++i) { acc[i] = 0ull; acc1[i] = 0ull; }
        } else {
#pragma unroll
            for (int i = 0; i < NIFO; ++i) {
                float re, im;
                unpack2(acc[i], re, im);
                zr[i] = fma((double)re, corr, zr[i]);
                zi[i] = fma((double)im, corr, zi[i]);
                acc[i] = 0ull;
                if constexpr (SPCAL) {
                    unpack2(acc1[i], re, im);
                    yr[i] = fma((double)re, corr, yr[i]);
                    yi[i] = fma((double)im, corr, yi[i]);
                    acc1[i] = 0ull;
                }
            }
        }
    };

    for (int tl = 0; tl < n_tiles; ++tl) {
        const int slot = tl % kStagesT;
        const unsigned char* tbasis = stages + (size_t)slot * kStageBytes;
        const unsigned char* tdata = tbasis + (size_t)R::kTileBins * BB;
        const int nb = min(R::kTileBins, nb_total - tl * R::kTileBins);
        // systematic factor of the FP32 amplitude series (amp_series_correction), one value per PAIR of 32-bin steps,
        // taken at the pair's centre bin: lane t of a walker's quad evaluates pair t of the
        // stage, the step loop fetches it by shuffle.  (u, L) of a bin come from the pair-format table.
        static_assert(R::kTileBins <= 8 * kStepBins, "one correction value per lane covers at most 4 step pairs");
        double corr_mine = 1.0;
        if (2 * t * kStepBins < nb) {
            const long centre = bin0 + (long)tl * R::kTileBins + min((2 * t + 1) * kStepBins, nb - 1);
            const double2 h0 = *reinterpret_cast<const double2*>(rec_ul + (size_t)centre * rec_ul_stride);
            corr_mine = amp_series_correction(w.q, w.q6l, h0.x, h0.y, coef, stride, wl);
        }
        mbar_wait(&full[slot], (uint32_t)((tl / kStagesT) & 1));
        const int n_steps = nb / kStepBins;
#pragma unroll 1
        for (int st = 0; st < n_steps; ++st) {
            if constexpr (ROT)
                tile_step_rot<NIFO, MODEL, SC, false, SPCAL>(tbasis + (size_t)st * kStepBins * BB,
                                                             tdata + (size_t)st * kStepBins * RB, afrag, w, rot, g, t,
                                                             kStepBins, df16, acc[0], acc[1], acc[NIFO - 1], acc1[0],
                                                             acc1[1], acc1[NIFO - 1]);
            else
            tile_step<NIFO, MODEL, SC, false, IRAMP, SPCAL>(tbasis + (size_t)st * kStepBins * BB,
                                                            tdata + (size_t)st * kStepBins * RB, afrag, w, g, t,
                                                            kStepBins, df16, acc, acc1);
            // FP32 sums of 8 bins per lane go to FP64 (16 would save 1.5 % and double the worst-case rounding noise:
            // 1 walker in 50 000 then leaves the tolerance)
            corr = __shfl_sync(0xffffffffu, corr_mine, (lane & ~3) | (st >> 1));
            flush();
        }
        if (n_steps * kStepBins < nb) {
            if constexpr (ROT)
                tile_step_rot<NIFO, MODEL, SC, true, SPCAL>(tbasis + (size_t)n_steps * kStepBins * BB,
                                                            tdata + (size_t)n_steps * kStepBins * RB, afrag, w, rot, g, t,
                                                            nb - n_steps * kStepBins, df16, acc[0], acc[1], acc[NIFO - 1],
                                                            acc1[0], acc1[1], acc1[NIFO - 1]);
            else
            tile_step<NIFO, MODEL, SC, true, IRAMP, SPCAL>(tbasis + (size_t)n_steps * kStepBins * BB,
                                                           tdata + (size_t)n_steps * kStepBins * RB, afrag, w, g, t,
                                                           nb - n_steps * kStepBins, df16, acc, acc1);
            corr = __shfl_sync(0xffffffffu, corr_mine, (lane & ~3) | (n_steps >> 1));
            flush();
        }
        __syncthreads();
        if (threadIdx.x == 0 && tl + kStagesT < n_tiles) {
            fence_proxy_async();
            issue(tl + kStagesT, slot);
        }
    }

    if constexpr (SPCAL) {
        // calibration envelope of the chunk's segment, exact FP64 node values: z <- c0 z + (c1 - c0) y
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            const double c0r = coef_ex[ex_cal_slot(dims, i, chunk.seg, 0) * stride + wl];
            const double c0i = coef_ex[ex_cal_slot(dims, i, chunk.seg, 1) * stride + wl];
            const double dr = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 0) * stride + wl] - c0r;
            const double di = coef_ex[ex_cal_slot(dims, i, chunk.seg + 1, 1) * stride + wl] - c0i;
            const double nr = c0r * zr[i] - c0i * zi[i] + dr * yr[i] - di * yi[i];
            const double ni = c0r * zi[i] + c0i * zr[i] + dr * yi[i] + di * yr[i];
            zr[i] = nr;
            zi[i] = ni;
        }
    }
    // the four lanes of a walker hold disjoint bins: fixed-order butterfly, lane t = 0 stores
#pragma unroll
    for (int i = 0; i < NIFO; ++i) {
        zr[i] += __shfl_xor_sync(0xffffffffu, zr[i], 1);
        zi[i] += __shfl_xor_sync(0xffffffffu, zi[i], 1);
        zr[i] += __shfl_xor_sync(0xffffffffu, zr[i], 2);
        zi[i] += __shfl_xor_sync(0xffffffffu, zi[i], 2);
    }
    if (t == 0 && live) {
        double* out = partial + (size_t)chunk_id * (2 * NIFO) * stride + wi;
#pragma unroll
        for (int i = 0; i < NIFO; ++i) {
            out[(2 * i + 0) * stride] = zr[i];
            out[(2 * i + 1) * stride] = zi[i];
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

// ---- epilogue: fixed-order reduction over chunks, <h|h> from the Gram moments, logL -------------------
// sums[W][n_ifo][3] (optional) receives Re<d|h>, Im<d|h>, <h|h> per detector.
constexpr int kEpiWalkers = 32;   // walkers per epilogue CTA
constexpr int kEpiSlices = 8;     // chunk slices summed in parallel, then combined in fixed order
// Static tables of the epilogue.  Cells = runs of bins with constant (PSD-weight band, calibration segment); without
// extras there is ONE cell.  gram[det][cell][3][kGram]: Gram moments of <h|h> weighted by (1-t)^2, t(1-t), t^2 (the
// piecewise-linear |cal|^2); without calibration only the first set is non-zero and carries weight 1.
struct EpilogueTables {
    const ChunkDesc* chunks;
    int n_chunks;
    int n_cells;
    const int* cell_band;
    const int* cell_seg;
    const double* gram;       // [n_ifo][n_cells][3][kGram]
    const double* dd_band;    // [n_ifo][nweights]  (4/T) sum_{k in band} |d|^2 / S
};

// mode 0: reduce the chunks [chunk_begin, chunk_end) and finish (the normal single-GPU epilogue)
// mode 1: reduce the chunks only and write sums_io[2 n_ifo][n_walkers]            (frequency-split, before all-reduce)
// mode 2: take sums_io as the already reduced sums and finish                     (frequency-split, after all-reduce)
__global__ void __launch_bounds__(kEpiWalkers * kEpiSlices)
epilogue_kernel(DeviceConfig cfg, EpilogueTables tab, ExtrasDims dims, const double* __restrict__ partial,
                const double* __restrict__ coef, const double* __restrict__ coef_ex, long stride, long n_walkers,
                double* __restrict__ logl, double* __restrict__ sums, int mode, int chunk_begin, int chunk_end,
                double* __restrict__ sums_io, Sel sel, const __grid_constant__ MultiTable mt) {
    __shared__ double red[kEpiSlices][2 * BJ_MAX_IFO][kEpiWalkers];
    const int lane_w = threadIdx.x % kEpiWalkers, slice = threadIdx.x / kEpiWalkers;
    long n_sel, entry0 = (long)blockIdx.x * kEpiWalkers;
    if (mt.n > 0) {
        int lv;
        if (!resolve_tile(mt, kEpiWalkers, (int)blockIdx.x, lv, entry0, n_sel)) return;
        const TableRef& T = mt.t[lv];
        tab.chunks = T.chunks;
        tab.n_chunks = T.n_chunks;
        chunk_begin = 0;
        chunk_end = T.n_chunks;
        cfg.dh_fix_re = T.dh_fix_re;
        cfg.dh_fix_im = T.dh_fix_im;
        sel.idx = mt.lists + (size_t)lv * mt.list_stride;
    } else {
        n_sel = sel.idx ? (long)*sel.count : n_walkers;
        if (entry0 >= n_sel) return;
    }
    const long entry = entry0 + lane_w;
    const bool live = entry < n_sel;
    const long w = sel.idx ? (long)sel.idx[live ? entry : n_sel - 1] : (live ? entry : n_sel - 1);
    const long wc = w;
    const int nifo = cfg.n_ifo;
    // stage 1: slice s adds chunks s, s + kEpiSlices, ... in ascending order (fixed tree: bits do not depend on W);
    // with PSD weights every chunk lies in one band and is scaled by 1 / w_band (detector.py:529).  All 2 n_ifo sums of a
    // chunk are loaded together and two chunks per round, so that ~12 loads are in flight (a call with few walkers is
    // bound by this chain); every sum still receives its terms in ascending chunk order.
    {
        double acc[2 * BJ_MAX_IFO];
#pragma unroll
        for (int j = 0; j < 2 * BJ_MAX_IFO; ++j) acc[j] = 0.0;
        if (mode != 2) {
#pragma unroll 2
            for (int ch = chunk_begin + slice; ch < chunk_end; ch += kEpiSlices) {
                double v[2 * BJ_MAX_IFO];
#pragma unroll
                for (int j = 0; j < 2 * BJ_MAX_IFO; ++j)
                    v[j] = j < 2 * nifo ? partial[((size_t)ch * (2 * nifo) + j) * stride + wc] : 0.0;
                if (dims.nweights > 0) {
                    const int band = tab.chunks[ch].band;
#pragma unroll
                    for (int j = 0; j < 2 * BJ_MAX_IFO; ++j)
                        if (j < 2 * nifo) v[j] *= coef_ex[ex_invw_slot(dims, nifo, j >> 1, band) * stride + wc];
                }
#pragma unroll
                for (int j = 0; j < 2 * BJ_MAX_IFO; ++j) acc[j] += v[j];
            }
        } else if (slice == 0) {
#pragma unroll
            for (int j = 0; j < 2 * BJ_MAX_IFO; ++j)
                if (j < 2 * nifo) acc[j] = sums_io[(size_t)j * n_walkers + wc];
        }
#pragma unroll
        for (int j = 0; j < 2 * BJ_MAX_IFO; ++j)
            if (j < 2 * nifo) red[slice][j][lane_w] = acc[j];
    }
    __syncthreads();
    if (slice != 0 || !live) return;
    if (mode == 1) {
        for (int j = 0; j < 2 * nifo; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int sl = 0; sl < kEpiSlices; ++sl) acc += red[sl][j][lane_w];
            sums_io[(size_t)j * n_walkers + w] = acc;
        }
        return;
    }

    const double amp = coef[C_AMP * stride + w];
    const bool valid = coef[C_VALID * stride + w] != 0.0;
    // amplitude-series coefficients in the basis phi = {1, u^2, ..., u^7, L u^6}
    double g[kBasis];
    g[0] = 1.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) g[1 + j] = coef[(C_Q2 + j) * stride + w];
    g[7] = coef[C_Q6L * stride + w];
    double dh_re = 0.0, dh_im = 0.0, hh_tot = 0.0, dd_tot = 0.0;
    for (int i = 0; i < nifo; ++i) {
        double zr = 0.0, zi = 0.0;
#pragma unroll
        for (int sl = 0; sl < kEpiSlices; ++sl) {
            zr += red[sl][2 * i][lane_w];
            zi += red[sl][2 * i + 1][lane_w];
        }
        // <h|h>_i: quadratic form of the Gram moments, per cell scaled by 1/w_band and |cal|^2 of its segment
        double h = 0.0;
        for (int cell = 0; cell < tab.n_cells; ++cell) {
            double a3[3] = {1.0, 0.0, 0.0};
            int nset = 1;
            if (dims.nspcal > 0) {
                const int sg = tab.cell_seg[cell];
                const double c0r = coef_ex[ex_cal_slot(dims, i, sg, 0) * stride + w];
                const double c0i = coef_ex[ex_cal_slot(dims, i, sg, 1) * stride + w];
                const double c1r = coef_ex[ex_cal_slot(dims, i, sg + 1, 0) * stride + w];
                const double c1i = coef_ex[ex_cal_slot(dims, i, sg + 1, 1) * stride + w];
                a3[0] = c0r * c0r + c0i * c0i;                  // |(1-t) c0 + t c1|^2
                a3[1] = 2.0 * (c0r * c1r + c0i * c1i);
                a3[2] = c1r * c1r + c1i * c1i;
                nset = 3;
            }
            double hc = 0.0;
            for (int st = 0; st < nset; ++st) {
                const double* G = tab.gram + (((size_t)i * tab.n_cells + cell) * 3 + st) * kGram;
                double q = 0.0;
                int idx = 0;
                for (int m = 0; m < kBasis; ++m) {
                    double row = 0.0;
                    for (int n = m; n < kBasis; ++n, ++idx) row = fma((n == m ? 1.0 : 2.0) * g[n], G[idx], row);
                    q = fma(g[m], row, q);
                }
                hc = fma(a3[st], q, hc);
            }
            if (dims.nweights > 0) hc *= coef_ex[ex_invw_slot(dims, nifo, i, tab.cell_band[cell]) * stride + w];
            h += hc;
        }
        if (dims.nweights > 0) {
            for (int b = 0; b < dims.nweights; ++b)
                dd_tot += tab.dd_band[i * dims.nweights + b] * coef_ex[ex_invw_slot(dims, nifo, i, b) * stride + w];
        }
        const double cr = coef[(C_CR0 + i) * stride + w];
        const double ci = coef[(C_CI0 + i) * stride + w];
        // undo the (power-of-two) table scale and the calibrated phasor bias: z -> z * (kr + i ki)
        const double xr = zr * cfg.dh_fix_re - zi * cfg.dh_fix_im;
        const double xi = zr * cfg.dh_fix_im + zi * cfg.dh_fix_re;
        const double re = amp * (cr * xr - ci * xi);
        const double im = amp * (cr * xi + ci * xr);
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
    // log_like.py:181:  -1/2 (hh + dd) + R - logZ_noise - 1/2 _psd_fact
    const double dd = dims.nweights > 0 ? dd_tot : cfg.dd_sum;
    const double pf = dims.nweights > 0 ? coef_ex[ex_psdfact_slot(dims, nifo) * stride + w] : 0.0;
    double v = -0.5 * (hh_tot + dd) + R - cfg.logZ_noise - 0.5 * pf;
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
// the same with the in-house FP64 polynomials (|error| < 1e-14: ROQ / relative-binning kernels, where the libdevice
// sincospi was a third of the instructions)
__device__ __forceinline__ void cis_neg_turns_fast(double turns, double& c, double& s_neg) {
    double s;
    sincos_turns_f64(turns, s, c);
    s_neg = -s;
}
// eval_point for the 3.5PN + 6PN-tides model: the coefficients a[1], a[8], a[9], a[11], a[13..15], b[2..6], c8 are
// structurally zero (walker.cuh), so the phase needs 13 instead of 30 FMAs
template <int MODEL>
__device__ __forceinline__ void eval_point_m(const WalkerCoefReg& r, double u, double L, double w5, double& turns,
                                             double& Q) {
    if constexpr (MODEL == MODEL_35) {
        const double u2 = u * u;
        double p = fma(r.a[12], u2, r.a[10]);
        p = fma(p, u2 * u, r.a[7]);
#pragma unroll
        for (int k = 6; k >= 2; --k) p = fma(p, u, r.a[k]);
        p = fma(p, u2, r.a[0]);
        const double bl = fma(r.b[1], u, r.b[0]);
        turns = fma(w5, p, fma(L, bl, r.k0));
        double qq = fma(r.q6l, L, r.q[4]);
        qq = fma(r.q[5], u, qq);
        qq = fma(qq, u, r.q[3]);
        qq = fma(qq, u, r.q[2]);
        qq = fma(qq, u, r.q[1]);
        qq = fma(qq, u, r.q[0]);
        Q = fma(u2, qq, 1.0);
    } else {
        eval_point(r, u, L, w5, turns, Q);
    }
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
