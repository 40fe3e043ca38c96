// C ABI of bajes_b200 (see include/bajes_b200.h).  Host-side context management, static table
// construction (the set-up rows of SURVEY.md section 8a: Detector.store_measurement,
// detector.py:452-493) and kernel launches.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "roq_relbin.cuh"
#include "roq_basis.cuh"
#include "timemarg.cuh"

using namespace bj;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static bj::CufftApi g_cufft;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return fail(BJ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* bj_last_error(void) { return g_last_error.c_str(); }
extern "C" int bj_abi_version(void) { return BJ_ABI_VERSION; }

extern "C" int bj_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                              int* clock_khz) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(BJ_ERR_NO_DEVICE, "no CUDA device visible");
    }
    if (device < 0 || device >= n) return fail(BJ_ERR_INVALID, "device %d out of range (have %d)", device, n);
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (name && name_len > 0) {
        strncpy(name, p.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    if (clock_khz) *clock_khz = khz;
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
enum { KIND_FULLGRID = 0, KIND_ROQ = 1, KIND_RELBIN = 2 };

struct bj_ctx {
    int kind = KIND_FULLGRID;
    int device = 0;
    int sm_count = 0;
    int sincos = BJ_SINCOS_MUFU;
    DeviceConfig cfg;
    // full grid
    long n_bins = 0;
    long n_bins_kernel = 0;       // n_bins padded to even for the mixed-precision kernel
    int rec_doubles = 0;
    double calib[2] = {0.0, 0.0}; // measured phasor bias (eps_a, eps_p)
    void* d_rec = nullptr;
    void* d_tile_basis = nullptr; // tables of the tensor-core kernel; null: Horner kernels only
    double grid_df = 0.0;         // bin spacing if the in-band grid is uniform (integer time-delay ramps), else 0
    double window_df = 0.0;       // bin spacing of the slots the FP64 window covers if they are uniform (fullgrid_window_kernel), else 0
    bool rot = false;             // tile tables in the rotation layout (kernels.cuh: tile_step_rot)
    int n_precise = 0;            // chunks [0, n_precise): the FP64 window (heaviest bins), evaluated by fullgrid_kernel
    double* d_rec64 = nullptr;    // FP64 records of the window's slots [prec_lo, prec_hi)
    long prec_lo = 0, prec_hi = 0;
    void* d_tile_data = nullptr;
    void* d_tile_data_plain = nullptr;   // the same records in the per-detector layout (contexts with compressed axes: launch_multi)
    int table_exp = 0;                   // 2^-e scale of the mixed-precision tables
    double* d_freqs = nullptr;
    double* d_gram = nullptr;
    int bins_per_chunk = 0;
    int n_chunks = 0;
    // cells / chunks / extras (calibration envelopes, PSD weights)
    ExtrasDims dims;
    int n_cells = 1;
    ChunkDesc* d_chunks = nullptr;
    int* d_cell_band = nullptr;
    int* d_cell_seg = nullptr;
    double* d_dd_band = nullptr;
    double* d_spcal_freqs = nullptr;
    double* d_coef_ex = nullptr;
    // time-shift marginalisation
    int tm = 0;
    long tm_n_full = 0, tm_offset = 0, tm_batch = 0;
    double2* d_tm = nullptr;
    double* d_tm_R = nullptr;
    long cap_tm_R = 0;
    int tm_plan = 0;              // plan of tm_batch rows
    bool tm_plan_ok = false;
    int tm_plan_small = 0;        // plan of tm_small rows (calls with few walkers: a sampler's single points)
    long tm_small = 0;
    // ROQ / relbin static tables
    RoqDevice roq;
    RelbinDevice relbin;
    // per-call workspace (grown on demand)
    long cap_walkers = 0;
    double* d_coef = nullptr;
    double* d_partial = nullptr;
    double* d_x = nullptr;
    long cap_x = 0;
    double* d_logl = nullptr;
    long cap_logl = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_tile[2] = {nullptr, nullptr};      // around the launch of the dominant (tile) kernel of the last call
    bool tile_timed = false;
    bool timed = false;
    bool shadow = false;          // a compressed level of another context: owns static tables only
    int launch_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // compressed frequency axis (build_compressed_axis): a second set of mixed-precision tables on Chebyshev nodes.  `cmp`
    // owns its static tables only; workspace, Gram moments, stream and events are this context's and are lent per call.
    bj_ctx* cmp[kMaxLevels] = {nullptr, nullptr};   // levels of growing design budget (kernels.cuh: classify_kernel)
    int n_levels = 0;
    SteepTable steep;
    int* d_lists = nullptr;           // [n_levels + 1][cap_walkers] work lists: one per level, the last = the full grid
    int* d_counts = nullptr;          // [kMaxLevels + 1]
    int cmp_info[kMaxLevels][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // nodes, nodes that are plain bins, blocks, nodes per block
    double cmp_spec[kMaxLevels][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};   // mchirp_min, tau_max, reach, margin
    long last_walkers = 0;
    // focus table (bj_focus): all-FP64 node records around a fiducial point; host copies of the grid's weights to build it
    std::vector<double> h_freqs, h_wts[BJ_MAX_IFO];
    bj_ctx* focus = nullptr;          // owns d_rec (FP64 records), d_chunks
    double* d_fid = nullptr;          // {A_0..15, B_0..6, C8, tau_0..2} of the fiducial point
    int focus_info[4] = {0, 0, 0, 0};
};

static int check_config(const bj_config* cfg) {
    if (!cfg) return fail(BJ_ERR_INVALID, "null config");
    if (cfg->abi_version != BJ_ABI_VERSION)
        return fail(BJ_ERR_INVALID, "ABI version mismatch: caller %d, library %d", cfg->abi_version, BJ_ABI_VERSION);
    if (cfg->n_ifo < 1 || cfg->n_ifo > BJ_MAX_IFO)
        return fail(BJ_ERR_UNSUPPORTED, "n_ifo=%d outside 1..%d", cfg->n_ifo, BJ_MAX_IFO);
    if (cfg->approx < 0 || cfg->approx > BJ_APPROX_TAYLORF2_55PN_75PNTIDES2020)
        return fail(BJ_ERR_UNSUPPORTED, "approximant id %d is not a TaylorF2 variant", cfg->approx);
    if (cfg->sincos < 0 || cfg->sincos > BJ_SINCOS_FP64) return fail(BJ_ERR_INVALID, "bad sincos mode %d", cfg->sincos);
    if (!(cfg->seglen > 0)) return fail(BJ_ERR_INVALID, "seglen must be positive");
    return BJ_OK;
}

static int init_common(bj_ctx* c, const bj_config* cfg) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(BJ_ERR_NO_DEVICE, "no CUDA device visible: bajes_b200 has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= n) return fail(BJ_ERR_INVALID, "device %d out of range", cfg->device);
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, cfg->device));
    if (p.major != 10)
        return fail(BJ_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device,
                    p.major, p.minor);
    c->device = cfg->device;
    c->sm_count = p.multiProcessorCount;
    c->sincos = cfg->sincos;
    c->cfg.approx = cfg->approx;
    c->cfg.n_ifo = cfg->n_ifo;
    c->cfg.marg_phi_ref = cfg->marg_phi_ref;
    c->cfg.seglen = cfg->seglen;
    c->cfg.dd_sum = cfg->dd_sum;
    c->cfg.logZ_noise = cfg->logZ_noise;
    c->cfg.dh_fix_re = 1.0;
    c->cfg.dh_fix_im = 0.0;
    c->cfg.f_lo = 0.0;
    c->cfg.f_hi = 0.0;
    c->cfg.grid_df = 0.0;
    memset(&c->dims, 0, sizeof(c->dims));
    for (int i = 0; i < BJ_MAX_IFO; ++i) c->cfg.det[i] = cfg->det[i];
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) CUDA_TRY(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreate(&c->ev_tile[i]));
    return BJ_OK;
}

// per-bin quantities shared by both table layouts
struct BinStatics {
    double u, L, w5, f, g, cs, sn;
};
static BinStatics bin_statics(double f, double seglen) {
    BinStatics b;
    b.f = f;
    b.u = std::cbrt(f);
    b.L = std::log(f);
    b.w5 = 1.0 / (b.u * b.u * b.u * b.u * b.u);
    b.g = std::pow(f, -7.0 / 6.0);
    // exp(+i pi f T): exactly +-1 on the FFT grid (waveform.py:392-393)
    const double kk = f * seglen;
    const double kr = std::nearbyint(kk);
    if (std::fabs(kk - kr) < 1e-6) {
        b.cs = (std::fmod(std::fabs(kr), 2.0) == 1.0) ? -1.0 : 1.0;
        b.sn = 0.0;
    } else {
        b.cs = std::cos(M_PI * kk);
        b.sn = std::sin(M_PI * kk);
    }
    return b;
}
// (4/T) conj(d)/S * f^(-7/6) * exp(i pi f T)
static void data_weight(const BinStatics& b, double four_over_t, double dre, double dim, double psd, double& wr,
                        double& wi) {
    const double inv_s = 1.0 / psd;          // psd = inf outside the ASD table -> weight 0
    const double sc = four_over_t * inv_s * b.g;
    const double cre = dre, cim = -dim;
    wr = sc * (cre * b.cs - cim * b.sn);
    wi = sc * (cre * b.sn + cim * b.cs);
}

// A frequency axis the table builders work on: bin (or node) frequencies and, per detector, the complex weight
// (4/T) conj(d)/S f^(-7/6) exp(i pi f T) that multiplies Q e^{-i theta} in <d|h>.  The full grid carries the weights of its
// bins (full_axis_weights); the compressed axis carries Chebyshev-node weights (build_compressed_axis).
struct Axis {
    long n = 0;
    const double* f = nullptr;
    const double* w[BJ_MAX_IFO] = {nullptr, nullptr, nullptr};    // interleaved (re, im)
};
static void full_axis_weights(std::vector<double>* wts, int nifo, long n, const double* freqs,
                              const double* const* data_ri, const double* const* psd, double seglen) {
    const double four_over_t = 4.0 / seglen;
    for (int i = 0; i < nifo; ++i) wts[i].assign((size_t)2 * n, 0.0);
    for (long k = 0; k < n; ++k) {
        const BinStatics b = bin_statics(freqs[k], seglen);
        for (int i = 0; i < nifo; ++i)
            data_weight(b, four_over_t, data_ri[i][2 * k], data_ri[i][2 * k + 1], psd[i][k], wts[i][2 * k], wts[i][2 * k + 1]);
    }
}

// ------------------------------------------------------------------------------------------------
// table layout: cells (runs of bins with constant PSD-weight band and calibration segment), padded to an even
// number of records, cut into chunks
// ------------------------------------------------------------------------------------------------
struct Layout {
    std::vector<long> slot_bin;      // table slot -> in-band bin whose statics it carries
    std::vector<char> slot_real;     // 0 for zero-weight padding slots
    std::vector<double> slot_t;      // calibration interpolation fraction in [0, 1]
    std::vector<int> slot_cell;
    std::vector<int> cell_band, cell_seg;
    std::vector<ChunkDesc> chunks;
    int per = 0;
    int n_precise = 0;               // chunks [0, n_precise) of the list: the window evaluated in FP64
    long prec_lo = 0, prec_hi = 0;   // its slot range
};

// binweight[k] ~ share of bin k in <h|h> of an inspiral (sum_i f^(-7/3) / S_i).  The shortest run of consecutive slots
// that holds the fraction prec_share of the sum of SQUARED weights (the bins whose terms are signal-dominated at high SNR,
// where FP32 rounding does not average down), at most prec_frac of the table, is cut into 64-record chunks that a launch of
// the FP64 window kernel evaluates.  For the design PSDs of H1/L1/V1 a share of 0.9 is the band 20-87 Hz: 1.6 % of the
// bins of a 128 s segment sampled to 4096 Hz.
static int build_layout(Layout& L, long n, const double* freqs, const bj_extras* ex, const std::vector<double>* binweight,
                        double prec_share, double prec_frac, long window_limit = -1) {
    const int nsp = ex ? ex->nspcal : 0, nw = ex ? ex->nweights : 0;
    std::vector<int> band(n, 0), seg(n, 0);
    std::vector<double> tt(n, 0.0);
    if (nw > 0) {
        long k = 0;
        for (int b = 0; b < nw; ++b)
            for (int64_t m = 0; m < ex->len_weights[b] && k < n; ++m) band[k++] = b;
        if (k != n) return -1;
    }
    if (nsp > 0) {
        const double* nf = ex->spcal_freqs;
        for (long k = 0; k < n; ++k) {
            const double f = freqs[k];
            if (f <= nf[0]) { seg[k] = 0; tt[k] = 0.0; }                       // np.interp clamps (detector.py:519)
            else if (f >= nf[nsp - 1]) { seg[k] = nsp - 2; tt[k] = 1.0; }
            else {
                int j = 0;
                while (j + 2 < nsp && f >= nf[j + 1]) ++j;
                seg[k] = j;
                tt[k] = (f - nf[j]) / (nf[j + 1] - nf[j]);
            }
        }
    }
    long k = 0;
    while (k < n) {
        const int cell = (int)L.cell_band.size();
        L.cell_band.push_back(band[k]);
        L.cell_seg.push_back(seg[k]);
        long k1 = k;
        while (k1 < n && band[k1] == band[k] && seg[k1] == seg[k]) {
            L.slot_bin.push_back(k1); L.slot_real.push_back(1); L.slot_t.push_back(tt[k1]); L.slot_cell.push_back(cell);
            ++k1;
        }
        if ((k1 - k) & 1) {                        // pad the cell to an even number of records
            L.slot_bin.push_back(k1 - 1); L.slot_real.push_back(0); L.slot_t.push_back(tt[k1 - 1]); L.slot_cell.push_back(cell);
        }
        k = k1;
    }
    // chunk size depends on the table ONLY, so the reduction tree (and every output bit) is independent of how many
    // walkers a call or a rank evaluates
    const long total = (long)L.slot_bin.size();
#ifndef BJ_TARGET_CHUNKS
#define BJ_TARGET_CHUNKS 128
#endif
    long target = BJ_TARGET_CHUNKS;
    if (const char* ev = getenv("BAJES_B200_CHUNKS")) target = std::max(1L, atol(ev));
    long per = (total + target - 1) / target;
    per = ((per + kTileBins - 1) / kTileBins) * kTileBins;
    if (per < 256) per = 256;
    L.per = (int)per;
    // the precise window [prec_lo, prec_hi) in slot space (multiples of 256 slots)
    L.prec_lo = L.prec_hi = 0;
    if (binweight && prec_frac > 0.0 && prec_share > 0.0 && total > 512) {
        std::vector<double> cum(total + 1, 0.0);
        for (long sidx = 0; sidx < total; ++sidx) {
            const double w = L.slot_real[sidx] ? (*binweight)[L.slot_bin[sidx]] : 0.0;
            cum[sidx + 1] = cum[sidx] + (std::isfinite(w) ? w * w : 0.0);
        }
        // shortest window with the requested share (two pointers), then rounded up to whole 256-slot chunks and capped
        const double want = std::fmin(prec_share, 1.0) * cum[total];
        long best_len = total, best_lo = 0, j = 0;
        for (long a = 0; a < total; a += 2) {
            while (j < total && cum[j] - cum[a] < want) ++j;
            if (cum[j] - cum[a] < want) break;
            if (j - a < best_len) { best_len = j - a; best_lo = a; }
        }
        long win = ((best_len + 255) / 256) * 256;
        const long cap = (long)std::ceil(prec_frac * (double)total / 256.0) * 256;
        if (win > cap) {                 // capped: the window of that length with the largest share
            win = cap;
            best_lo = 0;
            for (long a = 0; a + win <= total; a += 2)
                if (cum[a + win] - cum[a] > cum[best_lo + win] - cum[best_lo]) best_lo = a;
        }
        if (win > total) win = total & ~1L;
        if (best_lo + win > total) best_lo = (total - win) & ~1L;
        L.prec_lo = best_lo;
        L.prec_hi = best_lo + win;
        if (window_limit >= 0 && L.prec_hi > window_limit) {
            // compressed axis: the window stays inside the leading run of plain (uniformly spaced) bins
            L.prec_hi = (window_limit / 256) * 256;
            if (L.prec_lo > L.prec_hi) L.prec_lo = L.prec_hi;
            L.prec_lo = (L.prec_lo / 2) * 2;
            if ((L.prec_hi - L.prec_lo) % 256) L.prec_lo = L.prec_hi - ((L.prec_hi - L.prec_lo) / 256) * 256;
        }
    }
    std::vector<ChunkDesc> fast;
    long s0 = 0;
    while (s0 < total) {
        const int cell = L.slot_cell[s0];
        long s1 = s0;
        while (s1 < total && L.slot_cell[s1] == cell) ++s1;
        // pieces of this cell: before / inside / after the precise window
        const long cut[4] = {s0, std::min(std::max(L.prec_lo, s0), s1), std::min(std::max(L.prec_hi, s0), s1), s1};
        for (int piece = 0; piece < 3; ++piece) {
            const bool precise = piece == 1;
            const long step = precise ? kTileBins : per;      // one TMA tile per FP64 CTA: short serial chains, many CTAs
            for (long a = cut[piece]; a < cut[piece + 1]; a += step) {
                ChunkDesc cd;
                cd.start = (int)a;
                cd.len = (int)std::min(step, cut[piece + 1] - a);
                cd.band = L.cell_band[cell];
                cd.seg = L.cell_seg[cell];
                (precise ? L.chunks : fast).push_back(cd);
            }
        }
        s0 = s1;
    }
    L.n_precise = (int)L.chunks.size();
    L.chunks.insert(L.chunks.end(), fast.begin(), fast.end());
    return 0;
}

// Gram moments of <h|h> per detector, cell and t-weight: sum_k phi_m phi_n (4/T) f^(-7/3) / S_i * {(1-t)^2, t(1-t), t^2}
// (detector.py:539; without calibration only the first set, with weight 1); and (4/T) sum |d|^2/S per weight band
static void build_gram(std::vector<double>& gram, std::vector<double>& dd_band, const Layout& L, int nifo, int nsp,
                       int nw, const double* freqs, const double* const* data_ri, const double* const* psd,
                       double seglen) {
    const int ncell = (int)L.cell_band.size();
    std::vector<long double> acc((size_t)nifo * ncell * 3 * kGram, 0.0L);
    std::vector<long double> dd((size_t)nifo * (nw > 0 ? nw : 1), 0.0L);
    const double four_over_t = 4.0 / seglen;
    for (size_t sidx = 0; sidx < L.slot_bin.size(); ++sidx) {
        if (!L.slot_real[sidx]) continue;
        const long k = L.slot_bin[sidx];
        const int cell = L.slot_cell[sidx];
        const BinStatics b = bin_statics(freqs[k], seglen);
        const double u2 = b.u * b.u, u3 = u2 * b.u, u4 = u2 * u2, u5 = u4 * b.u, u6 = u3 * u3, u7 = u6 * b.u;
        const long double phi[kBasis] = {1.0L, u2, u3, u4, u5, u6, u7, (long double)b.L * u6};
        const long double t = L.slot_t[sidx];
        const long double tw[3] = {nsp > 0 ? (1 - t) * (1 - t) : 1.0L, nsp > 0 ? t * (1 - t) : 0.0L, nsp > 0 ? t * t : 0.0L};
        for (int i = 0; i < nifo; ++i) {
            const long double sw = (long double)four_over_t * (1.0 / psd[i][k]) * b.g * b.g;
            for (int st = 0; st < 3; ++st) {
                if (tw[st] == 0.0L) continue;
                long double* a = &acc[(((size_t)i * ncell + cell) * 3 + st) * kGram];
                int idx = 0;
                for (int m = 0; m < kBasis; ++m)
                    for (int q = m; q < kBasis; ++q, ++idx) a[idx] += phi[m] * phi[q] * sw * tw[st];
            }
            if (nw > 0) {
                const long double re = data_ri[i][2 * k], im = data_ri[i][2 * k + 1];
                dd[(size_t)i * nw + L.cell_band[cell]] += (long double)four_over_t * (re * re + im * im) / psd[i][k];
            }
        }
    }
    gram.resize(acc.size());
    for (size_t i = 0; i < acc.size(); ++i) gram[i] = (double)acc[i];
    dd_band.resize(dd.size());
    for (size_t i = 0; i < dd.size(); ++i) dd_band[i] = (double)dd[i];
}

// all-FP64 records (validation mode)
template <int NIFO>
static void build_records(std::vector<double>& rec, const Layout& L, const Axis& ax, double seglen) {
    constexpr int RS = Rec<NIFO>::kDoubles;
    const size_t n = L.slot_bin.size();
    rec.assign(n * RS, 0.0);
    for (size_t sidx = 0; sidx < n; ++sidx) {
        const long k = L.slot_bin[sidx];
        const BinStatics b = bin_statics(ax.f[k], seglen);
        double* r = &rec[sidx * RS];
        r[0] = b.u; r[1] = b.L; r[2] = b.w5; r[3] = b.f;
        if (!L.slot_real[sidx]) continue;
        for (int i = 0; i < NIFO; ++i) { r[4 + 2 * i] = ax.w[i][2 * k]; r[5 + 2 * i] = ax.w[i][2 * k + 1]; }
    }
}

// all-FP64 records of the precise window [lo, hi), data weights times the complex factor (sr + i si) that the epilogue's
// common <d|h> correction (table scale, calibrated phasor bias) will undo again
template <int NIFO>
static void build_records_window(std::vector<double>& rec, const Layout& L, long lo, long hi, const Axis& ax,
                                 double seglen, double sr, double si) {
    constexpr int RS = Rec<NIFO>::kDoubles;
    rec.assign((size_t)(hi - lo) * RS, 0.0);
    for (long sidx = lo; sidx < hi; ++sidx) {
        const long k = L.slot_bin[sidx];
        const BinStatics b = bin_statics(ax.f[k], seglen);
        double* r = &rec[(size_t)(sidx - lo) * RS];
        r[0] = b.u; r[1] = b.L; r[2] = b.w5; r[3] = b.f;
        if (!L.slot_real[sidx]) continue;
        for (int i = 0; i < NIFO; ++i) {
            const double wr = ax.w[i][2 * k], wi = ax.w[i][2 * k + 1];
            r[4 + 2 * i] = wr * sr - wi * si;
            r[5 + 2 * i] = wr * si + wi * sr;
        }
    }
}

// mixed-precision records: FP64 phase statics, FP32 amplitude statics and data weights scaled by 2^-e
template <int NIFO>
static int build_records_mixed(std::vector<unsigned char>& rec, const Layout& L, const Axis& ax, double seglen,
                               int* exp_out) {
    constexpr int RB = RecM<NIFO>::kBytes;
    const size_t n = L.slot_bin.size();
    rec.assign(n * RB, 0);
    const double* freqs = ax.f;
    double mx = 0.0;
    for (size_t sidx = 0; sidx < n; ++sidx) {
        if (!L.slot_real[sidx]) continue;
        const long k = L.slot_bin[sidx];
        for (int i = 0; i < NIFO; ++i) {
            const double wr = ax.w[i][2 * k], wi = ax.w[i][2 * k + 1];
            if (!std::isfinite(wr) || !std::isfinite(wi)) return -1;
            mx = std::fmax(mx, std::fmax(std::fabs(wr), std::fabs(wi)));
        }
    }
    int e = 0;
    if (mx > 0.0) std::frexp(mx, &e);       // mx = m 2^e, m in [0.5, 1)  ->  scaled weights below 1
    const double sc = std::ldexp(1.0, -e);
    for (size_t sidx = 0; sidx < n; ++sidx) {
        const long k = L.slot_bin[sidx];
        const BinStatics b = bin_statics(freqs[k], seglen);
        unsigned char* r = &rec[sidx * RB];
        double hd[4] = {b.u, b.L, b.w5, b.f};
        memcpy(r, hd, 32);
        // pair slots (records are consumed in (even, odd) pairs; cells and chunks start on even records):
        //   even record: (uf_even, uf_odd) and (t_even, t_odd)      odd record: (Lf_even, Lf_odd)
        const size_t se = sidx & ~(size_t)1, so = se + 1;
        const BinStatics be = bin_statics(freqs[L.slot_bin[se]], seglen);
        const BinStatics bo = bin_statics(freqs[L.slot_bin[so < n ? so : se]], seglen);
        float hf[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if ((sidx & 1) == 0) {
            hf[0] = (float)be.u; hf[1] = (float)bo.u;
            hf[2] = (float)L.slot_t[se]; hf[3] = (float)L.slot_t[so < n ? so : se];
        } else {
            hf[0] = (float)be.L; hf[1] = (float)bo.L;
        }
        memcpy(r + 32, hf, 16);
        for (int i = 0; i < NIFO; ++i) {
            double wr = 0.0, wi = 0.0;
            if (L.slot_real[sidx]) { wr = ax.w[i][2 * k]; wi = ax.w[i][2 * k + 1]; }
            const float fr = (float)(wr * sc), fi = (float)(wi * sc);
            float d4[4] = {fr, fi, fi, -fr};
            memcpy(r + 48 + 16 * i, d4, 16);
        }
    }
    *exp_out = e;
    return 0;
}

// tables of the tensor-core kernel (kernels.cuh: TileRec): basis[bin][K] -- the K basis values of the phase GEMM in the
// order of tile_basis_slot<MODEL>, products formed in long double from the same (u, L, w5) the Horner kernels read --
// and data[bin] = f, the FP32 pair slot and the scaled data weights of build_records_mixed (same 2^-e)
template <int MODEL>
static void build_records_tile(std::vector<double>& tab_basis, std::vector<unsigned char>& rec, const Layout& L,
                               const Axis& ax, double seglen, int nifo, int e2, bool rot) {
    using R = TileRec<MODEL>;
    constexpr int RB = R::kDataBytes;
    const size_t n = L.slot_bin.size();
    rec.assign(n * RB, 0);
    tab_basis.assign(n * R::K, 0.0);
    const double* freqs = ax.f;
    const double sc = std::ldexp(1.0, -e2);
    for (size_t sidx = 0; sidx < n; ++sidx) {
        const long k = L.slot_bin[sidx];
        const BinStatics b = bin_statics(freqs[k], seglen);
        unsigned char* r = &rec[sidx * RB];
        double* basis = &tab_basis[sidx * R::K];
        long double up[16];
        up[0] = 1.0L;
        for (int j = 1; j < 16; ++j) up[j] = up[j - 1] * (long double)b.u;
        const long double w5 = b.w5, Lg = b.L;
        if (MODEL == MODEL_35) {
            const int ex[9] = {0, 2, 3, 4, 5, 6, 7, 10, 12};
            for (int j = 0; j < 9; ++j) basis[j] = (double)(w5 * up[ex[j]]);
            basis[9] = (double)Lg;
            basis[10] = (double)(Lg * up[1]);
            basis[11] = 1.0;
        } else {
            for (int j = 0; j < 16; ++j) basis[j] = (double)(w5 * up[j]);
            for (int j = 0; j < 7; ++j) basis[16 + j] = (double)(Lg * up[j]);
            basis[23] = (double)(Lg * Lg * up[3]);
            basis[24] = 1.0;
        }
        memcpy(r + R::kF, &b.f, 8);
        const size_t se = sidx & ~(size_t)1, so = se + 1;
        const BinStatics be = bin_statics(freqs[L.slot_bin[se]], seglen);
        const BinStatics bo = bin_statics(freqs[L.slot_bin[so < n ? so : se]], seglen);
        float hf[2];
        if ((sidx & 1) == 0) { hf[0] = (float)be.u; hf[1] = (float)bo.u; }
        else                 { hf[0] = (float)be.L; hf[1] = (float)bo.L; }
        memcpy(r + R::kPair, hf, 8);
        const float tcal = (float)L.slot_t[sidx];
        memcpy(r + R::kT, &tcal, 4);
        float dri[BJ_MAX_IFO][2];
        for (int i = 0; i < nifo; ++i) {
            double wr = 0.0, wi = 0.0;
            if (L.slot_real[sidx]) { wr = ax.w[i][2 * k]; wi = ax.w[i][2 * k + 1]; }
            dri[i][0] = (float)(wr * sc);
            dri[i][1] = (float)(wi * sc);
        }
        if (rot && nifo == 3) {
            // rotation layout: (dr0, di0), (dr1, dr2), (di1, di2) -- the two rotated detectors share FFMA2 lanes
            const float d6[6] = {dri[0][0], dri[0][1], dri[1][0], dri[2][0], dri[1][1], dri[2][1]};
            memcpy(r + R::kData, d6, 24);
        } else {
            for (int i = 0; i < nifo; ++i) memcpy(r + R::kData + 8 * i, dri[i], 8);
        }
    }
}

template <int SC>
static int calibrate(double* eps2) {
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 2 * sizeof(double)));
    cudaMemset(d, 0, 2 * sizeof(double));
    calibrate_kernel<SC><<<256, 256>>>(d);
    cudaError_t e = cudaMemcpy(eps2, d, 2 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "phasor calibration failed: %s", cudaGetErrorString(e));
    return BJ_OK;
}

template <typename T>
static cudaError_t upload(T** dst, const std::vector<T>& src) {
    if (src.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, src.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static void adopt_layout(bj_ctx* c, const Layout& L) {
    c->n_precise = L.n_precise;
    c->prec_lo = L.prec_lo;
    c->prec_hi = L.prec_hi;
    c->n_cells = (int)L.cell_band.size();
    c->n_chunks = (int)L.chunks.size();
    c->bins_per_chunk = L.per;
    c->n_bins_kernel = (long)L.slot_bin.size();
}

// The mixed-precision table set of one frequency axis (the full grid, or the compressed node axis): layout, Horner
// records (fix-up / time marginalisation / amplitude correction), FP64 window, tensor-core tables; uploads them into c.
static int build_mixed_set(bj_ctx* c, const bj_config* cfg, const Axis& ax, const bj_extras* ex,
                           const std::vector<double>& bw, const double* eps, bool uniform_ok, bool tm, Layout& L,
                           long uniform_prefix = -1, double prefix_df = 0.0) {
    double share = 0.9, frac = 0.08;
    if (const char* ev = getenv("BAJES_B200_PRECISE_SHARE")) share = atof(ev);
    if (const char* ev = getenv("BAJES_B200_PRECISE_FRAC")) frac = atof(ev);
    if (c->shadow) {
        // on a compressed axis the plain low-frequency bins are a third to a half of the table, not 2 % of it: the window that
        // holds 90 % of their weight would double the cost.  16 % of the nodes (profiles/r02_compress_window.txt, 22 560 nodes:
        // worst walker in 2e6 0.42 tol at +7 % time; 8 %: 0.51, 40 %: 0.35 at +23 %; profiles/r02_compress_window_13312.txt,
        // the default 13 312 nodes, 4e6 walkers: 0.52 / 0.54 / 0.54 / 0.56 tol for 16 / 22 / 27 / 35 % -- the tail no longer
        // depends on it)
        frac = 0.16;
        if (const char* ev = getenv("BAJES_B200_CMP_PRECISE_SHARE")) share = atof(ev);
        if (const char* ev = getenv("BAJES_B200_CMP_PRECISE_FRAC")) frac = atof(ev);
    }
    if (tm || getenv("BAJES_B200_NO_TC")) frac = 0.0;
    if (build_layout(L, ax.n, ax.f, ex, &bw, share, frac, uniform_prefix)) return fail(BJ_ERR_INVALID, "len_weights do not cover the band");
    adopt_layout(c, L);
    if (uniform_prefix >= 0) c->window_df = prefix_df;
    std::vector<unsigned char> recm;
    int e2 = 0, bad = 0;
    switch (cfg->n_ifo) {
        case 1: bad = build_records_mixed<1>(recm, L, ax, cfg->seglen, &e2); c->rec_doubles = RecM<1>::kDoubles; break;
        case 2: bad = build_records_mixed<2>(recm, L, ax, cfg->seglen, &e2); c->rec_doubles = RecM<2>::kDoubles; break;
        default: bad = build_records_mixed<3>(recm, L, ax, cfg->seglen, &e2); c->rec_doubles = RecM<3>::kDoubles; break;
    }
    if (bad) return fail(BJ_ERR_INVALID, "non-finite static table entry (check data / PSD)");
    // computed phasor = true phasor * (1 + eps_a - i eps_p): divide it out, and undo the 2^-e table scale
    const double ar = 1.0 + eps[0], ai = -eps[1];
    const double den = ar * ar + ai * ai;
    const double sc = std::ldexp(1.0, e2);
    c->table_exp = e2;
    c->cfg.dh_fix_re = sc * ar / den;
    c->cfg.dh_fix_im = -sc * ai / den;
    if (c->n_precise > 0) {
        // the epilogue multiplies every chunk sum by 2^e / (ar + i ai): pre-multiply the FP64 window by its inverse
        std::vector<double> rec64;
        const double inv = std::ldexp(1.0, -e2);
        switch (cfg->n_ifo) {
            case 1: build_records_window<1>(rec64, L, L.prec_lo, L.prec_hi, ax, cfg->seglen, ar * inv, ai * inv); break;
            case 2: build_records_window<2>(rec64, L, L.prec_lo, L.prec_hi, ax, cfg->seglen, ar * inv, ai * inv); break;
            default: build_records_window<3>(rec64, L, L.prec_lo, L.prec_hi, ax, cfg->seglen, ar * inv, ai * inv); break;
        }
        cudaError_t e64 = upload(&c->d_rec64, rec64);
        if (e64 != cudaSuccess) return fail(BJ_ERR_CUDA, "uploading the FP64 window failed: %s", cudaGetErrorString(e64));
    }
    // the tensor-core kernel (phase as an FP64 GEMM) is the production path
    if (!getenv("BAJES_B200_NO_TC")) {
        // uniform in-band grid (every FFT grid bajes produces): integer time-delay ramps, and one phasor per bin
        // rotated to the other detectors
        if (uniform_ok && ax.n >= 2 && !getenv("BAJES_B200_NO_IRAMP")) {
            const double df = (ax.f[ax.n - 1] - ax.f[0]) / (double)(ax.n - 1);
            bool uniform = df > 0.0;
            for (long k = 0; k < ax.n && uniform; ++k)
                uniform = std::fabs(ax.f[k] - (ax.f[0] + (double)k * df)) <= 1e-7 * df;
            if (uniform) c->grid_df = df;
            if (uniform) c->window_df = df;
        }
        c->rot = c->grid_df > 0.0 && cfg->n_ifo >= 2 && !getenv("BAJES_B200_NO_ROT");
        if (c->rot) c->cfg.grid_df = c->grid_df;
        std::vector<unsigned char> rect;
        std::vector<double> basis;
        if (cfg->approx == BJ_APPROX_TAYLORF2_35PN) build_records_tile<MODEL_35>(basis, rect, L, ax, cfg->seglen, cfg->n_ifo, e2, c->rot);
        else                                        build_records_tile<MODEL_GENERIC>(basis, rect, L, ax, cfg->seglen, cfg->n_ifo, e2, c->rot);
        cudaError_t et = cudaMalloc(&c->d_tile_data, rect.size());
        if (et == cudaSuccess) et = cudaMemcpy(c->d_tile_data, rect.data(), rect.size(), cudaMemcpyHostToDevice);
        if (et == cudaSuccess) et = cudaMalloc(&c->d_tile_basis, basis.size() * sizeof(double));
        if (et == cudaSuccess) et = cudaMemcpy(c->d_tile_basis, basis.data(), basis.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (et != cudaSuccess) return fail(BJ_ERR_CUDA, "uploading the tensor-core tables failed: %s", cudaGetErrorString(et));
    }
    cudaError_t e = cudaMalloc(&c->d_rec, recm.size());
    if (e == cudaSuccess) e = cudaMemcpy(c->d_rec, recm.data(), recm.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = upload(&c->d_chunks, L.chunks);
    if (e == cudaSuccess) e = upload(&c->d_cell_band, L.cell_band);
    if (e == cudaSuccess) e = upload(&c->d_cell_seg, L.cell_seg);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "uploading static tables failed: %s", cudaGetErrorString(e));
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// Compressed frequency axis.
//
// <d|h>_i = sum_k w_ik g_i(f_k) with static weights w_ik = (4/T) conj(d_ik)/S_ik f_k^(-7/6) e^{i pi f_k T} and the walker's
// SMOOTH factor g_i(f) = Q(f) e^{-2 pi i [Phi(f)/2pi + f tau_i]}.  On a run of bins [a, b] over which the phase of g turns
// by at most `reach` radians to either side of the centre, g is reproduced to ~1e-12 by its Chebyshev interpolant on n
// nodes f_j, so
//     sum_{k in run} w_ik g_i(f_k) = sum_j [ sum_k w_ik l_j(f_k) ] g_i(f_j) = sum_j W_ij g_i(f_j)
// with the Lagrange cardinal functions l_j: the run's bins collapse into n "bins" with weights W_ij that the very same
// kernels evaluate (non-uniform axis).  The run length follows from a DESIGN bound on the phase slope,
//     |d(Phi/2pi + f tau)/df| <= margin t_N(f; mchirp_min) + tau_max        (t_N: Newtonian time to coalescence),
// that steep_kernel checks per walker; walkers outside the design are evaluated on the full grid (run_device).  Where a run
// would hold fewer than 1.5 n bins (low frequencies of a long signal) the bins are kept as they are.  Weights are formed in
// long double with the barycentric formula; nothing about the walker enters, and <h|h> never used the grid (Gram moments).
// ------------------------------------------------------------------------------------------------
struct CompressSpec {
    int on, nodes;
    double reach, mchirp_min, tau_max, margin;
    int focus = 0;       // budget relative to a fiducial point (bj_focus): a (f / f0)^(-5/3) + b + c (f / f1)^(2/3)
    double focus_a = 0.0, focus_b = 0.0, focus_c = 0.0, focus_f0 = 1.0, focus_f1 = 1.0;
};
static CompressSpec compress_spec(const bj_extras* ex) {
    // 64 nodes per run of +-32 rad: the Chebyshev truncation error 2 (reach/2)^n / n! is ~1e-12 there, and longer runs leave
    // fewer nodes (C2: 17 728; 32 nodes of +-10 rad: 22 560) at the same worst-walker error (profiles/r02_compress_nodes.txt)
    CompressSpec sp;
    sp.on = 1; sp.nodes = 64; sp.reach = 32.0; sp.mchirp_min = 0.8; sp.tau_max = 0.11; sp.margin = 1.3;
    if (ex && ex->compress != 0) {
        sp.on = ex->compress > 0;
        if (ex->compress_mchirp_min > 0.0) sp.mchirp_min = ex->compress_mchirp_min;
        if (ex->compress_tau_max > 0.0) sp.tau_max = ex->compress_tau_max;
    }
    if (const char* ev = getenv("BAJES_B200_COMPRESS")) sp.on = atoi(ev) != 0;
    if (const char* ev = getenv("BAJES_B200_COMPRESS_NODES")) sp.nodes = std::max(4, std::min(64, atoi(ev)));
    if (const char* ev = getenv("BAJES_B200_COMPRESS_REACH")) sp.reach = atof(ev);
    if (const char* ev = getenv("BAJES_B200_COMPRESS_MCHIRP_MIN")) sp.mchirp_min = atof(ev);
    if (const char* ev = getenv("BAJES_B200_COMPRESS_TAU_MAX")) sp.tau_max = atof(ev);
    return sp;
}
// seconds: the design bound on |d(Phi/2pi)/df| + |tau_i| at frequency f  (gt: taylorf2.py:6)
static double design_budget(const CompressSpec& sp, double f) {
    if (sp.focus)
        return sp.focus_a * std::pow(f / sp.focus_f0, -5.0 / 3.0) + sp.focus_b + sp.focus_c * std::pow(f / sp.focus_f1, 2.0 / 3.0);
    const double t = 5.0 / 256.0 * std::pow(sp.mchirp_min * 4.92549094830932e-6, -5.0 / 3.0) * std::pow(M_PI * f, -8.0 / 3.0);
    return sp.margin * t + sp.tau_max;
}
static void build_compressed_axis(const Axis& full, int nifo, const std::vector<double>& bw, const CompressSpec& sp,
                                  std::vector<double>& nf, std::vector<double>* nw, std::vector<double>& nbw, int* info) {
    const int n = sp.nodes;
    const long N = full.n;
    nf.clear();
    nbw.clear();
    for (int i = 0; i < nifo; ++i) nw[i].clear();
    std::vector<long double> xn(n), lam(n), ell(n);
    for (int j = 0; j < n; ++j) {
        const long double th = M_PIl * (2 * j + 1) / (2.0L * n);
        xn[j] = -cosl(th);                               // ascending
        lam[j] = ((j & 1) ? -1.0L : 1.0L) * sinl(th);    // barycentric weights of the first-kind Chebyshev points
    }
    std::vector<long double> acc((size_t)2 * nifo * n), accb(n);
    long k = 0, kept = 0, blocks = 0;
    while (k < N) {
        const double width = 2.0 * sp.reach / (2.0 * M_PI * design_budget(sp, full.f[k]));    // Hz
        long k1 = k + 1;
        while (k1 < N && full.f[k1] - full.f[k] <= width) ++k1;
        const long nb = k1 - k;
        if (nb <= n + n / 2) {                           // too few bins to gain: keep n of them as they are
            const long ke = std::min(N, k + (long)n);
            for (; k < ke; ++k, ++kept) {
                nf.push_back(full.f[k]);
                nbw.push_back(bw[k]);
                for (int i = 0; i < nifo; ++i) { nw[i].push_back(full.w[i][2 * k]); nw[i].push_back(full.w[i][2 * k + 1]); }
            }
            continue;
        }
        const long double a = full.f[k], b = full.f[k1 - 1];
        const long double mid = 0.5L * (a + b), half = 0.5L * (b - a);
        std::fill(acc.begin(), acc.end(), 0.0L);
        std::fill(accb.begin(), accb.end(), 0.0L);
        for (long q = k; q < k1; ++q) {
            const long double x = ((long double)full.f[q] - mid) / half;
            int hit = -1;
            long double den = 0.0L;
            for (int j = 0; j < n; ++j) {
                const long double d = x - xn[j];
                if (d == 0.0L) { hit = j; break; }
                ell[j] = lam[j] / d;
                den += ell[j];
            }
            if (hit >= 0) { for (int j = 0; j < n; ++j) ell[j] = j == hit ? 1.0L : 0.0L; }
            else          { for (int j = 0; j < n; ++j) ell[j] /= den; }
            for (int i = 0; i < nifo; ++i) {
                const long double wr = full.w[i][2 * q], wi = full.w[i][2 * q + 1];
                long double* ai = &acc[(size_t)2 * i * n];
                for (int j = 0; j < n; ++j) { ai[2 * j] += wr * ell[j]; ai[2 * j + 1] += wi * ell[j]; }
            }
            for (int j = 0; j < n; ++j) accb[j] += (long double)bw[q] * fabsl(ell[j]);
        }
        for (int j = 0; j < n; ++j) {
            nf.push_back((double)(mid + half * xn[j]));
            nbw.push_back((double)accb[j]);
            for (int i = 0; i < nifo; ++i) { nw[i].push_back((double)acc[(size_t)2 * (i * n + j)]); nw[i].push_back((double)acc[(size_t)2 * (i * n + j) + 1]); }
        }
        ++blocks;
        k = k1;
    }
    info[0] = (int)nf.size();
    info[1] = (int)kept;
    info[2] = (int)blocks;
    info[3] = n;
}

extern "C" int bj_create_ex(bj_ctx** out, const bj_config* cfg, int64_t n_bins, const double* freqs,
                            const double* const* data_ri, const double* const* psd, const bj_extras* ex) {
    if (!out) return fail(BJ_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int rc = check_config(cfg);
    if (rc) return rc;
    if (n_bins <= 0 || !freqs || !data_ri || !psd) return fail(BJ_ERR_INVALID, "empty frequency axis or null tables");
    for (int64_t k = 0; k < n_bins; ++k) {
        if (!(freqs[k] > 0.0) || !std::isfinite(freqs[k]))
            return fail(BJ_ERR_INVALID, "freqs[%lld]=%g: in-band frequencies must be positive and finite", (long long)k,
                        freqs[k]);
        if (k > 0 && !(freqs[k] > freqs[k - 1])) return fail(BJ_ERR_INVALID, "in-band frequencies must be ascending");
    }
    const int nsp = ex ? ex->nspcal : 0, nw = ex ? ex->nweights : 0;
    if (nsp < 0 || nsp == 1 || nsp > BJ_MAX_SPCAL) return fail(BJ_ERR_INVALID, "nspcal=%d outside {0, 2..%d}", nsp, BJ_MAX_SPCAL);
    if (nw < 0 || nw > BJ_MAX_WEIGHTS) return fail(BJ_ERR_INVALID, "nweights=%d outside 0..%d", nw, BJ_MAX_WEIGHTS);
    if (nsp > 0) {
        if (!ex->spcal_freqs) return fail(BJ_ERR_INVALID, "null spcal_freqs");
        for (int j = 1; j < nsp; ++j)
            if (!(ex->spcal_freqs[j] > ex->spcal_freqs[j - 1])) return fail(BJ_ERR_INVALID, "spcal_freqs must be ascending");
    }
    if (nw > 0) {
        if (!ex->len_weights) return fail(BJ_ERR_INVALID, "null len_weights");
        int64_t tot = 0;
        for (int b = 0; b < nw; ++b) {
            if (ex->len_weights[b] < 0) return fail(BJ_ERR_INVALID, "negative len_weights[%d]", b);
            tot += ex->len_weights[b];
        }
        if (tot != n_bins) return fail(BJ_ERR_INVALID, "sum(len_weights)=%lld differs from n_bins=%lld", (long long)tot, (long long)n_bins);
    }
    bj_ctx* c = new bj_ctx();
    rc = init_common(c, cfg);
    if (rc) { delete c; return rc; }
    c->kind = KIND_FULLGRID;
    c->n_bins = n_bins;
    c->dims.nspcal = nsp;
    c->dims.nweights = nw;
    for (int b = 0; b < nw; ++b) c->dims.len_weights[b] = (int)ex->len_weights[b];
    c->cfg.f_lo = freqs[0];
    c->cfg.f_hi = freqs[n_bins - 1];
    if (ex && ex->marg_time_shift) {
        if (nsp || nw) { bj_destroy(c); return fail(BJ_ERR_UNSUPPORTED, "time-shift marginalisation with spcal / PSD weights"); }
        if (cfg->sincos == BJ_SINCOS_FP64) { bj_destroy(c); return fail(BJ_ERR_UNSUPPORTED, "time-shift marginalisation uses the mixed-precision tables: choose sincos mufu or poly32"); }
        if (ex->n_full < ex->band_offset + n_bins || ex->band_offset < 0) { bj_destroy(c); return fail(BJ_ERR_INVALID, "band [%lld, %lld) does not fit the full axis of %lld samples", (long long)ex->band_offset, (long long)(ex->band_offset + n_bins), (long long)ex->n_full); }
        if (!g_cufft.load()) { bj_destroy(c); return fail(BJ_ERR_UNSUPPORTED, "time-shift marginalisation needs libcufft.so.11, which could not be loaded"); }
        c->tm = 1;
        c->tm_n_full = (long)ex->n_full;
        c->tm_offset = (long)ex->band_offset;
    }

    // per-bin weights of the full grid, shared by every table builder
    std::vector<double> wts[BJ_MAX_IFO];
    full_axis_weights(wts, cfg->n_ifo, (long)n_bins, freqs, data_ri, psd, cfg->seglen);
    Axis ax;
    ax.n = (long)n_bins;
    ax.f = freqs;
    for (int i = 0; i < cfg->n_ifo; ++i) ax.w[i] = wts[i].data();
    // share of each bin in <h|h> of an inspiral; the precise window only exists in the mixed-precision modes
    std::vector<double> bw((size_t)n_bins, 0.0);
    for (int64_t k = 0; k < n_bins; ++k) {
        double w = 0.0;
        for (int i = 0; i < cfg->n_ifo; ++i) w += 1.0 / psd[i][k];
        bw[k] = w * std::pow(freqs[k], -7.0 / 3.0);
    }
    const bool tm = ex && ex->marg_time_shift;

    Layout L;
    if (cfg->sincos == BJ_SINCOS_FP64) {
        if (build_layout(L, n_bins, freqs, ex, &bw, 0.0, 0.0)) { bj_destroy(c); return fail(BJ_ERR_INVALID, "len_weights do not cover the band"); }
        adopt_layout(c, L);
        std::vector<double> rec;
        switch (cfg->n_ifo) {
            case 1: build_records<1>(rec, L, ax, cfg->seglen); c->rec_doubles = Rec<1>::kDoubles; break;
            case 2: build_records<2>(rec, L, ax, cfg->seglen); c->rec_doubles = Rec<2>::kDoubles; break;
            default: build_records<3>(rec, L, ax, cfg->seglen); c->rec_doubles = Rec<3>::kDoubles; break;
        }
        for (size_t i = 0; i < rec.size(); ++i)
            if (!std::isfinite(rec[i])) {
                bj_destroy(c);
                return fail(BJ_ERR_INVALID, "non-finite static table entry at record %zu (check data / PSD)", i / c->rec_doubles);
            }
        cudaError_t e = cudaMalloc(&c->d_rec, rec.size() * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(c->d_rec, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = upload(&c->d_chunks, L.chunks);
        if (e == cudaSuccess) e = upload(&c->d_cell_band, L.cell_band);
        if (e == cudaSuccess) e = upload(&c->d_cell_seg, L.cell_seg);
        if (e != cudaSuccess) { bj_destroy(c); return fail(BJ_ERR_CUDA, "uploading static tables failed: %s", cudaGetErrorString(e)); }
    } else {
        double eps[2] = {0.0, 0.0};
        rc = cfg->sincos == BJ_SINCOS_MUFU ? calibrate<BJ_SINCOS_MUFU>(eps) : calibrate<BJ_SINCOS_POLY32>(eps);
        if (rc) { bj_destroy(c); return rc; }
        c->calib[0] = eps[0];
        c->calib[1] = eps[1];
        rc = build_mixed_set(c, cfg, ax, ex, bw, eps, /*uniform_ok=*/true, tm, L);
        if (rc) { bj_destroy(c); return rc; }
    }

    std::vector<double> gram, dd_band;
    build_gram(gram, dd_band, L, cfg->n_ifo, nsp, nw, freqs, data_ri, psd, cfg->seglen);
    for (size_t i = 0; i < gram.size(); ++i)
        if (!std::isfinite(gram[i])) {
            bj_destroy(c);
            return fail(BJ_ERR_INVALID, "non-finite <h|h> moment (check the PSD)");
        }
    cudaError_t e = upload(&c->d_gram, gram);
    if (e == cudaSuccess) e = upload(&c->d_dd_band, dd_band);
    if (e == cudaSuccess && nsp > 0) {
        std::vector<double> nf(ex->spcal_freqs, ex->spcal_freqs + nsp);
        e = upload(&c->d_spcal_freqs, nf);
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->d_freqs, (size_t)n_bins * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(c->d_freqs, freqs, (size_t)n_bins * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        bj_destroy(c);
        return fail(BJ_ERR_CUDA, "uploading static tables failed: %s", cudaGetErrorString(e));
    }

    // ---- compressed frequency axes (mixed-precision modes, tensor-core tables, no extras, no time marginalisation) ----
    CompressSpec sp = compress_spec(ex);
    if (sp.on && cfg->sincos != BJ_SINCOS_FP64 && c->d_tile_basis && c->grid_df > 0.0 && nsp == 0 && nw == 0 && !tm) {
        // level 0: the design; level 1: time shifts up to ~1 s (the default prior of the reference's pipeline,
        // pipe/scripts/bajes_pipe:526) -- samplers start there and end in level 0
        // and lighter systems: half the design's chirp mass
        double taus[kMaxLevels] = {sp.tau_max, std::fmax(1.1, 2.0 * sp.tau_max)};
        double mcs[kMaxLevels] = {sp.mchirp_min, 0.5 * sp.mchirp_min};
        int want = sp.tau_max >= 1.1 ? 1 : kMaxLevels;
        if (const char* ev = getenv("BAJES_B200_COMPRESS_LEVELS")) want = std::max(1, std::min(kMaxLevels, atoi(ev)));
        for (int lv = 0; lv < want; ++lv) {
            CompressSpec sl = sp;
            sl.tau_max = taus[lv];
            sl.mchirp_min = mcs[lv];
            std::vector<double> nf, nbw, nwt[BJ_MAX_IFO];
            int info[4] = {0, 0, 0, 0};
            build_compressed_axis(ax, cfg->n_ifo, bw, sl, nf, nwt, nbw, info);
            // worth it only on long axes (a call on a short one is bound by its launches, and a compressed context
            // launches more of them) and when the node axis is markedly shorter than the grid
            long min_bins = 65536;
            if (const char* ev = getenv("BAJES_B200_COMPRESS_MIN_BINS")) min_bins = atol(ev);
            if (n_bins < min_bins || (double)nf.size() > (double)n_bins / 3.0) break;
            Axis cx;
            cx.n = (long)nf.size();
            cx.f = nf.data();
            for (int i = 0; i < cfg->n_ifo; ++i) cx.w[i] = nwt[i].data();
            bj_ctx* m = new bj_ctx();
            m->kind = KIND_FULLGRID;
            m->shadow = true;
            m->device = c->device;
            m->sm_count = c->sm_count;
            m->sincos = c->sincos;
            m->cfg = c->cfg;
            memset(&m->dims, 0, sizeof(m->dims));
            m->n_bins = cx.n;
            c->cmp[c->n_levels] = m;
            Layout Lc;
            // the leading run of plain bins is a piece of the (uniform) grid: the FP64 window is placed there and
            // evaluated by the fast window kernel
            long prefix = 0;
            while (prefix < cx.n && prefix < (long)n_bins && nf[prefix] == freqs[prefix]) ++prefix;
            if (!(c->grid_df > 0.0)) prefix = -1;
            if (prefix >= 0) for (long q = prefix; q < cx.n; ++q) nbw[q] = 0.0;     // the window is chosen among the plain bins
            rc = build_mixed_set(m, cfg, cx, nullptr, nbw, c->calib, /*uniform_ok=*/false, false, Lc, prefix, c->grid_df);
            if (rc) { c->n_levels++; bj_destroy(c); return rc; }
            const int q0 = c->n_levels++;
            for (int q = 0; q < 4; ++q) c->cmp_info[q0][q] = info[q];
            c->cmp_spec[q0][0] = sl.mchirp_min; c->cmp_spec[q0][1] = sl.tau_max; c->cmp_spec[q0][2] = sl.reach; c->cmp_spec[q0][3] = sl.margin;
            for (int q = 0; q < kSteepChecks; ++q) {
                const double f = freqs[0] * std::pow(freqs[n_bins - 1] / freqs[0], (double)q / (double)(kSteepChecks - 1));
                c->steep.budget[q0][q] = design_budget(sl, f);
            }
        }
        if (c->n_levels > 0 && c->rot) {
            // launch_multi evaluates every table with the per-detector record layout: a second copy of the grid's data
            // records (the rotation layout interleaves the two rotated detectors)
            std::vector<unsigned char> rect;
            std::vector<double> basis;
            if (cfg->approx == BJ_APPROX_TAYLORF2_35PN) build_records_tile<MODEL_35>(basis, rect, L, ax, cfg->seglen, cfg->n_ifo, c->table_exp, false);
            else                                        build_records_tile<MODEL_GENERIC>(basis, rect, L, ax, cfg->seglen, cfg->n_ifo, c->table_exp, false);
            cudaError_t et = cudaMalloc(&c->d_tile_data_plain, rect.size());
            if (et == cudaSuccess) et = cudaMemcpy(c->d_tile_data_plain, rect.data(), rect.size(), cudaMemcpyHostToDevice);
            if (et != cudaSuccess) { bj_destroy(c); return fail(BJ_ERR_CUDA, "uploading the per-detector data records failed: %s", cudaGetErrorString(et)); }
        }
        if (c->n_levels > 0) {
            c->h_freqs.assign(freqs, freqs + n_bins);
            for (int i = 0; i < cfg->n_ifo; ++i) c->h_wts[i] = wts[i];
        }
        if (c->n_levels > 0) {
            // check frequencies of classify_kernel: log-spaced over the band, both ends included
            c->steep.n = kSteepChecks;
            c->steep.n_levels = c->n_levels;
            for (int q = 0; q < kSteepChecks; ++q) {
                const double f = freqs[0] * std::pow(freqs[n_bins - 1] / freqs[0], (double)q / (double)(kSteepChecks - 1));
                c->steep.u[q] = std::cbrt(f);
                c->steep.L[q] = std::log(f);
                c->steep.w5[q] = std::pow(f, -5.0 / 3.0);
                c->steep.inv_f[q] = 1.0 / f;
            }
            c->steep.focus = 0;
            if (cudaMalloc(&c->d_counts, (kMaxLevels + 2) * sizeof(int)) != cudaSuccess) {
                bj_destroy(c);
                return fail(BJ_ERR_CUDA, "allocating the work-list counters failed");
            }
        }
    }
    *out = c;
    return BJ_OK;
}

extern "C" int bj_create(bj_ctx** out, const bj_config* cfg, int64_t n_bins, const double* freqs,
                         const double* const* data_ri, const double* const* psd) {
    return bj_create_ex(out, cfg, n_bins, freqs, data_ri, psd, nullptr);
}

extern "C" int bj_destroy(bj_ctx* c) {
    if (!c) return BJ_OK;
    cudaSetDevice(c->device);
    for (int lv = 0; lv < kMaxLevels; ++lv) {
        bj_ctx* m = c->cmp[lv];
        if (!m) continue;
        cudaFree(m->d_rec);
        cudaFree(m->d_tile_basis);
        cudaFree(m->d_tile_data);
        cudaFree(m->d_rec64);
        cudaFree(m->d_chunks);
        cudaFree(m->d_cell_band);
        cudaFree(m->d_cell_seg);
        delete m;
        c->cmp[lv] = nullptr;
    }
    if (c->focus) {
        cudaFree(c->focus->d_rec);
        cudaFree(c->focus->d_chunks);
        delete c->focus;
        c->focus = nullptr;
    }
    cudaFree(c->d_fid);
    cudaFree(c->d_lists);
    cudaFree(c->d_counts);
    cudaFree(c->d_rec);
    cudaFree(c->d_tile_basis);
    cudaFree(c->d_tile_data);
    cudaFree(c->d_tile_data_plain);
    cudaFree(c->d_rec64);
    cudaFree(c->d_freqs);
    cudaFree(c->d_gram);
    cudaFree(c->d_chunks);
    cudaFree(c->d_cell_band);
    cudaFree(c->d_cell_seg);
    cudaFree(c->d_dd_band);
    cudaFree(c->d_spcal_freqs);
    cudaFree(c->d_coef_ex);
    cudaFree(c->d_tm);
    cudaFree(c->d_tm_R);
    if (c->tm_plan_ok) g_cufft.Destroy(c->tm_plan);
    if (c->tm_small > 0) g_cufft.Destroy(c->tm_plan_small);
    cudaFree(c->d_coef);
    cudaFree(c->d_partial);
    cudaFree(c->d_x);
    cudaFree(c->d_logl);
    c->roq.release();
    c->relbin.release();
    for (int i = 0; i < 4; ++i)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; ++i)
        if (c->ev_tile[i]) cudaEventDestroy(c->ev_tile[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BJ_OK;
}

static long pad_walkers(long w) { return ((w + 127) / 128) * 128; }

static int ensure_workspace(bj_ctx* c, long n_walkers) {
    if (n_walkers <= c->cap_walkers) return BJ_OK;
    long cap = pad_walkers(n_walkers);
    if (cap < 2 * c->cap_walkers && 2 * c->cap_walkers >= n_walkers) cap = pad_walkers(2 * c->cap_walkers);
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(c->d_coef); c->d_coef = nullptr;
    cudaFree(c->d_partial); c->d_partial = nullptr;
    cudaFree(c->d_coef_ex); c->d_coef_ex = nullptr;
    c->cap_walkers = 0;
    CUDA_TRY(cudaMalloc(&c->d_coef, (size_t)C_NUM * cap * sizeof(double)));
    if (c->dims.nspcal > 0 || c->dims.nweights > 0)
        CUDA_TRY(cudaMalloc(&c->d_coef_ex, (size_t)ex_num_slots(c->dims, c->cfg.n_ifo) * cap * sizeof(double)));
    size_t part = 0;
    if (c->kind == KIND_FULLGRID) {
        int nch = c->n_chunks;
        for (int lv = 0; lv < c->n_levels; ++lv) nch = std::max(nch, c->cmp[lv]->n_chunks);
        if (c->focus) nch = std::max(nch, c->focus->n_chunks);
        part = (size_t)nch * 2 * c->cfg.n_ifo * cap;
    }
    if (c->n_levels > 0) {
        cudaFree(c->d_lists); c->d_lists = nullptr;
        CUDA_TRY(cudaMalloc(&c->d_lists, (size_t)(c->n_levels + 2) * cap * sizeof(int)));
    }
    if (part) CUDA_TRY(cudaMalloc(&c->d_partial, part * sizeof(double)));
    c->cap_walkers = cap;
    return BJ_OK;
}

static void to_dev_map(const bj_param_map* m, ParamMapDev& d) {
    for (int i = 0; i < BJ_NUM_PARAMS; ++i) {
        d.column[i] = m->column[i];
        d.value[i] = m->value[i];
    }
}

static int check_map(const bj_param_map* m, int ndim) {
    if (!m) return fail(BJ_ERR_INVALID, "null parameter map");
    for (int i = 0; i < BJ_NUM_PARAMS; ++i) {
        int col = m->column[i];
        if (col >= ndim || col < BJ_COL_ABSENT) return fail(BJ_ERR_INVALID, "parameter slot %d maps to column %d (ndim=%d)", i, col, ndim);
    }
    auto present = [&](int s) { return m->column[s] != BJ_COL_ABSENT; };
    if (!present(BJ_P_Q)) return fail(BJ_ERR_INVALID, "parameter 'q' is required");
    if (!present(BJ_P_MCHIRP) && !present(BJ_P_MTOT)) return fail(BJ_ERR_INVALID, "one of 'mchirp' / 'mtot' is required");
    if (!present(BJ_P_COS_IOTA) && !present(BJ_P_IOTA)) return fail(BJ_ERR_INVALID, "one of 'cos_iota' / 'iota' is required");
    const int need[] = {BJ_P_DISTANCE, BJ_P_RA, BJ_P_DEC, BJ_P_PSI, BJ_P_TIME_SHIFT, BJ_P_PHI_REF, BJ_P_T_GPS};
    for (int s : need)
        if (!present(s)) return fail(BJ_ERR_INVALID, "required parameter slot %d is absent", s);
    return BJ_OK;
}

static void note_launch(bj_ctx* c, dim3 grid, size_t smem, int rec8, int model, int threads = kThreads) {
    c->launch_info[0] = (int)grid.x;
    c->launch_info[1] = (int)grid.y;
    c->launch_info[2] = threads;
    c->launch_info[3] = (int)smem;
    c->launch_info[4] = c->bins_per_chunk;
    c->launch_info[5] = c->n_chunks;
    c->launch_info[6] = rec8;
    c->launch_info[7] = model;
}

// sel (kernels.cuh: Sel): the work list the launch evaluates (contexts with compressed axes: one list per table); the grid
// is sized for the whole batch and CTAs beyond the end of the list return at once.  {null, null} = every walker.
template <int NIFO, int MODEL, bool SPCAL>
static int launch_fullgrid_fp64(bj_ctx* c, long n_walkers, cudaStream_t st, int c0, int c1, Sel sel) {
    constexpr int RS = Rec<NIFO>::kDoubles;
    const size_t smem = (size_t)kStages * kTileBins * RS * 8 + kStages * sizeof(uint64_t);
    auto kern = fullgrid_kernel<NIFO, MODEL, BJ_SINCOS_FP64, SPCAL>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(c1 - c0), (unsigned)((n_walkers + kThreads - 1) / kThreads));
    kern<<<grid, kThreads, smem, st>>>((const double*)c->d_rec, c->d_chunks, c->d_coef, c->cap_walkers, n_walkers,
                                       c->d_partial, c->d_coef_ex, c->dims, c->d_spcal_freqs, c0, sel, nullptr);
    CUDA_TRY(cudaGetLastError());
    note_launch(c, grid, smem, RS, MODEL);
    return BJ_OK;
}

template <int NIFO, int MODEL, int SC, bool SPCAL>
static int launch_fullgrid_mixed(bj_ctx* c, long n_walkers, cudaStream_t st, int c0, int c1, Sel sel) {
    constexpr int RB = RecM<NIFO>::kBytes;
    if (c0 < c->n_precise) {
        // the heaviest bins in FP64 (chunks of 64 records; the table pointer is offset so that chunk.start indexes it)
        constexpr int RS = Rec<NIFO>::kDoubles;
        const int p1 = c1 < c->n_precise ? c1 : c->n_precise;
        const size_t smem64 = (size_t)kStages * kTileBins * RS * 8 + kStages * sizeof(uint64_t);
        dim3 pgrid((unsigned)(p1 - c0), (unsigned)((n_walkers + kThreads - 1) / kThreads));
        if (c->window_df > 0.0 && !getenv("BAJES_B200_WINDOW_GENERIC")) {
            auto pk = fullgrid_window_kernel<NIFO, MODEL, SPCAL>;
            CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
            pk<<<pgrid, kThreads, smem64, st>>>(c->d_rec64 - (size_t)c->prec_lo * RS, c->d_chunks, c->d_coef,
                                                c->cap_walkers, n_walkers, c->d_partial, c->d_coef_ex, c->dims,
                                                c->d_spcal_freqs, c0, c->window_df, sel, MultiTable{});
        } else {
            auto pk = fullgrid_kernel<NIFO, MODEL, BJ_SINCOS_FP64, SPCAL>;
            CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
            pk<<<pgrid, kThreads, smem64, st>>>(c->d_rec64 - (size_t)c->prec_lo * RS, c->d_chunks, c->d_coef,
                                                c->cap_walkers, n_walkers, c->d_partial, c->d_coef_ex, c->dims,
                                                c->d_spcal_freqs, c0, sel, nullptr);
        }
        CUDA_TRY(cudaGetLastError());
        c0 = p1;
        if (c0 >= c1) return BJ_OK;
    }
    const size_t smem = (size_t)kStages * kTileBins * RB + kStages * sizeof(uint64_t);
    auto kern = fullgrid_mixed_kernel<NIFO, MODEL, SC, false, SPCAL>;
    auto fixup = fullgrid_mixed_kernel<NIFO, MODEL, SC, true, SPCAL>;
    CUDA_TRY(cudaFuncSetAttribute(fixup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(c1 - c0), (unsigned)((n_walkers + kThreads - 1) / kThreads));
    if (c->d_tile_basis) {
        constexpr int RT = TileRec<MODEL>::kBasisBytes + TileRec<MODEL>::kDataBytes;
        const size_t smem_t = (size_t)kStagesT * TileRec<MODEL>::kTileBins * RT + kStagesT * sizeof(uint64_t);
        auto tkern = c->grid_df > 0.0 ? fullgrid_tile_kernel<NIFO, MODEL, SC, true, SPCAL, false>
                                      : fullgrid_tile_kernel<NIFO, MODEL, SC, false, SPCAL, false>;
        if constexpr (NIFO >= 2) {
            if (c->rot) tkern = fullgrid_tile_kernel<NIFO, MODEL, SC, true, SPCAL, true>;
        }
        CUDA_TRY(cudaFuncSetAttribute(tkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
        dim3 tgrid((unsigned)(c1 - c0), (unsigned)((n_walkers + kTileWalkers - 1) / kTileWalkers));
        if (!c->shadow) { CUDA_TRY(cudaEventRecord(c->ev_tile[0], st)); }
        tkern<<<tgrid, kTileThreads, smem_t, st>>>((const unsigned char*)c->d_tile_basis, (const unsigned char*)c->d_tile_data,
                                                   (const unsigned char*)c->d_rec, RB, c->d_chunks, c->d_coef,
                                                   c->cap_walkers, n_walkers, c->d_partial, c0, c->grid_df,
                                                   c->d_coef_ex, c->dims, sel, MultiTable{});
        if (!c->shadow) { CUDA_TRY(cudaEventRecord(c->ev_tile[1], st)); c->tile_timed = true; }
        note_launch(c, tgrid, smem_t, RT / 8, MODEL, kTileThreads);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kThreads, smem, st>>>((const unsigned char*)c->d_rec, c->d_chunks, c->d_coef, c->cap_walkers,
                                           n_walkers, c->d_partial, c->d_coef_ex, c->dims, c0);
        note_launch(c, grid, smem, RB / 8, MODEL);
    }
    // wide-phase fix-up: every CTA returns immediately unless the prologue flagged one of its walkers.  (Compressed
    // tables never see such walkers: classify_kernel sends them to the full grid.)
    if (c->shadow) return BJ_OK;
    fixup<<<grid, kThreads, smem, st>>>((const unsigned char*)c->d_rec, c->d_chunks, c->d_coef, c->cap_walkers,
                                        n_walkers, c->d_partial, c->d_coef_ex, c->dims, c0);
    CUDA_TRY(cudaGetLastError());
    return BJ_OK;
}

template <int NIFO, int MODEL, bool SPCAL>
static int launch_fullgrid_s(bj_ctx* c, long n_walkers, cudaStream_t st, int c0, int c1, Sel sel) {
    switch (c->sincos) {
        case BJ_SINCOS_MUFU: return launch_fullgrid_mixed<NIFO, MODEL, BJ_SINCOS_MUFU, SPCAL>(c, n_walkers, st, c0, c1, sel);
        case BJ_SINCOS_POLY32: return launch_fullgrid_mixed<NIFO, MODEL, BJ_SINCOS_POLY32, SPCAL>(c, n_walkers, st, c0, c1, sel);
        default: return launch_fullgrid_fp64<NIFO, MODEL, SPCAL>(c, n_walkers, st, c0, c1, sel);
    }
}

template <int NIFO, int MODEL>
static int launch_fullgrid_m(bj_ctx* c, long n_walkers, cudaStream_t st, int c0, int c1, Sel sel) {
    if (c->dims.nspcal > 0) return launch_fullgrid_s<NIFO, MODEL, true>(c, n_walkers, st, c0, c1, sel);
    return launch_fullgrid_s<NIFO, MODEL, false>(c, n_walkers, st, c0, c1, sel);
}

template <int NIFO>
static int launch_fullgrid_n(bj_ctx* c, long n_walkers, cudaStream_t st, int c0, int c1, Sel sel) {
    if (c->cfg.approx == BJ_APPROX_TAYLORF2_35PN) return launch_fullgrid_m<NIFO, MODEL_35>(c, n_walkers, st, c0, c1, sel);
    return launch_fullgrid_m<NIFO, MODEL_GENERIC>(c, n_walkers, st, c0, c1, sel);
}

static int launch_fullgrid(bj_ctx* c, long n_walkers, cudaStream_t st, int c0, int c1, Sel sel = Sel{nullptr, nullptr}) {
    switch (c->cfg.n_ifo) {
        case 1: return launch_fullgrid_n<1>(c, n_walkers, st, c0, c1, sel);
        case 2: return launch_fullgrid_n<2>(c, n_walkers, st, c0, c1, sel);
        default: return launch_fullgrid_n<3>(c, n_walkers, st, c0, c1, sel);
    }
}

static EpilogueTables epilogue_tables(const bj_ctx* c) {
    EpilogueTables tab;
    tab.chunks = c->d_chunks;
    tab.n_chunks = c->n_chunks;
    tab.n_cells = c->n_cells;
    tab.cell_band = c->d_cell_band;
    tab.cell_seg = c->d_cell_seg;
    tab.gram = c->d_gram;
    tab.dd_band = c->d_dd_band;
    return tab;
}

static int launch_epilogue(bj_ctx* c, long n_walkers, cudaStream_t st, double* logl, double* parts, int mode, int c0,
                           int c1, double* sums_io, Sel sel = Sel{nullptr, nullptr}) {
    epilogue_kernel<<<(unsigned)((n_walkers + kEpiWalkers - 1) / kEpiWalkers), kEpiWalkers * kEpiSlices, 0, st>>>(
        c->cfg, epilogue_tables(c), c->dims, c->d_partial, c->d_coef, c->d_coef_ex, c->cap_walkers, n_walkers, logl,
        parts, mode, c0, c1, sums_io, sel, MultiTable{});
    CUDA_TRY(cudaGetLastError());
    return BJ_OK;
}

// ---- contexts with compressed axes: ONE launch per kernel type over all tables (kernels.cuh: MultiTable) -----------------
static MultiTable multi_table(const bj_ctx* c, bool for_epilogue) {
    MultiTable mt;
    memset(&mt, 0, sizeof(mt));
    mt.n = c->n_levels + 1;
    mt.list_stride = c->cap_walkers;
    mt.lists = c->d_lists;
    mt.counts = c->d_counts;
    if (c->focus) {
        // the focus table is list n_levels + 1: evaluated by its own FP64 launch (no window / tile chunks), finished by the
        // common epilogue (FP64 records are unscaled: <d|h> correction 1)
        mt.n = c->n_levels + 2;
        TableRef& F = mt.t[c->n_levels + 1];
        F.chunks = c->focus->d_chunks;
        F.n_chunks = for_epilogue ? c->focus->n_chunks : 0;
        F.n_precise = 0;
        F.dh_fix_re = 1.0;
        F.dh_fix_im = 0.0;
    }
    for (int lv = 0; lv <= c->n_levels; ++lv) {
        const bj_ctx* m = lv < c->n_levels ? c->cmp[lv] : c;
        TableRef& T = mt.t[lv];
        T.basis = (const unsigned char*)m->d_tile_basis;
        T.data = (const unsigned char*)(m->d_tile_data_plain ? m->d_tile_data_plain : m->d_tile_data);
        T.rec_ul = (const unsigned char*)m->d_rec;
        T.rec64 = m->d_rec64 ? m->d_rec64 - (size_t)m->prec_lo * (4 + 2 * c->cfg.n_ifo) : nullptr;
        T.chunks = m->d_chunks;
        T.n_chunks = m->n_chunks;
        T.n_precise = m->n_precise;
        T.dh_fix_re = m->cfg.dh_fix_re;
        T.dh_fix_im = m->cfg.dh_fix_im;
    }
    return mt;
}

template <int NIFO, int MODEL, int SC>
static int launch_multi(bj_ctx* c, long n_walkers, cudaStream_t st, double* logl, double* parts) {
    const MultiTable mt = multi_table(c, false);
    constexpr int RB = RecM<NIFO>::kBytes;
    constexpr int RS = Rec<NIFO>::kDoubles;
    int max_precise = 0, max_fast = 0;
    for (int lv = 0; lv < mt.n; ++lv) {
        max_precise = std::max(max_precise, mt.t[lv].n_precise);
        max_fast = std::max(max_fast, mt.t[lv].n_chunks - mt.t[lv].n_precise);
    }
    const Sel none = {nullptr, nullptr};
    if (max_precise > 0) {
        const size_t smem64 = (size_t)kStages * kTileBins * RS * 8 + kStages * sizeof(uint64_t);
        auto pk = fullgrid_window_kernel<NIFO, MODEL, false>;
        CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
        dim3 pgrid((unsigned)max_precise, (unsigned)((n_walkers + kThreads - 1) / kThreads + mt.n - 1));
        pk<<<pgrid, kThreads, smem64, st>>>(nullptr, nullptr, c->d_coef, c->cap_walkers, n_walkers, c->d_partial, nullptr,
                                            c->dims, nullptr, 0, c->grid_df, none, mt);
        CUDA_TRY(cudaGetLastError());
    }
    {
        // the non-uniform-axis instantiation serves every table, the full grid included (FP64 ramp per node and detector)
        constexpr int RT = TileRec<MODEL>::kBasisBytes + TileRec<MODEL>::kDataBytes;
        const size_t smem_t = (size_t)kStagesT * TileRec<MODEL>::kTileBins * RT + kStagesT * sizeof(uint64_t);
        auto tkern = fullgrid_tile_kernel<NIFO, MODEL, SC, false, false, false>;
        CUDA_TRY(cudaFuncSetAttribute(tkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
        dim3 tgrid((unsigned)max_fast, (unsigned)((n_walkers + kTileWalkers - 1) / kTileWalkers + mt.n - 1));
        CUDA_TRY(cudaEventRecord(c->ev_tile[0], st));
        tkern<<<tgrid, kTileThreads, smem_t, st>>>(nullptr, nullptr, nullptr, RB, nullptr, c->d_coef, c->cap_walkers, n_walkers,
                                                   c->d_partial, 0, 0.0, nullptr, c->dims, none, mt);
        CUDA_TRY(cudaEventRecord(c->ev_tile[1], st));
        c->tile_timed = true;
        CUDA_TRY(cudaGetLastError());
        note_launch(c, tgrid, smem_t, RT / 8, MODEL, kTileThreads);
    }
    {
        // wide-phase fix-up on the full grid (classify_kernel sends such walkers there): early exit unless one is flagged
        const size_t smem = (size_t)kStages * kTileBins * RB + kStages * sizeof(uint64_t);
        auto fixup = fullgrid_mixed_kernel<NIFO, MODEL, SC, true, false>;
        CUDA_TRY(cudaFuncSetAttribute(fixup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)(c->n_chunks - c->n_precise), (unsigned)((n_walkers + kThreads - 1) / kThreads));
        fixup<<<grid, kThreads, smem, st>>>((const unsigned char*)c->d_rec, c->d_chunks, c->d_coef, c->cap_walkers, n_walkers,
                                            c->d_partial, nullptr, c->dims, c->n_precise);
        CUDA_TRY(cudaGetLastError());
    }
    if (c->focus) {
        // walkers next to the fiducial point: a few hundred nodes, all FP64 (fullgrid_kernel with the fiducial subtracted)
        bj_ctx* f = c->focus;
        const size_t smem64 = (size_t)kStages * kTileBins * RS * 8 + kStages * sizeof(uint64_t);
        auto fk = fullgrid_kernel<NIFO, MODEL, BJ_SINCOS_FP64, false>;
        CUDA_TRY(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
        const int lf = c->n_levels + 1;
        const Sel sf = {c->d_lists + (size_t)lf * c->cap_walkers, c->d_counts + lf};
        dim3 fgrid((unsigned)f->n_chunks, (unsigned)((n_walkers + kThreads - 1) / kThreads));
        fk<<<fgrid, kThreads, smem64, st>>>((const double*)f->d_rec, f->d_chunks, c->d_coef, c->cap_walkers, n_walkers,
                                            c->d_partial, nullptr, c->dims, nullptr, 0, sf, c->d_fid);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(c->ev[2], st));
    const MultiTable me = multi_table(c, true);
    epilogue_kernel<<<(unsigned)((n_walkers + kEpiWalkers - 1) / kEpiWalkers + me.n - 1), kEpiWalkers * kEpiSlices, 0, st>>>(
        c->cfg, epilogue_tables(c), c->dims, c->d_partial, c->d_coef, nullptr, c->cap_walkers, n_walkers, logl, parts, 0, 0, 0,
        nullptr, none, me);
    CUDA_TRY(cudaGetLastError());
    return BJ_OK;
}
template <int NIFO, int MODEL>
static int launch_multi_s(bj_ctx* c, long n_walkers, cudaStream_t st, double* logl, double* parts) {
    if (c->sincos == BJ_SINCOS_MUFU) return launch_multi<NIFO, MODEL, BJ_SINCOS_MUFU>(c, n_walkers, st, logl, parts);
    return launch_multi<NIFO, MODEL, BJ_SINCOS_POLY32>(c, n_walkers, st, logl, parts);
}
template <int NIFO>
static int launch_multi_m(bj_ctx* c, long n_walkers, cudaStream_t st, double* logl, double* parts) {
    if (c->cfg.approx == BJ_APPROX_TAYLORF2_35PN) return launch_multi_s<NIFO, MODEL_35>(c, n_walkers, st, logl, parts);
    return launch_multi_s<NIFO, MODEL_GENERIC>(c, n_walkers, st, logl, parts);
}
static int launch_multi_n(bj_ctx* c, long n_walkers, cudaStream_t st, double* logl, double* parts) {
    switch (c->cfg.n_ifo) {
        case 1: return launch_multi_m<1>(c, n_walkers, st, logl, parts);
        case 2: return launch_multi_m<2>(c, n_walkers, st, logl, parts);
        default: return launch_multi_m<3>(c, n_walkers, st, logl, parts);
    }
}

template <int NIFO, int MODEL, int SC>
static void launch_tm_integrand(bj_ctx* c, long w0, long n_sub, cudaStream_t st) {
    dim3 grid((unsigned)c->n_chunks, (unsigned)((n_sub + kThreads - 1) / kThreads));
    timemarg_integrand_kernel<NIFO, MODEL, SC><<<grid, kThreads, 0, st>>>(
        c->cfg, (const unsigned char*)c->d_rec, c->d_chunks, c->n_bins, c->d_coef, c->cap_walkers, w0, n_sub,
        c->tm_n_full, c->tm_offset, c->d_tm);
}
template <int NIFO, int MODEL>
static void launch_tm_integrand_sc(bj_ctx* c, long w0, long n_sub, cudaStream_t st) {
    if (c->sincos == BJ_SINCOS_MUFU) launch_tm_integrand<NIFO, MODEL, BJ_SINCOS_MUFU>(c, w0, n_sub, st);
    else                             launch_tm_integrand<NIFO, MODEL, BJ_SINCOS_POLY32>(c, w0, n_sub, st);
}
template <int NIFO>
static void launch_tm_integrand_m(bj_ctx* c, long w0, long n_sub, cudaStream_t st) {
    if (c->cfg.approx == BJ_APPROX_TAYLORF2_35PN) launch_tm_integrand_sc<NIFO, MODEL_35>(c, w0, n_sub, st);
    else                                          launch_tm_integrand_sc<NIFO, MODEL_GENERIC>(c, w0, n_sub, st);
}

// time-marginalised logL of walkers [0, n_walkers): sub-batches of integrand -> batched FFT -> logsumexp
static int run_timemarg(bj_ctx* c, long n_walkers, cudaStream_t st, double* out_logl_dev) {
    if (!c->d_tm) {
        // sub-batch: at most ~2 GiB of complex128 spectra
        long b = (long)((2LL << 30) / (c->tm_n_full * (long)sizeof(double2)));
        if (b < 1) b = 1;
        if (b > 1024) b = 1024;
        c->tm_batch = b;
        CUDA_TRY(cudaMalloc(&c->d_tm, (size_t)b * c->tm_n_full * sizeof(double2)));
        int n = (int)c->tm_n_full;
        if (g_cufft.PlanMany(&c->tm_plan, 1, &n, nullptr, 1, n, nullptr, 1, n, kCufftZ2Z, (int)b) != 0)
            return fail(BJ_ERR_CUDA, "cufftPlanMany(n=%d, batch=%ld) failed", n, b);
        c->tm_plan_ok = true;
    }
    if (n_walkers > c->cap_tm_R) {
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(c->d_tm_R);
        c->d_tm_R = nullptr;
        CUDA_TRY(cudaMalloc(&c->d_tm_R, (size_t)n_walkers * sizeof(double)));
        c->cap_tm_R = n_walkers;
    }
    for (long w0 = 0; w0 < n_walkers; w0 += c->tm_batch) {
        const long n_sub = std::min(c->tm_batch, n_walkers - w0);
        // a batched plan transforms a fixed number of rows: the full sub-batch, or -- for calls with few walkers -- a
        // second, small plan (next power of two; rebuilt when it does not fit), so that one point does not pay for 1024
        int plan = c->tm_plan;
        long rows = c->tm_batch;
        if (n_sub * 4 <= c->tm_batch) {
            long want = 1;
            while (want < n_sub) want *= 2;
            if (c->tm_small != want) {
                if (c->tm_small > 0) g_cufft.Destroy(c->tm_plan_small);
                c->tm_small = 0;
                int n = (int)c->tm_n_full;
                if (g_cufft.PlanMany(&c->tm_plan_small, 1, &n, nullptr, 1, n, nullptr, 1, n, kCufftZ2Z, (int)want) != 0)
                    return fail(BJ_ERR_CUDA, "cufftPlanMany(n=%d, batch=%ld) failed", n, want);
                c->tm_small = want;
            }
            plan = c->tm_plan_small;
            rows = want;
        }
        if (g_cufft.SetStream(plan, st) != 0) return fail(BJ_ERR_CUDA, "cufftSetStream failed");
        // rows beyond n_sub are zero and ignored
        CUDA_TRY(cudaMemsetAsync(c->d_tm, 0, (size_t)rows * c->tm_n_full * sizeof(double2), st));
        switch (c->cfg.n_ifo) {
            case 1: launch_tm_integrand_m<1>(c, w0, n_sub, st); break;
            case 2: launch_tm_integrand_m<2>(c, w0, n_sub, st); break;
            default: launch_tm_integrand_m<3>(c, w0, n_sub, st); break;
        }
        CUDA_TRY(cudaGetLastError());
        if (g_cufft.ExecZ2Z(plan, c->d_tm, c->d_tm, kCufftForward) != 0) return fail(BJ_ERR_CUDA, "cufftExecZ2Z failed");
        timemarg_reduce_kernel<<<(unsigned)n_sub, 256, 0, st>>>(c->d_tm, c->tm_n_full, c->cfg.marg_phi_ref, w0, c->d_tm_R);
        CUDA_TRY(cudaGetLastError());
    }
    timemarg_finish_kernel<<<(unsigned)((n_walkers + 127) / 128), 128, 0, st>>>(c->cfg, c->d_gram, c->d_coef, c->cap_walkers,
                                                                                n_walkers, c->d_tm_R, out_logl_dev);
    CUDA_TRY(cudaGetLastError());
    c->launch_info[0] = c->n_chunks; c->launch_info[1] = (int)c->tm_batch; c->launch_info[2] = kThreads; c->launch_info[3] = 0;
    c->launch_info[4] = c->bins_per_chunk; c->launch_info[5] = c->n_chunks; c->launch_info[6] = (int)(c->tm_n_full & 0x7fffffff);
    c->launch_info[7] = 300;
    return BJ_OK;
}

static int check_extra_map(const bj_ctx* c, const bj_extra_map* em, int ndim) {
    const bool need = c->dims.nspcal > 0 || c->dims.nweights > 0;
    if (need && !em)
        return fail(BJ_ERR_INVALID, "this context was created with calibration envelopes / PSD weights: use the _ex entry");
    if (!need) return BJ_OK;
    if (c->kind != KIND_FULLGRID) return fail(BJ_ERR_UNSUPPORTED, "extras are only defined for the full-grid likelihood");
    auto bad = [&](int col) { return col >= ndim || col < BJ_COL_CONST; };
    for (int i = 0; i < c->cfg.n_ifo; ++i) {
        for (int j = 0; j < c->dims.nspcal; ++j)
            if (bad(em->spcal_amp_col[i][j]) || bad(em->spcal_phi_col[i][j]))
                return fail(BJ_ERR_INVALID, "calibration node %d of detector %d is not mapped", j, i);
        for (int b = 0; b < c->dims.nweights; ++b)
            if (bad(em->weight_col[i][b])) return fail(BJ_ERR_INVALID, "PSD weight %d of detector %d is not mapped", b, i);
    }
    return BJ_OK;
}

static int run_device(bj_ctx* c, const double* x_dev, long n_walkers, int ndim, const bj_param_map* map,
                      const bj_extra_map* emap, double* out_logl_dev, double* out_parts_dev, cudaStream_t st) {
    int rc = check_map(map, ndim);
    if (rc) return rc;
    rc = check_extra_map(c, emap, ndim);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    rc = ensure_workspace(c, n_walkers);
    if (rc) return rc;
    ParamMapDev pm;
    to_dev_map(map, pm);
    const int tpb = 128;
    const unsigned nblk = (unsigned)((n_walkers + tpb - 1) / tpb);
    const unsigned nblk_pro = (unsigned)((n_walkers + kProWalkers - 1) / kProWalkers);
    const dim3 blk_pro(kProWalkers, 1 + BJ_MAX_IFO);
    CUDA_TRY(cudaEventRecord(c->ev[0], st));
    prologue_kernel<<<nblk_pro, blk_pro, 0, st>>>(c->cfg, pm, x_dev, ndim, n_walkers, c->d_coef, c->cap_walkers,
                                                  c->n_levels > 0 ? c->d_counts : nullptr);
    if (c->d_coef_ex)
        prologue_extras_kernel<<<nblk, tpb, 0, st>>>(c->cfg, c->dims, *emap, x_dev, ndim, n_walkers, c->d_coef,
                                                     c->d_coef_ex, c->cap_walkers);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev[1], st));
    if (c->kind == KIND_FULLGRID && c->tm) {
        if (out_parts_dev) {
            // per-detector <d|h>, <h|h> (Detector.compute_inner_products) do not depend on the marginalisation: one pass
            // of the fused kernels fills them (its logL is overwritten by the marginalised one below)
            rc = launch_fullgrid(c, n_walkers, st, 0, c->n_chunks);
            if (rc) return rc;
            rc = launch_epilogue(c, n_walkers, st, out_logl_dev, out_parts_dev, 0, 0, c->n_chunks, nullptr);
            if (rc) return rc;
        }
        rc = run_timemarg(c, n_walkers, st, out_logl_dev);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(c->ev[2], st));
    } else if (c->kind == KIND_FULLGRID && c->n_levels > 0) {
        // compressed frequency axes: classify_kernel sorts the walkers into one work list per table (levels of growing
        // design budget, then the full grid); ONE launch of each bin kernel and of the epilogue then serves all lists, every
        // CTA on the table of the list its walker tile belongs to (launch_multi).  The lists are disjoint, so the tables
        // share the partial-sum buffer.
        c->last_walkers = n_walkers;
        classify_kernel<<<(unsigned)((n_walkers + 3) / 4), 128, 0, st>>>(c->steep, c->cfg.n_ifo, c->d_coef, c->cap_walkers,
                                                                        n_walkers, c->d_lists, c->d_counts);
        CUDA_TRY(cudaGetLastError());
        rc = launch_multi_n(c, n_walkers, st, out_logl_dev, out_parts_dev);
        if (rc) return rc;
    } else if (c->kind == KIND_FULLGRID) {
        rc = launch_fullgrid(c, n_walkers, st, 0, c->n_chunks);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(c->ev[2], st));
        rc = launch_epilogue(c, n_walkers, st, out_logl_dev, out_parts_dev, 0, 0, c->n_chunks, nullptr);
        if (rc) return rc;
    } else if (c->kind == KIND_ROQ) {
        rc = launch_roq(c->cfg, c->roq, c->d_coef, c->cap_walkers, n_walkers, out_logl_dev, out_parts_dev, st,
                        c->launch_info);
        if (rc) return fail(BJ_ERR_CUDA, "ROQ kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        CUDA_TRY(cudaEventRecord(c->ev[2], st));
    } else {
        rc = launch_relbin(c->cfg, c->relbin, c->d_coef, c->cap_walkers, n_walkers, out_logl_dev, out_parts_dev, st,
                           c->launch_info);
        if (rc) return fail(BJ_ERR_CUDA, "relative-binning kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        CUDA_TRY(cudaEventRecord(c->ev[2], st));
    }
    CUDA_TRY(cudaEventRecord(c->ev[3], st));
    c->timed = true;
    return BJ_OK;
}

extern "C" int bj_loglike_device_ex(bj_ctx* c, const double* x_dev, int64_t n_walkers, int32_t ndim,
                                    const bj_param_map* map, const bj_extra_map* emap, double* out_logl_dev,
                                    double* out_parts_dev, void* stream) {
    if (!c) return fail(BJ_ERR_INVALID, "null context");
    if (n_walkers == 0) return BJ_OK;
    if (n_walkers < 0 || ndim <= 0 || !x_dev || !out_logl_dev) return fail(BJ_ERR_INVALID, "bad batch arguments");
    // walkers are independent: batches beyond 2^21 are evaluated in slices (the walker-tile index of the bin kernels is a
    // grid.y coordinate, limited to 65 535 tiles of 64)
    const int64_t slice = 1LL << 21;
    for (int64_t w0 = 0; w0 < n_walkers; w0 += slice) {
        const long nw = (long)std::min(slice, n_walkers - w0);
        int rc = run_device(c, x_dev + (size_t)w0 * ndim, nw, ndim, map, emap, out_logl_dev + w0,
                            out_parts_dev ? out_parts_dev + (size_t)w0 * c->cfg.n_ifo * 3 : nullptr, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return BJ_OK;
}

// ---- frequency-axis split: local partial sums of a chunk range, and the epilogue on reduced sums -----------
extern "C" int bj_num_chunks(bj_ctx* c, int32_t* n_chunks) {
    if (!c || !n_chunks) return fail(BJ_ERR_INVALID, "null argument");
    if (c->kind != KIND_FULLGRID) return fail(BJ_ERR_UNSUPPORTED, "only the full-grid likelihood has frequency chunks");
    *n_chunks = c->n_chunks;
    return BJ_OK;
}

extern "C" int bj_partial_device(bj_ctx* c, const double* x_dev, int64_t n_walkers, int32_t ndim,
                                 const bj_param_map* map, const bj_extra_map* emap, int32_t chunk_begin,
                                 int32_t chunk_end, double* out_sums_dev, void* stream) {
    if (!c) return fail(BJ_ERR_INVALID, "null context");
    if (c->kind != KIND_FULLGRID) return fail(BJ_ERR_UNSUPPORTED, "only the full-grid likelihood can be split in frequency");
    if (c->tm) return fail(BJ_ERR_UNSUPPORTED, "the time-marginalised likelihood needs the whole spectrum of a walker on one GPU: it cannot be split in frequency");
    if (n_walkers <= 0 || ndim <= 0 || !x_dev || !out_sums_dev) return fail(BJ_ERR_INVALID, "bad batch arguments");
    if (chunk_begin < 0 || chunk_end > c->n_chunks || chunk_begin > chunk_end)
        return fail(BJ_ERR_INVALID, "chunk range [%d, %d) outside [0, %d)", chunk_begin, chunk_end, c->n_chunks);
    int rc = check_map(map, ndim);
    if (rc) return rc;
    rc = check_extra_map(c, emap, ndim);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(c->device));
    rc = ensure_workspace(c, (long)n_walkers);
    if (rc) return rc;
    ParamMapDev pm;
    to_dev_map(map, pm);
    const unsigned nblk = (unsigned)((n_walkers + 127) / 128);
    prologue_kernel<<<(unsigned)((n_walkers + kProWalkers - 1) / kProWalkers), dim3(kProWalkers, 1 + BJ_MAX_IFO), 0, st>>>(
        c->cfg, pm, x_dev, ndim, n_walkers, c->d_coef, c->cap_walkers, nullptr);
    if (c->d_coef_ex)
        prologue_extras_kernel<<<nblk, 128, 0, st>>>(c->cfg, c->dims, *emap, x_dev, ndim, n_walkers, c->d_coef,
                                                     c->d_coef_ex, c->cap_walkers);
    CUDA_TRY(cudaGetLastError());
    if (chunk_end > chunk_begin) {
        rc = launch_fullgrid(c, (long)n_walkers, st, chunk_begin, chunk_end);
        if (rc) return rc;
    }
    return launch_epilogue(c, (long)n_walkers, st, nullptr, nullptr, 1, chunk_begin, chunk_end, out_sums_dev);
}

extern "C" int bj_finish_device(bj_ctx* c, const double* sums_dev, int64_t n_walkers, double* out_logl_dev,
                                double* out_parts_dev, void* stream) {
    if (!c) return fail(BJ_ERR_INVALID, "null context");
    if (c->kind != KIND_FULLGRID) return fail(BJ_ERR_UNSUPPORTED, "only the full-grid likelihood can be split in frequency");
    if (c->tm) return fail(BJ_ERR_UNSUPPORTED, "the time-marginalised likelihood cannot be split in frequency");
    if (n_walkers <= 0 || !sums_dev || !out_logl_dev) return fail(BJ_ERR_INVALID, "bad arguments");
    if (n_walkers > c->cap_walkers) return fail(BJ_ERR_INVALID, "bj_finish_device must follow bj_partial_device on the same batch");
    CUDA_TRY(cudaSetDevice(c->device));
    return launch_epilogue(c, (long)n_walkers, (cudaStream_t)stream, out_logl_dev, out_parts_dev, 2, 0, 0,
                           const_cast<double*>(sums_dev));
}

extern "C" int bj_loglike_device(bj_ctx* c, const double* x_dev, int64_t n_walkers, int32_t ndim,
                                 const bj_param_map* map, double* out_logl_dev, double* out_parts_dev, void* stream) {
    return bj_loglike_device_ex(c, x_dev, n_walkers, ndim, map, nullptr, out_logl_dev, out_parts_dev, stream);
}

extern "C" int bj_loglike_ex(bj_ctx* c, const double* x_host, int64_t n_walkers, int32_t ndim, const bj_param_map* map,
                             const bj_extra_map* emap, double* out_logl_host) {
    if (!c) return fail(BJ_ERR_INVALID, "null context");
    if (n_walkers == 0) return BJ_OK;
    if (n_walkers < 0 || ndim <= 0 || !x_host || !out_logl_host) return fail(BJ_ERR_INVALID, "bad batch arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    const long need = (long)n_walkers * ndim;
    if (need > c->cap_x) {
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(c->d_x);
        c->d_x = nullptr;
        c->cap_x = 0;
        CUDA_TRY(cudaMalloc(&c->d_x, (size_t)need * sizeof(double)));
        c->cap_x = need;
    }
    if ((long)n_walkers > c->cap_logl) {
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(c->d_logl);
        c->d_logl = nullptr;
        c->cap_logl = 0;
        const long cap = pad_walkers(std::max((long)n_walkers, 2 * c->cap_logl));
        CUDA_TRY(cudaMalloc(&c->d_logl, (size_t)cap * sizeof(double)));
        c->cap_logl = cap;
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_x, x_host, (size_t)need * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int rc = bj_loglike_device_ex(c, c->d_x, n_walkers, ndim, map, emap, c->d_logl, nullptr, c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_logl_host, c->d_logl, (size_t)n_walkers * sizeof(double), cudaMemcpyDeviceToHost,
                             c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return BJ_OK;
}

extern "C" int bj_loglike(bj_ctx* c, const double* x_host, int64_t n_walkers, int32_t ndim, const bj_param_map* map,
                          double* out_logl_host) {
    return bj_loglike_ex(c, x_host, n_walkers, ndim, map, nullptr, out_logl_host);
}

extern "C" int bj_last_kernel_ms(bj_ctx* c, float* ms3) {
    if (!c || !ms3) return fail(BJ_ERR_INVALID, "null argument");
    if (!c->timed) return fail(BJ_ERR_INVALID, "no call has been timed yet");
    CUDA_TRY(cudaEventSynchronize(c->ev[3]));
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventElapsedTime(&ms3[i], c->ev[i], c->ev[i + 1]));
    return BJ_OK;
}

extern "C" int bj_last_tile_ms(bj_ctx* c, float* ms) {
    if (!c || !ms) return fail(BJ_ERR_INVALID, "null argument");
    if (!c->tile_timed) return fail(BJ_ERR_INVALID, "no call has launched the tile kernel yet");
    CUDA_TRY(cudaEventSynchronize(c->ev_tile[1]));
    CUDA_TRY(cudaEventElapsedTime(ms, c->ev_tile[0], c->ev_tile[1]));
    return BJ_OK;
}

extern "C" int bj_table_info(bj_ctx* c, int32_t table, int64_t* out4) {
    if (!c || !out4) return fail(BJ_ERR_INVALID, "null argument");
    const bj_ctx* m = nullptr;
    if (table >= 0 && table < c->n_levels) m = c->cmp[table];
    else if (table == kMaxLevels) m = c;
    else if (table == kMaxLevels + 1) m = c->focus;
    for (int q = 0; q < 4; ++q) out4[q] = 0;
    if (!m || c->kind != KIND_FULLGRID) return BJ_OK;
    out4[0] = m->n_bins_kernel;
    out4[1] = m->prec_hi - m->prec_lo;
    out4[2] = m->n_chunks;
    out4[3] = m->n_precise;
    return BJ_OK;
}

extern "C" int bj_last_launch_info(bj_ctx* c, int32_t* info8) {
    if (!c || !info8) return fail(BJ_ERR_INVALID, "null argument");
    for (int i = 0; i < 8; ++i) info8[i] = c->launch_info[i];
    return BJ_OK;
}

extern "C" int bj_last_levels(bj_ctx* c, int64_t n_walkers, int32_t* levels) {
    if (!c || !levels || n_walkers < 0) return fail(BJ_ERR_INVALID, "bad arguments");
    if (c->n_levels == 0 || n_walkers > c->last_walkers) {
        for (int64_t w = 0; w < n_walkers; ++w) levels[w] = kMaxLevels;
        return BJ_OK;
    }
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<double> lv((size_t)n_walkers);
    CUDA_TRY(cudaMemcpy(lv.data(), c->d_coef + (size_t)C_STEEP * c->cap_walkers, lv.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t w = 0; w < n_walkers; ++w)
        levels[w] = (int)lv[w] > c->n_levels ? kMaxLevels + 1 : (int)lv[w] == c->n_levels ? kMaxLevels : (int)lv[w];
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// Focus table: the compressed axis taken to its limit around ONE point (the region a converged sampler lives in).
// With the fiducial point's phase P_fid(f) and delays tau_fid,i folded into the data weights,
//     w'_ik = w_ik exp(-2 pi i [P_fid(f_k) + f_k tau_fid,i]),
// a walker's smooth factor is g'_i = Q exp(-2 pi i [(P - P_fid)(f) + f (tau_i - tau_fid,i)]): the kernels subtract the
// fiducial's coefficients (the phase is linear in them, so this costs nothing per node), and the design bound is on the
// slope RELATIVE to the fiducial,  |d(P - P_fid)/df| + max_i |tau_i - tau_fid,i| <= a (f / f_lo)^(-5/3) + b + c (f / f_hi)^(2/3)
// -- seconds, not the minutes of a chirp (a: masses and spins move the low-frequency end, b: sky position and time shift,
// c: tidal deformabilities act at the top of the band): the whole band collapses into a few runs of 64 nodes (C2: 521 729
// bins -> ~700), which are evaluated in FP64 throughout.  classify_kernel sends a walker there only if it respects that bound at every check
// frequency; everybody else is routed as before.  Relative binning (binning.py) is the first-order version of this with no
// bound and no fallback; this is exact to the truncation of the Chebyshev interpolant (~1e-12).
// ------------------------------------------------------------------------------------------------
static void drop_focus(bj_ctx* c) {
    if (c->focus) {
        cudaFree(c->focus->d_rec);
        cudaFree(c->focus->d_chunks);
        delete c->focus;
        c->focus = nullptr;
    }
    c->steep.focus = 0;
    for (int q = 0; q < 4; ++q) c->focus_info[q] = 0;
}

extern "C" int bj_unfocus(bj_ctx* c) {
    if (!c) return fail(BJ_ERR_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    drop_focus(c);
    return BJ_OK;
}

extern "C" int bj_focus(bj_ctx* c, const double* x_host, int32_t ndim, const bj_param_map* map, double slope_a,
                        double slope_b, double slope_c, int32_t* info4) {
    if (!c || !x_host) return fail(BJ_ERR_INVALID, "null argument");
    if (c->kind != KIND_FULLGRID || c->n_levels == 0)
        return fail(BJ_ERR_UNSUPPORTED, "a focus table needs a full-grid context with compressed axes");
    int rc = check_map(map, ndim);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    drop_focus(c);
    if (!(slope_a > 0.0)) slope_a = 0.25;
    if (!(slope_b > 0.0)) slope_b = 2e-3;
    if (!(slope_c >= 0.0)) slope_c = 0.03;
    // the fiducial point's coefficient vector, from the prologue itself
    ParamMapDev pm;
    to_dev_map(map, pm);
    double *d_x = nullptr, *d_coef = nullptr;
    std::vector<double> coef(C_NUM, 0.0);
    cudaError_t e = cudaMalloc(&d_x, ndim * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_coef, (size_t)C_NUM * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_x, x_host, ndim * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        prologue_kernel<<<1, dim3(kProWalkers, 1 + BJ_MAX_IFO)>>>(c->cfg, pm, d_x, ndim, 1, d_coef, 1, nullptr);
        e = cudaMemcpy(coef.data(), d_coef, C_NUM * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_x);
    cudaFree(d_coef);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "bj_focus: %s", cudaGetErrorString(e));
    if (coef[C_VALID] != 1.0) return fail(BJ_ERR_INVALID, "the fiducial point is unphysical or needs the wide-phase path");
    const int nifo = c->cfg.n_ifo;
    const long N = (long)c->h_freqs.size();
    double fid[24 + BJ_MAX_IFO];
    for (int j = 0; j < 16; ++j) fid[j] = coef[C_A0 + j];
    for (int j = 0; j < 7; ++j) fid[16 + j] = coef[C_B0 + j];
    fid[23] = coef[C_C8];
    for (int i = 0; i < BJ_MAX_IFO; ++i) fid[24 + i] = i < nifo ? coef[C_TAU0 + i] : 0.0;
    // heterodyned weights (long double: P_fid ~ 1e4..1e5 turns, wanted to 1e-12 of a turn)
    std::vector<double> hw[BJ_MAX_IFO];
    for (int i = 0; i < nifo; ++i) hw[i].assign((size_t)2 * N, 0.0);
    for (long k = 0; k < N; ++k) {
        const long double f = c->h_freqs[k];
        const long double u = cbrtl(f), L = logl(f), w5 = 1.0L / (u * u * u * u * u);
        long double up = 1.0L, sa = 0.0L, sb = 0.0L;
        for (int j = 0; j < 16; ++j) {
            sa += (long double)fid[j] * up;
            if (j < 7) sb += (long double)fid[16 + j] * up;
            up *= u;
        }
        const long double P = w5 * sa + L * sb + L * L * (long double)fid[23] * u * u * u;
        for (int i = 0; i < nifo; ++i) {
            long double t = P + f * (long double)fid[24 + i];
            t -= rintl(t);
            const long double cs = cosl(-2.0L * M_PIl * t), sn = sinl(-2.0L * M_PIl * t);
            const long double wr = c->h_wts[i][2 * k], wi = c->h_wts[i][2 * k + 1];
            hw[i][2 * k] = (double)(wr * cs - wi * sn);
            hw[i][2 * k + 1] = (double)(wr * sn + wi * cs);
        }
    }
    Axis ax;
    ax.n = N;
    ax.f = c->h_freqs.data();
    for (int i = 0; i < nifo; ++i) ax.w[i] = hw[i].data();
    // runs from the focus budget: build_compressed_axis with budget(f) = a (f / f_lo)^(-5/3) + b
    CompressSpec sp = compress_spec(nullptr);
    sp.focus = 1;
    sp.focus_a = slope_a;
    sp.focus_b = slope_b;
    sp.focus_c = slope_c;
    sp.focus_f0 = c->h_freqs[0];
    sp.focus_f1 = c->h_freqs[N - 1];
    std::vector<double> bw((size_t)N, 1.0), nf, nbw, nwt[BJ_MAX_IFO];
    int info[4] = {0, 0, 0, 0};
    build_compressed_axis(ax, nifo, bw, sp, nf, nwt, nbw, info);
    Axis cx;
    cx.n = (long)nf.size();
    cx.f = nf.data();
    for (int i = 0; i < nifo; ++i) cx.w[i] = nwt[i].data();
    bj_ctx* m = new bj_ctx();
    m->kind = KIND_FULLGRID;
    m->shadow = true;
    m->device = c->device;
    m->cfg = c->cfg;
    memset(&m->dims, 0, sizeof(m->dims));
    m->n_bins = cx.n;
    Layout L;
    // all FP64: no window; chunks of 16 records -- a few hundred nodes make few CTAs as it is, and a thread walks its
    // chunk serially through ~250 dependent FP64 instructions per node (64-record chunks: 0.5 waves of CTAs, 50 us)
    if (build_layout(L, cx.n, cx.f, nullptr, nullptr, 0.0, 0.0)) { delete m; return fail(BJ_ERR_INVALID, "focus layout"); }
    {
        int per = 16;
        if (const char* ev = getenv("BAJES_B200_FOCUS_CHUNK")) per = std::max(2, std::min(kTileBins, atoi(ev) & ~1));
        std::vector<ChunkDesc> small;
        for (const ChunkDesc& cd : L.chunks)
            for (int a0 = 0; a0 < cd.len; a0 += per) {
                ChunkDesc s = cd;
                s.start = cd.start + a0;
                s.len = std::min(per, cd.len - a0);
                small.push_back(s);
            }
        L.chunks = small;
    }
    adopt_layout(m, L);
    std::vector<double> rec;
    switch (nifo) {
        case 1: build_records<1>(rec, L, cx, c->cfg.seglen); break;
        case 2: build_records<2>(rec, L, cx, c->cfg.seglen); break;
        default: build_records<3>(rec, L, cx, c->cfg.seglen); break;
    }
    for (double v : rec)
        if (!std::isfinite(v)) { delete m; return fail(BJ_ERR_INVALID, "non-finite focus table entry"); }
    e = cudaMalloc(&m->d_rec, rec.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_rec, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = upload(&m->d_chunks, L.chunks);
    if (e == cudaSuccess && !c->d_fid) e = cudaMalloc(&c->d_fid, sizeof(fid));
    if (e == cudaSuccess) e = cudaMemcpy(c->d_fid, fid, sizeof(fid), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(m->d_rec);
        cudaFree(m->d_chunks);
        delete m;
        return fail(BJ_ERR_CUDA, "uploading the focus table failed: %s", cudaGetErrorString(e));
    }
    c->focus = m;
    c->steep.focus = 1;
    for (int q = 0; q < 24 + BJ_MAX_IFO; ++q) c->steep.fid[q] = fid[q];
    for (int q = 0; q < kSteepChecks; ++q) {
        const double f = c->h_freqs[0] * std::pow(c->h_freqs[N - 1] / c->h_freqs[0], (double)q / (double)(kSteepChecks - 1));
        c->steep.budget_focus[q] = design_budget(sp, f);
    }
    for (int q = 0; q < 4; ++q) c->focus_info[q] = info[q];
    if (info4) for (int q = 0; q < 4; ++q) info4[q] = info[q];
    // the partial-sum buffer may have to grow for the focus table's chunk count
    if (c->cap_walkers > 0 && m->n_chunks > 0) {
        const long keep = c->cap_walkers;
        c->cap_walkers = 0;
        rc = ensure_workspace(c, keep);
        if (rc) return rc;
    }
    return BJ_OK;
}

// host-only: the node axis build_compressed_axis makes of a weighted frequency axis (no device needed)
extern "C" int bj_compress_axis_host(int64_t n_bins, const double* freqs, int32_t n_ifo, const double* const* weights_ri,
                                     int32_t nodes_per_run, double reach_rad, double mchirp_min, double tau_max,
                                     double margin, int64_t* out_n, double* out_freqs, double* const* out_weights_ri,
                                     int32_t* info4) {
    if (n_bins <= 0 || !freqs || !weights_ri || !out_n || !out_freqs || !out_weights_ri || n_ifo < 1 || n_ifo > BJ_MAX_IFO)
        return fail(BJ_ERR_INVALID, "bad arguments");
    CompressSpec sp = compress_spec(nullptr);
    if (nodes_per_run > 0) sp.nodes = std::max(4, std::min(64, (int)nodes_per_run));
    if (reach_rad > 0.0) sp.reach = reach_rad;
    if (mchirp_min > 0.0) sp.mchirp_min = mchirp_min;
    if (tau_max > 0.0) sp.tau_max = tau_max;
    if (margin > 0.0) sp.margin = margin;
    Axis ax;
    ax.n = (long)n_bins;
    ax.f = freqs;
    for (int i = 0; i < n_ifo; ++i) ax.w[i] = weights_ri[i];
    std::vector<double> bw((size_t)n_bins, 1.0), nf, nbw, nwt[BJ_MAX_IFO];
    int info[4];
    build_compressed_axis(ax, n_ifo, bw, sp, nf, nwt, nbw, info);
    *out_n = (int64_t)nf.size();
    memcpy(out_freqs, nf.data(), nf.size() * sizeof(double));
    for (int i = 0; i < n_ifo; ++i) memcpy(out_weights_ri[i], nwt[i].data(), nwt[i].size() * sizeof(double));
    if (info4) for (int q = 0; q < 4; ++q) info4[q] = info[q];
    return BJ_OK;
}

extern "C" int bj_compress_info(bj_ctx* c, int32_t level, int32_t* info4, double* spec4) {
    if (!c || !info4 || !spec4) return fail(BJ_ERR_INVALID, "null argument");
    const bool have = level >= 0 && level < c->n_levels;
    for (int q = 0; q < 4; ++q) { info4[q] = have ? c->cmp_info[level][q] : 0; spec4[q] = have ? c->cmp_spec[level][q] : 0.0; }
    return BJ_OK;
}

extern "C" int bj_last_level_counts(bj_ctx* c, int64_t* counts3) {
    if (!c || !counts3) return fail(BJ_ERR_INVALID, "null argument");
    for (int q = 0; q <= kMaxLevels + 1; ++q) counts3[q] = 0;
    if (c->n_levels == 0 || c->last_walkers <= 0) return BJ_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    int h[kMaxLevels + 2];
    CUDA_TRY(cudaMemcpy(h, c->d_counts, sizeof(h), cudaMemcpyDeviceToHost));
    // levels that do not exist stay 0; the full grid is always reported third, the focus table fourth
    for (int q = 0; q < c->n_levels; ++q) counts3[q] = h[q];
    counts3[kMaxLevels] = h[c->n_levels];
    counts3[kMaxLevels + 1] = c->focus ? h[c->n_levels + 1] : 0;
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// per-point waveform entry points
// ------------------------------------------------------------------------------------------------
extern "C" int bj_project_waveform(bj_ctx* c, const double* x_host, int32_t ndim, const bj_param_map* map,
                                   double* out_h_host) {
    if (!c || !x_host || !out_h_host) return fail(BJ_ERR_INVALID, "null argument");
    if (c->kind != KIND_FULLGRID) return fail(BJ_ERR_INVALID, "bj_project_waveform needs a full-grid context");
    int rc = check_map(map, ndim);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    rc = ensure_workspace(c, 1);
    if (rc) return rc;
    double* d_x = nullptr;
    double2* d_h = nullptr;
    CUDA_TRY(cudaMalloc(&d_x, ndim * sizeof(double)));
    cudaError_t e = cudaMalloc(&d_h, (size_t)c->cfg.n_ifo * c->n_bins * sizeof(double2));
    if (e != cudaSuccess) { cudaFree(d_x); return fail(BJ_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    ParamMapDev pm;
    to_dev_map(map, pm);
    cudaMemcpyAsync(d_x, x_host, ndim * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    prologue_kernel<<<1, dim3(kProWalkers, 1 + BJ_MAX_IFO), 0, c->stream>>>(c->cfg, pm, d_x, ndim, 1, c->d_coef, c->cap_walkers, nullptr);
    const unsigned nblk = (unsigned)((c->n_bins + 255) / 256);
    waveform_kernel<<<nblk, 256, 0, c->stream>>>(c->cfg, c->d_freqs, c->n_bins, c->d_coef, c->cap_walkers, d_h);
    cudaMemcpyAsync(out_h_host, d_h, (size_t)c->cfg.n_ifo * c->n_bins * sizeof(double2), cudaMemcpyDeviceToHost,
                    c->stream);
    e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_x);
    cudaFree(d_h);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "waveform kernel failed: %s", cudaGetErrorString(e));
    return BJ_OK;
}

extern "C" int bj_polarizations(int device, int approx, double seglen, int centre_shift, int64_t n_freqs,
                                const double* freqs, const double* x_host, int32_t ndim, const bj_param_map* map,
                                double* out_hp_host, double* out_hc_host) {
    if (!freqs || !x_host || !out_hp_host || !out_hc_host || n_freqs <= 0) return fail(BJ_ERR_INVALID, "null / empty argument");
    if (approx < 0 || approx > BJ_APPROX_TAYLORF2_55PN_75PNTIDES2020)
        return fail(BJ_ERR_UNSUPPORTED, "approximant id %d is not a TaylorF2 variant", approx);
    // the polarisations do not depend on the sky position: fill the detector-only slots if absent
    bj_param_map m = *map;
    const int sky[] = {BJ_P_RA, BJ_P_DEC, BJ_P_PSI, BJ_P_TIME_SHIFT, BJ_P_T_GPS};
    for (int s : sky)
        if (m.column[s] == BJ_COL_ABSENT) { m.column[s] = BJ_COL_CONST; m.value[s] = 0.0; }
    int rc = check_map(&m, ndim);
    if (rc) return rc;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(BJ_ERR_NO_DEVICE, "no CUDA device visible: bajes_b200 has no CPU fallback");
    }
    CUDA_TRY(cudaSetDevice(device));
    DeviceConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.approx = approx;
    cfg.n_ifo = 0;
    cfg.seglen = seglen;
    ParamMapDev pm;
    to_dev_map(&m, pm);
    double *d_x = nullptr, *d_f = nullptr, *d_coef = nullptr;
    double2 *d_hp = nullptr, *d_hc = nullptr;
    cudaError_t e = cudaMalloc(&d_x, ndim * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_f, (size_t)n_freqs * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_coef, (size_t)C_NUM * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_hp, (size_t)n_freqs * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&d_hc, (size_t)n_freqs * sizeof(double2));
    if (e == cudaSuccess) e = cudaMemcpy(d_x, x_host, ndim * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_f, freqs, (size_t)n_freqs * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        prologue_kernel<<<1, dim3(kProWalkers, 1 + BJ_MAX_IFO)>>>(cfg, pm, d_x, ndim, 1, d_coef, 1, nullptr);
        polarizations_kernel<<<(unsigned)((n_freqs + 255) / 256), 256>>>(seglen, centre_shift, d_f, n_freqs, d_coef, 1,
                                                                         d_hp, d_hc);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_hp_host, d_hp, (size_t)n_freqs * sizeof(double2), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_hc_host, d_hc, (size_t)n_freqs * sizeof(double2), cudaMemcpyDeviceToHost);
    cudaFree(d_x); cudaFree(d_f); cudaFree(d_coef); cudaFree(d_hp); cudaFree(d_hc);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "bj_polarizations: %s", cudaGetErrorString(e));
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// ROQ basis construction (roq_basis.cuh): training bank and reduced-basis greedy, device-resident
// ------------------------------------------------------------------------------------------------
extern "C" int bj_wavebank(int device, int approx, int64_t n_freqs, const double* freqs_host, const double* x_host,
                           int64_t n_points, int32_t ndim, const bj_param_map* map, int32_t quadratic, void* out_dev) {
    if (!freqs_host || !x_host || !out_dev || n_freqs <= 0 || n_points <= 0 || ndim <= 0) return fail(BJ_ERR_INVALID, "null / empty argument");
    if (approx < 0 || approx > BJ_APPROX_TAYLORF2_55PN_75PNTIDES2020)
        return fail(BJ_ERR_UNSUPPORTED, "approximant id %d is not a TaylorF2 variant", approx);
    bj_param_map m = *map;
    const int sky[] = {BJ_P_RA, BJ_P_DEC, BJ_P_PSI, BJ_P_TIME_SHIFT, BJ_P_T_GPS};
    for (int sl : sky)
        if (m.column[sl] == BJ_COL_ABSENT) { m.column[sl] = BJ_COL_CONST; m.value[sl] = 0.0; }
    int rc = check_map(&m, ndim);
    if (rc) return rc;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(BJ_ERR_NO_DEVICE, "no CUDA device visible: bajes_b200 has no CPU fallback");
    }
    CUDA_TRY(cudaSetDevice(device));
    DeviceConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.approx = approx;
    cfg.n_ifo = 0;
    ParamMapDev pm;
    to_dev_map(&m, pm);
    const long stride = pad_walkers((long)n_points);
    double *d_x = nullptr, *d_f = nullptr, *d_coef = nullptr;
    cudaError_t e = cudaMalloc(&d_x, (size_t)n_points * ndim * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_f, (size_t)n_freqs * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_coef, (size_t)C_NUM * stride * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_x, x_host, (size_t)n_points * ndim * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_f, freqs_host, (size_t)n_freqs * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        prologue_kernel<<<(unsigned)((n_points + kProWalkers - 1) / kProWalkers), dim3(kProWalkers, 1 + BJ_MAX_IFO)>>>(
            cfg, pm, d_x, ndim, (long)n_points, d_coef, stride, nullptr);
        // grid.y is limited to 65 535 rows per launch
        for (int64_t w0 = 0; w0 < n_points && e == cudaSuccess; w0 += 32768) {
            const long nw = (long)std::min<int64_t>(32768, n_points - w0);
            dim3 grid((unsigned)((n_freqs + 255) / 256), (unsigned)nw);
            wavebank_kernel<<<grid, 256>>>(d_f, (long)n_freqs, d_coef + w0, stride, nw, quadratic,
                                           (double2*)out_dev + (size_t)w0 * n_freqs);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_x); cudaFree(d_f); cudaFree(d_coef);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "bj_wavebank: %s", cudaGetErrorString(e));
    return BJ_OK;
}

extern "C" int bj_rb_greedy(int device, int64_t n_train, int64_t n_freqs, void* train_dev, const double* weights_host,
                            double df, double tol, int32_t n_prebuilt, int32_t max_basis, void* basis_dev, int32_t* out_picks,
                            double* out_sigmas, int32_t* out_n, float* out_ms) {
    if (!train_dev || !weights_host || !basis_dev || !out_picks || !out_sigmas || !out_n || n_train <= 0 || n_freqs <= 0 ||
        max_basis <= 0 || n_prebuilt < 0 || n_prebuilt > max_basis)
        return fail(BJ_ERR_INVALID, "null / empty argument");
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) {
        cudaGetLastError();
        return fail(BJ_ERR_NO_DEVICE, "no CUDA device visible: bajes_b200 has no CPU fallback");
    }
    CUDA_TRY(cudaSetDevice(device));
    double2* R = (double2*)train_dev;
    double2* B = (double2*)basis_dev;
    double *d_w = nullptr, *d_sig2 = nullptr, *d_val = nullptr;
    int* d_pick = nullptr;
    cudaError_t e = cudaMalloc(&d_w, (size_t)n_freqs * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_sig2, (size_t)n_train * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_val, sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_pick, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(d_w, weights_host, (size_t)n_freqs * sizeof(double), cudaMemcpyHostToDevice);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (e == cudaSuccess) e = cudaEventCreate(&ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ev1);
    const double four_df = 4.0 * df;
    int n = n_prebuilt, np_out = 0;
    std::vector<char> taken((size_t)n_train, 0);
    if (e == cudaSuccess) {
        cudaEventRecord(ev0);
        // a prebuilt basis (roq.py:512-520): the training rows are first reduced to their residuals against it
        for (int j = 0; j < n_prebuilt; ++j)
            rb_project_kernel<<<(unsigned)n_train, kRbThreads>>>(R, (long)n_freqs, d_w, four_df, B + (size_t)j * n_freqs, d_sig2,
                                                                  nullptr);
        if (n_prebuilt == 0) rb_norm_kernel<<<(unsigned)n_train, kRbThreads>>>(R, (long)n_freqs, d_w, four_df, d_sig2);
        // rb_greedy (roq.py:537-612): the worst-approximated training function joins the basis until its error is below
        // tol, or it has been picked before (accuracy cannot be reached), or max_basis is full.  The reference seeds the
        // basis with training function 0 (:508-511): the first pick is forced to row 0.
        while (n < max_basis) {
            int pick = 0;
            double val = 0.0;
            if (n == 0) {        // (only without a prebuilt basis)
                const int zero = 0;
                cudaMemcpy(d_pick, &zero, sizeof(int), cudaMemcpyHostToDevice);
                cudaMemcpy(d_val, d_sig2, sizeof(double), cudaMemcpyDeviceToDevice);
            } else {
                rb_argmax_kernel<<<1, 1024>>>(d_sig2, (long)n_train, d_pick, d_val);
            }
            e = cudaMemcpy(&pick, d_pick, sizeof(int), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(&val, d_val, sizeof(double), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) break;
            const double sigma = std::sqrt(std::fmax(val, 0.0));
            if (n > 0 && (sigma <= tol || taken[pick] || !(val > 0.0))) {
                out_sigmas[np_out] = sigma;      // the error at which the loop ended
                break;
            }
            taken[pick] = 1;
            out_picks[np_out] = pick;
            out_sigmas[np_out] = sigma;
            ++np_out;
            rb_take_kernel<<<(unsigned)((n_freqs + 255) / 256), 256>>>(R, (long)n_freqs, d_pick, d_val, B + (size_t)n * n_freqs);
            rb_project_kernel<<<(unsigned)n_train, kRbThreads>>>(R, (long)n_freqs, d_w, four_df, B + (size_t)n * n_freqs, d_sig2,
                                                                  nullptr);
            ++n;
            e = cudaGetLastError();
            if (e != cudaSuccess) break;
        }
        cudaEventRecord(ev1);
        if (e == cudaSuccess) e = cudaEventSynchronize(ev1);
        if (e == cudaSuccess && out_ms) cudaEventElapsedTime(out_ms, ev0, ev1);
    }
    *out_n = n;
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    cudaFree(d_w); cudaFree(d_sig2); cudaFree(d_val); cudaFree(d_pick);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "bj_rb_greedy: %s", cudaGetErrorString(e));
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// ROQ / relative binning contexts
// ------------------------------------------------------------------------------------------------
extern "C" int bj_roq_create(bj_ctx** out, const bj_config* cfg, const bj_roq_tables* tab) {
    if (!out) return fail(BJ_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int rc = check_config(cfg);
    if (rc) return rc;
    if (!tab || tab->n_join <= 0 || tab->n_lin <= 0 || tab->n_quad <= 0 || tab->n_coef < 4 || tab->n_knots != tab->n_coef + 4)
        return fail(BJ_ERR_INVALID, "inconsistent ROQ tables (need a cubic B-spline: n_knots == n_coef + 4)");
    bj_ctx* c = new bj_ctx();
    rc = init_common(c, cfg);
    if (rc) { delete c; return rc; }
    c->kind = KIND_ROQ;
    cudaError_t e = c->roq.upload(cfg->n_ifo, tab);
    if (e != cudaSuccess) {
        bj_destroy(c);
        return fail(BJ_ERR_CUDA, "uploading ROQ tables failed: %s", cudaGetErrorString(e));
    }
    *out = c;
    return BJ_OK;
}

extern "C" int bj_relbin_create(bj_ctx** out, const bj_config* cfg, int64_t n_edges, const double* fbin,
                                const double* const* h0_ri, const double* sdat_ri) {
    if (!out) return fail(BJ_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int rc = check_config(cfg);
    if (rc) return rc;
    if (n_edges < 2 || !fbin || !h0_ri || !sdat_ri) return fail(BJ_ERR_INVALID, "bad relative-binning tables");
    bj_ctx* c = new bj_ctx();
    rc = init_common(c, cfg);
    if (rc) { delete c; return rc; }
    c->kind = KIND_RELBIN;
    cudaError_t e = c->relbin.upload(cfg->n_ifo, cfg->seglen, n_edges, fbin, h0_ri, sdat_ri);
    if (e != cudaSuccess) {
        bj_destroy(c);
        return fail(BJ_ERR_CUDA, "uploading relative-binning tables failed: %s", cudaGetErrorString(e));
    }
    *out = c;
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU result path over peer memory (include/bajes_b200.h)
// ------------------------------------------------------------------------------------------------
__global__ void peer_signal_kernel(long long* flags, int slot, long long value) {
    // everything this stream wrote before (the epilogue's logL stores into the peer buffer) becomes visible system-wide
    // before the flag does
    __threadfence_system();
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(flags + slot), "l"(value) : "memory");
}

__global__ void peer_wait_kernel(long long* flags, int first, int n, long long value, long long timeout_cycles) {
    const long long t0 = clock64();
    for (int s = first + (int)threadIdx.x; s < first + n; s += blockDim.x) {
        long long v;
        do {
            asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(flags + s) : "memory");
            if (v >= value) break;
            if (clock64() - t0 > timeout_cycles) {
                flags[BJ_PEER_FLAGS] = 1;                  // status word: timed out
                break;
            }
            __nanosleep(200);
        } while (true);
    }
    __threadfence_system();
}

extern "C" int bj_peer_create(int device, int64_t capacity_doubles, void** base_dev, unsigned char* handle64) {
    if (!base_dev || !handle64 || capacity_doubles <= 0) return fail(BJ_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    void* p = nullptr;
    const size_t bytes = (size_t)BJ_PEER_HEADER_BYTES + (size_t)capacity_doubles * sizeof(double);
    CUDA_TRY(cudaMalloc(&p, bytes));
    CUDA_TRY(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(BJ_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *base_dev = p;
    return BJ_OK;
}

extern "C" int bj_peer_open(int device, const unsigned char* handle64, void** base_dev) {
    if (!base_dev || !handle64) return fail(BJ_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(BJ_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle failed (no peer access between the two GPUs?): %s",
                    cudaGetErrorString(e));
    }
    *base_dev = p;
    return BJ_OK;
}

extern "C" int bj_peer_close(int device, void* base_dev, int opened_from_handle) {
    if (!base_dev) return BJ_OK;
    CUDA_TRY(cudaSetDevice(device));
    if (opened_from_handle) CUDA_TRY(cudaIpcCloseMemHandle(base_dev));
    else                    CUDA_TRY(cudaFree(base_dev));
    return BJ_OK;
}

extern "C" int bj_peer_signal(int device, void* base_dev, int32_t slot, int64_t value, void* stream) {
    if (!base_dev || slot < 0 || slot >= BJ_PEER_FLAGS) return fail(BJ_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)base_dev, slot, (long long)value);
    CUDA_TRY(cudaGetLastError());
    return BJ_OK;
}

extern "C" int bj_peer_scatter(int device, void* base_dev, int64_t dst_offset_bytes, const void* src_host,
                               void* stage_pinned, int32_t n_parts, const int64_t* part_begin, const int64_t* part_end,
                               const int32_t* part_flag_slot, int64_t value, void* stream) {
    if (!base_dev || !src_host || n_parts < 0 || !part_begin || !part_end || !part_flag_slot)
        return fail(BJ_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    char* dst = (char*)base_dev + dst_offset_bytes;
    const char* src = (const char*)src_host;
    char* stage = (char*)stage_pinned;
    const int64_t piece = 2 << 20;
    for (int p = 0; p < n_parts; ++p) {
        if (part_flag_slot[p] < 0 || part_flag_slot[p] >= BJ_PEER_FLAGS) return fail(BJ_ERR_INVALID, "bad flag slot");
        for (int64_t o = part_begin[p]; o < part_end[p]; o += piece) {
            const int64_t n = std::min(piece, part_end[p] - o);
            if (stage) {
                memcpy(stage + o, src + o, (size_t)n);
                CUDA_TRY(cudaMemcpyAsync(dst + o, stage + o, (size_t)n, cudaMemcpyHostToDevice, st));
            } else {
                CUDA_TRY(cudaMemcpyAsync(dst + o, src + o, (size_t)n, cudaMemcpyHostToDevice, st));
            }
        }
        peer_signal_kernel<<<1, 1, 0, st>>>((long long*)base_dev, part_flag_slot[p], (long long)value);
    }
    CUDA_TRY(cudaGetLastError());
    return BJ_OK;
}

extern "C" int bj_peer_wait(int device, void* base_dev, int32_t first_slot, int32_t n_slots, int64_t value,
                            double timeout_ms, void* stream) {
    if (!base_dev || first_slot < 0 || n_slots < 0 || first_slot + n_slots > BJ_PEER_FLAGS)
        return fail(BJ_ERR_INVALID, "bad arguments");
    if (n_slots == 0) return BJ_OK;
    CUDA_TRY(cudaSetDevice(device));
    // the clock-rate attribute is a slow driver query (~1 ms): once per device
    static int khz_cache[64] = {0};
    int khz = device < 64 ? khz_cache[device] : 0;
    if (khz == 0) {
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        if (device < 64) khz_cache[device] = khz;
    }
    const long long cycles = (long long)(timeout_ms * (double)(khz > 0 ? khz : 1000000));
    peer_wait_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((long long*)base_dev, first_slot, n_slots, (long long)value, cycles);
    CUDA_TRY(cudaGetLastError());
    return BJ_OK;
}

// ------------------------------------------------------------------------------------------------
// issue-rate micro-benchmarks (roofline denominators for the FP64 / MUFU bound kernels)
// ------------------------------------------------------------------------------------------------
__global__ void peak_dfma_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void peak_mufu_kernel(float* out, int iters, float seed) {
    float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + .1f, a2 = a0 + .2f, a3 = a0 + .3f, a4 = a0 + .4f, a5 = a0 + .5f,
          a6 = a0 + .6f, a7 = a0 + .7f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            a0 = __sinf(a0); a1 = __cosf(a1); a2 = __sinf(a2); a3 = __cosf(a3);
            a4 = __sinf(a4); a5 = __cosf(a5); a6 = __sinf(a6); a7 = __cosf(a7);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int bj_measure_peak(int device, int which, double* out) {
    if (!out) return fail(BJ_ERR_INVALID, "null out");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(BJ_ERR_NO_DEVICE, "no CUDA device visible");
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    const int threads = 256, blocks = p.multiProcessorCount * 8;
    const int iters = which == 0 ? 4096 : 2048;
    void* buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) peak_dfma_kernel<<<blocks, threads>>>((double*)buf, iters, 1.0);
        else            peak_mufu_kernel<<<blocks, threads>>>((float*)buf, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)threads * blocks * (double)iters * (which == 0 ? 64.0 : 32.0);
        const double rate = instr / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(BJ_ERR_CUDA, "peak kernel failed: %s", cudaGetErrorString(e));
    *out = best;
    return BJ_OK;
}
