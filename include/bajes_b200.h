/*
 * bajes_b200 -- C ABI of the B200-native batched gravitational-wave likelihood.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch
 * types.  The reference (matteobreschi/bajes v1.1.0) is pure Python and has no FFI; each
 * entry point below names the Python call it replaces (paths relative to the reference tree).
 * The ctypes binding a bajes maintainer would add is shown in INTEGRATION.md and shipped as
 * bajes_b200/engine.py.
 *
 * Conventions
 *   - every function returns BJ_OK (0) or a negative bj_status; bj_last_error() gives the text
 *     of the last failure on the calling thread.
 *   - numerical failure of a parameter point is NOT an error: it yields -inf in the output,
 *     exactly as GWLikelihood.log_like does (bajes/pipe/log_like.py:121-129).
 *   - a context owns device copies of all static tables; the caller owns X / out buffers.
 *   - one context per (process, device); calls on a context are serialised on its stream (or on
 *     the stream passed to the *_device entry points).  No global mutable state.
 *   - reductions use a fixed tree: results are bit-identical run to run and for any sharding of
 *     the walker batch across GPUs.
 */
#ifndef BAJES_B200_H
#define BAJES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BJ_ABI_VERSION 2
#define BJ_MAX_IFO 3            /* detectors fused in one pass of the kernel            */
#define BJ_MAX_SPCAL 16         /* spectral-calibration nodes per detector               */
#define BJ_MAX_WEIGHTS 16       /* PSD-weight bands per detector                         */

typedef enum {
    BJ_OK = 0,
    BJ_ERR_INVALID = -1,        /* bad argument                                          */
    BJ_ERR_CUDA = -2,           /* CUDA runtime failure (text in bj_last_error)          */
    BJ_ERR_UNSUPPORTED = -3,    /* feature outside the fused path (spcal, weights, ...)  */
    BJ_ERR_NO_DEVICE = -4       /* no CUDA device / not an sm_100 part                   */
} bj_status;

/* waveform.py:72-86 + taylorf2.py:435-520 (the five TaylorF2 wrappers) */
typedef enum {
    BJ_APPROX_TAYLORF2_35PN = 0,               /* 'TaylorF2_3.5PN'                       */
    BJ_APPROX_TAYLORF2_55PN = 1,               /* 'TaylorF2_5.5PN'                       */
    BJ_APPROX_TAYLORF2_55PN_75PNTIDES = 2,     /* 'TaylorF2_5.5PN_7.5PNTides'            */
    BJ_APPROX_TAYLORF2_55PN_35PNQM_75PNTIDES = 3, /* 'TaylorF2_5.5PN_3.5PNQM_7.5PNTides' */
    BJ_APPROX_TAYLORF2_55PN_75PNTIDES2020 = 4  /* 'TaylorF2_5.5PN_7.5PNTides2020'        */
} bj_approx;

/* sin/cos evaluation inside the full-grid kernel (phase is always FP64 and reduced exactly) */
typedef enum {
    BJ_SINCOS_MUFU = 0,         /* FP32 MUFU.SIN/COS (fastest)                           */
    BJ_SINCOS_POLY32 = 1,       /* FP32 minimax polynomials after octant reduction       */
    BJ_SINCOS_FP64 = 2          /* FP64 sincospi (validation)                            */
} bj_sincos;

/* canonical parameter slots (names as in bajes/pipe/gw_init.py:328-687, waveform.py:276-330) */
typedef enum {
    BJ_P_MCHIRP = 0, BJ_P_MTOT, BJ_P_Q,
    BJ_P_S1X, BJ_P_S1Y, BJ_P_S1Z, BJ_P_S2X, BJ_P_S2Y, BJ_P_S2Z,
    BJ_P_S1, BJ_P_TILT1, BJ_P_PHI_1L, BJ_P_S2, BJ_P_TILT2, BJ_P_PHI_2L,
    BJ_P_LAMBDA1, BJ_P_LAMBDA2,
    BJ_P_DISTANCE, BJ_P_COS_IOTA, BJ_P_IOTA,
    BJ_P_RA, BJ_P_DEC, BJ_P_PSI, BJ_P_TIME_SHIFT, BJ_P_PHI_REF, BJ_P_T_GPS,
    BJ_NUM_PARAMS
} bj_param;

/* How a batch row x[0..ndim) maps onto the slots: column >= 0 reads x[column]; BJ_COL_CONST
 * uses value[slot]; BJ_COL_ABSENT marks a slot that is not given at all (e.g. 'mtot' when the
 * sampler works in 'mchirp', cartesian vs spherical spins, 'iota' vs 'cos_iota').
 * Replaces Prior.this_sample + the dict parsing of Waveform.compute_hphc
 * (bajes/inf/prior.py:262-269, bajes/obs/gw/waveform.py:334-377). */
#define BJ_COL_CONST  (-1)
#define BJ_COL_ABSENT (-2)
typedef struct {
    int32_t column[BJ_NUM_PARAMS];
    double  value[BJ_NUM_PARAMS];
} bj_param_map;

/* per-detector constants, taken from the constructed Detector (detector.py:196-205) */
typedef struct {
    double response[9];         /* row-major 3x3 response tensor   (detector.py:207-211) */
    double location[3];         /* metres, Earth-fixed             (detector.py:213-224) */
    double gmst_reference;      /* rad at reference_time           (detector.py:235-242) */
    double reference_time;      /* GPS s                                                  */
    double sday;                /* sidereal day in s                                      */
} bj_detector;

typedef struct {
    int32_t abi_version;        /* must be BJ_ABI_VERSION                                 */
    int32_t device;             /* CUDA device ordinal                                    */
    int32_t approx;             /* bj_approx                                              */
    int32_t n_ifo;              /* 1..BJ_MAX_IFO                                          */
    int32_t marg_phi_ref;       /* log_like.py:175-177                                    */
    int32_t sincos;             /* bj_sincos                                              */
    double  seglen;             /* s                                                      */
    double  dd_sum;             /* sum over detectors of Detector._dd (detector.py:493)   */
    double  logZ_noise;         /* GWLikelihood.logZ_noise (log_like.py:94)               */
    bj_detector det[BJ_MAX_IFO];
} bj_config;

typedef struct bj_ctx bj_ctx;

const char* bj_last_error(void);
int         bj_abi_version(void);
/* name, SM count, SM clock (kHz) of `device`; BJ_ERR_NO_DEVICE without a GPU */
int         bj_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major,
                           int* cc_minor, int* clock_khz);

/* ------------------------------------------------------------------------------------------
 * Full-grid likelihood.
 * bj_create replaces GWLikelihood.__init__ + Detector.store_measurement
 * (pipe/log_like.py:27-105, obs/gw/detector.py:452-493): it uploads, per in-band bin k and
 * detector i, freqs[k], data_i[k] (complex, interleaved re/im) and psd_i[k] (already multiplied
 * by the window factor) and builds the fused device tables.
 * ------------------------------------------------------------------------------------------ */
int bj_create(bj_ctx** out, const bj_config* cfg, int64_t n_bins, const double* freqs,
              const double* const* data_ri, const double* const* psd);
int bj_destroy(bj_ctx* ctx);

/* Calibration envelopes and PSD weights (the nspcal / nweights options of GWLikelihood,
 * pipe/log_like.py:27-33; Detector.compute_inner_products, detector.py:517-531):
 *   spcal  : h_i(f) *= interp(f; spcal_freqs, (1 + amp_j) exp(i phi_j))      (detector.py:17-20,517-519)
 *   weights: S_i(f) *= w_b on nweights contiguous bands of len_weights[b] bins; dd is recomputed and
 *            logL gets -1/2 sum log w                                          (detector.py:22-23,524-531) */
typedef struct {
    int32_t nspcal;               /* 0 = no calibration envelope, else 2..BJ_MAX_SPCAL            */
    int32_t nweights;             /* 0 = no PSD weights, else 1..BJ_MAX_WEIGHTS                   */
    const double*  spcal_freqs;   /* [nspcal] ascending node frequencies                          */
    const int64_t* len_weights;   /* [nweights] bins per band, sum == n_bins                      */
    /* time-shift marginalisation (GWLikelihood(marg_time_shift=True), pipe/log_like.py:135-156): the per-bin
     * <d|h> integrand summed over detectors is placed on the full one-sided axis of n_full samples (zeros outside
     * the band, which starts at band_offset), Fourier transformed (numpy.fft.fft convention, length n_full) and
     * R = logsumexp_m(Re S[m] - ln n_full)  (or ln I0|S[m]| with marg_phi_ref).  Needs libcufft at run time
     * (loaded lazily); not combinable with spcal / weights / ROQ. */
    int32_t marg_time_shift;
    int32_t reserved;
    int64_t n_full;               /* len(series.freqs): T * srate / 2 + 1                          */
    int64_t band_offset;          /* index of the first in-band bin on that axis                   */
    /* Compressed frequency axis (no counterpart in the reference: an exact-to-1e-11 regrouping of the bin sum of
     * detector.py:541).  Runs of bins over which the phase of any walker INSIDE THE DESIGN turns by a bounded amount are
     * replaced by Chebyshev nodes with pre-summed data weights; the same kernels then evaluate ~10-30x fewer "bins" on
     * long segments.  Design = bound on the phase slope: chirp mass >= compress_mchirp_min (solar masses) and
     * |geocentric delay + time_shift| <= compress_tau_max (s).  Walkers outside the design are detected per call and
     * evaluated on the full grid, so results never leave the parity tolerance; only speed depends on the design.
     * compress: 0 = default (on, unless BAJES_B200_COMPRESS=0), 1 = on, -1 = off.  Zero / negative bounds = defaults
     * (0.8 solar masses -- two 0.92 M_sun stars, below the lightest neutron stars known -- and 0.11 s = the Earth's
     * light-travel time + 90 ms of time shift; the second level admits half the chirp mass and 1.1 s).  Ignored with spcal / weights / time marginalisation / sincos fp64 (full grid only). */
    int32_t compress;
    int32_t reserved2;
    double  compress_mchirp_min;
    double  compress_tau_max;
} bj_extras;

/* where the per-point values of 'spcal_amp{j}_{ifo}', 'spcal_phi{j}_{ifo}', 'weight{b}_{ifo}' live:
 * column >= 0 reads x[column], BJ_COL_CONST uses the value (same convention as bj_param_map) */
typedef struct {
    int32_t spcal_amp_col[BJ_MAX_IFO][BJ_MAX_SPCAL];
    int32_t spcal_phi_col[BJ_MAX_IFO][BJ_MAX_SPCAL];
    int32_t weight_col[BJ_MAX_IFO][BJ_MAX_WEIGHTS];
    double  spcal_amp_val[BJ_MAX_IFO][BJ_MAX_SPCAL];
    double  spcal_phi_val[BJ_MAX_IFO][BJ_MAX_SPCAL];
    double  weight_val[BJ_MAX_IFO][BJ_MAX_WEIGHTS];
} bj_extra_map;

/* bj_create with extras (ex may be NULL = bj_create). */
int bj_create_ex(bj_ctx** out, const bj_config* cfg, int64_t n_bins, const double* freqs,
                 const double* const* data_ri, const double* const* psd, const bj_extras* ex);

/* Batched GWLikelihood.log_like (pipe/log_like.py:107-183) for W points; x is row-major
 * [W, ndim] float64 in HOST memory; out_logl[W] in HOST memory.  Includes H2D/D2H copies. */
int bj_loglike(bj_ctx* ctx, const double* x_host, int64_t n_walkers, int32_t ndim,
               const bj_param_map* map, double* out_logl_host);

/* Same with DEVICE pointers, asynchronous on `stream` (a cudaStream_t, may be NULL = default).
 * out_parts (optional, may be NULL) receives per-detector [W][n_ifo][3] = Re<d|h>, Im<d|h>, <h|h>
 * (Detector.compute_inner_products, detector.py:496-544, summed over bins). */
int bj_loglike_device(bj_ctx* ctx, const double* x_dev, int64_t n_walkers, int32_t ndim,
                      const bj_param_map* map, double* out_logl_dev, double* out_parts_dev,
                      void* stream);

/* The same two entries for contexts created with extras; emap gives the calibration / weight parameters. */
int bj_loglike_ex(bj_ctx* ctx, const double* x_host, int64_t n_walkers, int32_t ndim,
                  const bj_param_map* map, const bj_extra_map* emap, double* out_logl_host);
int bj_loglike_device_ex(bj_ctx* ctx, const double* x_dev, int64_t n_walkers, int32_t ndim,
                         const bj_param_map* map, const bj_extra_map* emap, double* out_logl_dev,
                         double* out_parts_dev, void* stream);

/* Frequency-axis split (small batches on several GPUs): rank r evaluates the table chunks
 * [chunk_begin, chunk_end) of EVERY walker and returns the local partial sums
 * sums[2 * n_ifo][n_walkers] (Re, Im of the un-normalised <d|h> per detector, PSD weights applied).  The caller
 * all-reduces (sums) them across ranks and hands the result to bj_finish_device, which runs the epilogue
 * (<h|h> from the Gram moments, amplitude, logL).  bj_num_chunks gives the number of table chunks. */
int bj_num_chunks(bj_ctx* ctx, int32_t* n_chunks);
int bj_partial_device(bj_ctx* ctx, const double* x_dev, int64_t n_walkers, int32_t ndim,
                      const bj_param_map* map, const bj_extra_map* emap, int32_t chunk_begin, int32_t chunk_end,
                      double* out_sums_dev, void* stream);
int bj_finish_device(bj_ctx* ctx, const double* sums_dev, int64_t n_walkers, double* out_logl_dev,
                     double* out_parts_dev, void* stream);

/* Projected frequency-domain waveform of ONE parameter point on every detector
 * (Waveform.compute_hphc + Detector.project_fdwave, waveform.py:270-395, detector.py:394-418).
 * out_h_host: [n_ifo][n_bins][2] (re, im), FP64 sin/cos.  Used for injections / SNR helpers. */
int bj_project_waveform(bj_ctx* ctx, const double* x_host, int32_t ndim, const bj_param_map* map,
                        double* out_h_host);

/* h_plus / h_cross of ONE point on an arbitrary frequency axis, no detector, no data
 * (Waveform.compute_hphc; `centre_shift` applies exp(i pi f seglen), waveform.py:390-393).
 * out_hp/out_hc: [n_freqs][2]. */
int bj_polarizations(int device, int approx, double seglen, int centre_shift, int64_t n_freqs,
                     const double* freqs, const double* x_host, int32_t ndim,
                     const bj_param_map* map, double* out_hp_host, double* out_hc_host);

/* ------------------------------------------------------------------------------------------
 * ROQ likelihood (Detector.compute_inner_products roq branch, detector.py:534-537).
 * Static inputs come from the host set-up that mirrors pipe/utils/roq.py:817-872 and
 * pipe/__init__.py:661-716: joined node axis, index masks, psi weights, and the fitted cubic
 * B-spline of the omega weights (knots t[n_knots], coefficients c[n_coef][n_lin] complex).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t n_join;  const double* freqs_join;
    int64_t n_lin;   const int32_t* mask_omega;
    int64_t n_quad;  const int32_t* mask_psi;
    int64_t n_knots; int64_t n_coef;
    const double* const* knots;          /* [n_ifo][n_knots]                              */
    const double* const* coef_ri;        /* [n_ifo][n_coef][n_lin][2]                     */
    const double* const* psi_weights;    /* [n_ifo][n_quad]                               */
    double tc_min, tc_max;               /* validity range of the spline (outside: -inf)  */
} bj_roq_tables;

int bj_roq_create(bj_ctx** out, const bj_config* cfg, const bj_roq_tables* tab);
/* bj_loglike / bj_loglike_device / bj_destroy work on ROQ contexts too. */

/* ROQ basis CONSTRUCTION.  The reference ships only the run-time side of ROQ; its own construction code is commented out
 * (pipe/utils/roq.py:16-662).  These two entry points follow it: the training set (h_plus of TaylorF2 on the in-band axis,
 * no centre shift, :471-473; quadratic != 0: |h_plus|^2) and the reduced-basis greedy selection by projection error under
 * (a, b) = 4 df sum_k w_k conj(a_k) b_k  (rb_greedy :537-612, inner_product :654-662 -- Hermitian here: the commented code
 * keeps the real part only, which makes e and i e orthogonal and the complex DEIM that follows singular; tests/test_roq_basis.py).
 * DEIM (:614-652) and the interpolant stay on the host (bajes_b200/roq_basis.py).
 *   bj_wavebank : x_host [n_points][ndim] -> out_dev, DEVICE memory, complex128 [n_points][n_freqs]
 *   bj_rb_greedy: train_dev [n_train][n_freqs] complex128 DEVICE memory, overwritten by the residuals; basis_dev
 *                 [max_basis][n_freqs] holds n_prebuilt basis vectors on entry (roq.py:512-520: the rows are first reduced to
 *                 their residuals against them) and receives the orthonormal basis; out_picks / out_sigmas [max_basis + 1] the training
 *                 row picked at every step and its error before the pick (the entry after the last pick: the error at which
 *                 the loop ended); *out_n the basis size; *out_ms (optional) device time of the loop. */
int bj_wavebank(int device, int approx, int64_t n_freqs, const double* freqs_host, const double* x_host, int64_t n_points,
                int32_t ndim, const bj_param_map* map, int32_t quadratic, void* out_dev);
int bj_rb_greedy(int device, int64_t n_train, int64_t n_freqs, void* train_dev, const double* weights_host, double df,
                 double tol, int32_t n_prebuilt, int32_t max_basis, void* basis_dev, int32_t* out_picks, double* out_sigmas,
                 int32_t* out_n, float* out_ms);

/* ------------------------------------------------------------------------------------------
 * Relative-binning likelihood (GWBinningLikelihood.log_like, pipe/utils/binning.py:248-306).
 * Static inputs: bin-edge frequencies fbin[n_edges], fiducial waveform at the edges per
 * detector h0[n_ifo][n_edges][2], summary data sdat[4][n_ifo][n_bins][2] (A0,A1,B0,B1).
 * ------------------------------------------------------------------------------------------ */
int bj_relbin_create(bj_ctx** out, const bj_config* cfg, int64_t n_edges, const double* fbin,
                     const double* const* h0_ri, const double* sdat_ri);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU result path over peer memory (one process per GPU; replaces the gather of the per-point results to the
 * sampler rank, the "results in requested order" step of MPIPool.map, pipe/utils/mpi.py:120-159).
 * The sampler rank owns ONE device buffer [flags: 64 x int64][status: int64][pad][logL: capacity doubles]; the other
 * ranks map it with a CUDA IPC handle and pass `peer_base + offset` as out_logl_dev of bj_loglike_device: the
 * epilogue kernel then stores every logL straight into the sampler rank's HBM over NVLink.  bj_peer_signal (stream-
 * ordered, after the epilogue) publishes `value` in flags[slot] with a system-scope release; bj_peer_wait on the
 * sampler rank's stream spins (bounded by timeout_ms; status word = 1 on time-out) until flags[first..first+n) have
 * all reached `value`.  No collective and no host synchronisation on the return path.
 * ------------------------------------------------------------------------------------------ */
#define BJ_PEER_FLAGS 64
#define BJ_PEER_HEADER_BYTES ((BJ_PEER_FLAGS + 2) * 8)
int bj_peer_create(int device, int64_t capacity_doubles, void** base_dev, unsigned char* handle64);
int bj_peer_open(int device, const unsigned char* handle64, void** base_dev);
int bj_peer_close(int device, void* base_dev, int opened_from_handle);
int bj_peer_signal(int device, void* base_dev, int32_t slot, int64_t value, void* stream);
/* Sampler rank: copy n_parts byte ranges [part_begin, part_end) of a host block to base_dev + dst_offset (same offsets),
 * each followed by a release of `value` in flags[part_flag_slot[p]] -- one call, in the given order, so that the rank
 * whose rows landed can start while the next rows travel.  src_host pageable: pass a pinned staging block of the same
 * size (pieces of 2 MiB are staged while the previous piece is in flight); src_host pinned: stage_pinned = NULL. */
int bj_peer_scatter(int device, void* base_dev, int64_t dst_offset_bytes, const void* src_host, void* stage_pinned,
                    int32_t n_parts, const int64_t* part_begin, const int64_t* part_end,
                    const int32_t* part_flag_slot, int64_t value, void* stream);
int bj_peer_wait(int device, void* base_dev, int32_t first_slot, int32_t n_slots, int64_t value, double timeout_ms,
                 void* stream);

/* ------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py): issue-rate micro-benchmarks used as roofline denominators.
 * which: 0 = FP64 FMA (instr/s), 1 = MUFU sin+cos (instr/s).  Result in *out (thread-instr/s).
 * ------------------------------------------------------------------------------------------ */
int bj_measure_peak(int device, int which, double* out);

/* timing of the last bj_loglike* call on this context, measured with CUDA events on the
 * launching stream: ms[0] = prologue, ms[1] = main kernel, ms[2] = epilogue.  Synchronises. */
int bj_last_kernel_ms(bj_ctx* ctx, float* ms3);
/* duration of the dominant (tensor-core tile) kernel of the last call alone, CUDA events around its launch.  Synchronises. */
int bj_last_tile_ms(bj_ctx* ctx, float* ms);
/* table: 0, 1 = compressed levels, 2 = the full grid, 3 = the focus table.  out4 = {table slots, slots of the FP64 window,
 * chunks, chunks of the FP64 window} (zeros when the table does not exist): the tile kernel evaluates slots - window slots. */
int bj_table_info(bj_ctx* ctx, int32_t table, int64_t* out4);
/* static description of the last launch: grid, block, dynamic smem, bins per chunk, #chunks */
int bj_last_launch_info(bj_ctx* ctx, int32_t* info8);
/* Compressed frequency axes of a full-grid context (bj_extras.compress*).  A context holds up to BJ_COMPRESS_LEVELS node
 * axes of growing design budget (level 0: the design; level 1: half its chirp mass, |tau| up to max(1.1 s, 2 tau_max) -- the
 * default time-shift prior of the reference's pipeline) and the full grid; every call sorts its walkers into the first table that covers them.
 * info4 = {nodes, nodes that are plain bins, compressed runs, nodes per run} (all 0 when `level` does not exist),
 * spec4 = {mchirp_min, tau_max, reach in rad, slope margin}. */
#define BJ_COMPRESS_LEVELS 2
int bj_compress_info(bj_ctx* ctx, int32_t level, int32_t* info4, double* spec4);
/* table of every walker of the last call: levels[w] = 0, 1 (compressed levels), 2 (full grid) or 3 (focus).  Synchronises. */
int bj_last_levels(bj_ctx* ctx, int64_t n_walkers, int32_t* levels);
/* Host-only (no device needed): the node axis made of a weighted frequency axis -- weights_ri[i] = interleaved (re, im)
 * per bin and detector.  out_freqs / out_weights_ri[i] must hold n_bins (2 n_bins) doubles; *out_n = number of nodes.
 * Non-positive nodes_per_run / reach_rad / mchirp_min / tau_max / margin = the library defaults (64, 32, 0.8, 0.11, 1.3).
 * For any g(f) whose phase slope respects the design, sum_j W_j g(f_j) reproduces sum_k w_k g(f_k) to ~1e-11 of
 * sum_k |w_k g(f_k)| (tests/test_compress_host.py). */
int bj_compress_axis_host(int64_t n_bins, const double* freqs, int32_t n_ifo, const double* const* weights_ri,
                          int32_t nodes_per_run, double reach_rad, double mchirp_min, double tau_max, double margin,
                          int64_t* out_n, double* out_freqs, double* const* out_weights_ri, int32_t* info4);
/* walkers of the last call per table: counts4 = {level 0, level 1, full grid, focus table}.  Synchronises the device. */
int bj_last_level_counts(bj_ctx* ctx, int64_t* counts4);
/* Focus table: the compressed axis taken to its limit around ONE point -- the region a converged sampler lives in.  The
 * fiducial point's phase and per-detector delays are folded into the data weights, the kernels evaluate every walker
 * RELATIVE to it (the phase is linear in the per-walker coefficients: the subtraction is free), and the design bound is on
 * the relative slope,  |d(P - P_fid)/df| + max_i |tau_i - tau_fid,i| <= slope_a (f / f_lo)^(-5/3) + slope_b + slope_c
 * (f / f_hi)^(2/3)  seconds (a: masses and spins, b: sky position and time shift, c: tidal deformabilities; non-positive a, b
 * / negative c = defaults 0.25, 2e-3, 0.03): the band collapses into a few runs of 64 nodes (C2: 521 729 bins -> ~700),
 * evaluated in FP64 throughout.  Only walkers that respect the bound at every check frequency are evaluated there; all
 * others are routed exactly as without a focus (and get the same bits).  Relative binning (pipe/utils/binning.py) is the
 * first-order, unbounded version of this idea; this one is exact to the Chebyshev truncation (~1e-12).
 * x_host [ndim] = the fiducial point (same parameter map as the batches); info4 as bj_compress_info.  bj_unfocus drops it. */
int bj_focus(bj_ctx* ctx, const double* x_host, int32_t ndim, const bj_param_map* map, double slope_a, double slope_b,
             double slope_c, int32_t* info4);
int bj_unfocus(bj_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BAJES_B200_H */
