#!/usr/bin/env python
"""
bench.py -- headline benchmark of the batched GW likelihood (BASELINE.json metric), all five BASELINE configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--secondary all|none|C1,C3,..]

HEADLINE (config C2: TaylorF2 3.5PN + 6PN tides BNS, H1/L1/V1, 128 s, 20-4096 Hz = 521 729 bins, 4096 walkers per GPU).
A "step" is ONE batched ``GWLikelihood.log_like`` over the walker batch: prologue, classification of the walkers by
phase slope, and per frequency table (compressed axes of DESIGN.md section 4.9, then the full grid) FP64 window kernel,
fused tensor-core kernel, epilogue.  The metric counts the reference's bins (521 729 per walker); the default path
evaluates them through ~22 600 Chebyshev nodes whose pre-summed data weights reproduce the bin sum to ~1e-11 (parity
against the oracle and the all-FP64 full-grid kernel is part of the line); ``full_grid`` is the same step with every
bin evaluated.

  value     walker*bins/s over all ranks, inputs resident in HBM, timed per step with CUDA events on the
            launching stream (L2 flushed between steps, outside the timed events), max over ranks
  e2e       N = 1: the host-pointer C-ABI entry ``bj_loglike`` (pinned host X -> H2D -> kernels -> D2H logL);
            N > 1: ``ShardedEvaluator.evaluate`` on the sampler rank, no collective: host X -> ONE H2D copy into the sampler
            rank's HBM -> every rank's prologue reads its rows and every rank's epilogue stores its logL there over NVLink
            (peer memory, flag-ordered) -> D2H
  strong    (N > 1) the same two numbers with the TOTAL batch fixed at 4096 walkers (a sampler's ensemble does not grow
            with the GPU count)
  full_grid the same step with the compressed frequency axis switched off (every bin evaluated: the round-1 / round-2a
            product path, still the path of contexts with calibration envelopes, PSD weights, short segments, and of
            every walker outside the design of the compression), with its own roofline
  focused   the same step after ``like.focus(injection)``: walkers whose phase evolution stays close to a fiducial point's are
            evaluated on a few hundred all-FP64 nodes (what a converged sampler gets); reported beside the headline, which
            assumes nothing about where the sampler is
  roofline  the fused kernel against the slower of the FP64-issue and SFU-issue rooflines (HBM is listed too, it is far
            from binding): achieved = algorithmic instructions per walker*bin x units / kernel
            time; peaks = DFMA and MUFU issue rates measured live on this GPU
  cpu_baseline  the oracle port (NumPy restatement of the reference algorithm, oracle/bajes_port.py) on the host cores,
            multiprocessing over all of them like bajes does, on a bounded sample of the SAME walkers (the reference
            itself is Python and cannot travel to the GPU box)
  secondary the other BASELINE configs, each with value / e2e / roofline / parity against the oracle:
            C1  TaylorF2 3.5PN BBH, H1+L1, 4 s, 64 walkers (launch-latency bound)
            C3  ROQ, 65 536 points, N_L = 400 / N_Q = 200 JenpyROQ-format basis, 10^4 tc points (HBM bound)
            C4  relative binning, 256 s, 2^20 points sharded over the N ranks
            C5  stretch-move ensemble sampler on the C2 data, 4096 walkers in all, through BatchedPool ->
                ShardedEvaluator (the unchanged-sampler contract: pool.map(posterior.log_post, points))

``--impl reference`` times the CPU implementation alone and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "likelihood evals/sec (walkers x freq bins / s)"
UNIT = "walker*bins/s"
DTYPE = ("f64 phase (DMMA) -> u32 binary angle -> f32 phasor (1 MUFU pair per bin, rotated to the other detectors) + f32 MAC "
         "-> f64 accumulate every 8 bins; all-f64 in the heaviest 1.6% of the bins")
DTYPE_COMPRESSED = ("f64 data weights pre-summed onto Chebyshev nodes (long double) -> f64 phase (DMMA) -> u32 binary angle -> f32 "
                    "phasor (MUFU pair per node and detector) + f32 MAC -> f64 accumulate every 8 nodes; all-f64 in the heaviest "
                    "16% of the nodes")
LAUNCHES_PER_STEP = 5      # prologue, FP64 window, fused tile kernel, fix-up (early exit), epilogue
# compressed axes: prologue, classify, then ONE window, tile, fix-up (early exit) and epilogue launch over all tables
LAUNCHES_PER_STEP_COMPRESSED = 6


# Algorithmic instruction counts per walker*bin of the fused kernel:
#   FP64 : K = 12 / 28 FMAs of the phase GEMM (executed as DMMA.8x8x4) + per 8 bins: 2 flush FMAs for detector 0 and
#          10 for every rotated detector (apply and advance its FP64 anchor) + 1 ramp anchor
#   SFU  : ONE MUFU.SIN + MUFU.COS per bin with the rotation path (uniform grids), else one pair per detector
def instr_per_unit(approx, n_ifo, sincos, tensor_core=True, rot=True):
    phase = 13 if approx == "TaylorF2_3.5PN" else 27
    if sincos == "fp64":
        return phase + 7 + 5 * n_ifo, 0
    if tensor_core:
        fp64 = (12 if approx == "TaylorF2_3.5PN" else 28) + ((3.0 + 10.0 * (n_ifo - 1)) / 8.0 if rot else 3.0 * n_ifo / 8.0)
    else:
        fp64 = phase + n_ifo + 2.0 * n_ifo / 8.0
    sfu = (2 if rot and n_ifo > 1 else 2 * n_ifo) if sincos == "mufu" else 0
    return fp64, sfu


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: the config's)")
    ap.add_argument("--sincos", default=os.environ.get("BAJES_B200_SINCOS", "mufu"))
    ap.add_argument("--cpu-points", type=int, default=0, help="points of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", default="all", help="all | none | comma list of C1,C3,C4,C5")
    ap.add_argument("--no-compress", action="store_true", help="evaluate every bin (compressed frequency axis off)")
    ap.add_argument("--no-focus", action="store_true", help="skip the focused sub-measurement")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores
# ---------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(cfg_name):
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"                        # as bajes does for its workers (bajes/__main__.py:35-37)
    from bajes_b200 import synthetic as syn
    from oracle import harness
    cfg = syn.CONFIGS[cfg_name]
    _CPU["like"] = harness.port_likelihood(cfg)
    _CPU["names"] = syn.WALKER_NAMES
    _CPU["const"] = syn.constants(cfg)


def _cpu_eval(x):
    p = {k: float(v) for k, v in zip(_CPU["names"], x)}
    p.update(_CPU["const"])
    return _CPU["like"].log_like(p)


class CpuArm:
    def __init__(self, cfg_name, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(cfg_name,))
        self.pool.map(_cpu_eval, [])            # spin up

    def run(self, X):
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_eval, list(X), chunksize=max(1, len(X) // (4 * self.cores)))
        return np.array(out), time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_points_auto(cfg, cores):
    # ~0.4 us per bin per detector per point on one core (measured: 0.2 s at 521 729 x 3)
    per_point = 1.3e-7 * cfg_bins(cfg) * len(cfg["ifos"]) + 1e-3
    return int(max(cores, min(4096, cores * max(1.0, 12.0 / per_point))))


def cfg_bins(cfg):
    return int(round((cfg["f_max"] - cfg["f_min"]) * cfg["seglen"])) + 1


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, 50 ms) -- runs during the timed region
# ---------------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock, throttle reasons and power of this rank's GPU, sampled INSIDE the timed region but BETWEEN the event-timed
    steps: one in-process NVML query after every step's synchronisation (a step is ~0.3 ms now; a polling nvidia-smi
    process holds a driver lock for longer than that whenever it fires, and did show up as 0.1 ms on the sampling rank's
    mean step).  Falls back to an ``nvidia-smi -lms`` subprocess when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.first = 0
        self.nv = None
        self.samples = []

    def start(self):
        if os.environ.get("BENCH_NO_CLOCKS"):
            return
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = (pynvml, h)
            self.sample()
            return
        except Exception:
            self.nv = None
        try:
            # nvidia-smi attaches to every GPU of the box when it starts (seconds on an 8-GPU node, and kernel launches of
            # all ranks stall meanwhile), so that must be over before the timed region begins
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 20.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def sample(self):
        """one NVML query (between two timed steps)"""
        if self.nv is None:
            return
        nv, h = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            self.samples.append((float(sm), float(mx), int(mask), float(pw)))
        except Exception:
            pass

    def mark(self):
        """Start of the timed region: only samples from here on (and the one just before) are reported."""
        self.first = max(0, len(self.rows) - 1)
        self.first_sample = max(0, len(self.samples) - 1)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nv is not None:
            nv, _ = self.nv
            rows = self.samples[getattr(self, "first_sample", 0):]
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, b in bits.items() if any(r[2] & b for r in rows))
            return {"sm_mhz": float(np.median([r[0] for r in rows])) if rows else None,
                    "sm_max_mhz": max(r[1] for r in rows) if rows else None, "reasons": reasons,
                    "power_w_max": max(r[3] for r in rows) if rows else None, "samples": len(rows),
                    "source": "NVML in process, one query after every timed step"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.08)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows[self.first:]:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[6]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "source": "nvidia-smi -lms 50"}


# ---------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """--impl reference: the CPU implementation of the path, all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bajes_b200 import synthetic as syn
    cfg = syn.CONFIGS[args.config]
    cores = os.cpu_count() or 1
    n = args.cpu_points or cpu_points_auto(cfg, cores)
    # keep the whole run within a few minutes
    n = max(cores, int(n / max(1, (args.steps + args.warmup) / 3)))
    X, _ = syn.draw_walkers(cfg, n, 5, "narrow")
    arm = CpuArm(args.config, cores)
    for _ in range(min(args.warmup, 1)):
        arm.run(X[:cores])
    times = []
    for _ in range(args.steps):
        _, dt = arm.run(X)
        times.append(dt)
    arm.close()
    nb = cfg_bins(cfg)
    total = float(np.sum(times))
    value = n * nb * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload(cfg, args, args.walkers or cfg["n_walkers"]),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d walkers x %d bins x %d detectors per step, oracle/bajes_port.py over a "
                                       "%d-process pool" % (n, nb, len(cfg["ifos"]), cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "evals_per_s": value / nb}
    print(json.dumps(line), flush=True)


def workload(cfg, args, walkers):
    return {"workload": "%s: %s, %s, %g s segment, %g-%g Hz (%d bins), %d walkers/GPU"
                        % (args.config, cfg["approx"], "/".join(cfg["ifos"]), cfg["seglen"], cfg["f_min"], cfg["f_max"],
                           cfg_bins(cfg), walkers),
            "approx": cfg["approx"], "ifos": cfg["ifos"], "n_bins": cfg_bins(cfg), "walkers_per_gpu": walkers,
            "sharding": ("walkers, dp%d, no collective: X read from / logL stored into the sampler rank's HBM by every rank's own "
                         "kernels (peer memory) -- inside e2e" % args.gpus) if args.gpus > 1
                        else "walkers, dp1",
            "l2": "flushed between timed steps (256 MiB memset, untimed)"}


def source_sha():
    """hash of the kernel sources: profiles/traffic.json carries the hash it was measured at"""
    h = hashlib.sha1()
    for name in ("kernels.cuh", "walker.cuh", "roq_relbin.cuh", "roq_basis.cuh", "capi.cu"):
        with open(os.path.join(ROOT, "bajes_b200", "csrc", name), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture -- only if it was taken at these very
    kernel sources (else null: a stale figure is worse than none)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None, "no profiles/traffic.json"
    if t.get("source_sha") != source_sha():
        return None, "profiles/traffic.json was captured at kernel sources %s, these are %s" % (t.get("source_sha"), source_sha())
    return t.get(key), t.get("_note")


# ---------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def timed(fn, dev, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms))


def wall(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return float(np.median(t)) * 1e3


def max_over_ranks(c, v):
    import torch
    if c.world == 1:
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=c.dev)
    c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return float(t.item())


def make_evaluator(c, like, names, const, capacity):
    """ShardedEvaluator over this likelihood (collective constructor)."""
    import torch
    from bajes_b200.dist import ShardedEvaluator
    pmap = like.param_map(names, const)

    def eval_dev(xs, outs):
        like.engine.loglike_device(xs.data_ptr(), xs.shape[0], len(names), pmap, outs.data_ptr(), 0,
                                   torch.cuda.current_stream(c.dev).cuda_stream)

    return ShardedEvaluator(lambda Xs: like.log_like_batch(Xs, names, const), ndim=len(names), device=c.dev,
                            local_eval_device=eval_dev, capacity=capacity)


def build_roofline(cfg, args, nifo, W, units, info, k_s, peaks, pk, tensor_core, rot, traffic_key, comp, interval_s=None,
                   interval_units=None):
    """Roofline object of the dominant (tile) kernel.  `units` = walker x (bins or nodes) IT evaluates, k_s its own duration;
    comp = None for the full grid, else {"counts": walkers per table, "nodes": tile-kernel nodes per table, "n_bins": ..} for a
    context with compressed axes.  interval_s / interval_units: everything between the prologue's and the epilogue's events."""
    fp64_pu, sfu_pu = instr_per_unit(cfg["approx"], nifo, args.sincos, tensor_core, rot)
    if comp:
        fp64_pu += nifo                    # non-uniform axis: the time-delay ramp is one DFMA per node and detector
    n_wtiles = (W + 63) // 64 if tensor_core else (W + 127) // 128
    # algorithmic bytes: every walker tile streams the static table once; coefficient block; partial sums
    if comp:
        tiles = [(comp["counts"][lv] + 63) // 64 for lv in range(3)]
        alg_bytes = (tiles[0] * comp["nodes"][0] + tiles[1] * comp["nodes"][1] + tiles[2] * comp["n_bins"]) * info[6] * 8 \
            + W * 71 * 8 + info[5] * 2 * nifo * W * 8
    else:
        alg_bytes = n_wtiles * (units // max(1, W)) * info[6] * 8 + W * 71 * 8 + info[5] * 2 * nifo * W * 8
    traffic, traffic_note = measured_traffic(traffic_key)
    lines = {"fp64": (fp64_pu, peaks["dfma"], "GDFMA/s"), "sfu": (sfu_pu, peaks["mufu"], "GMUFU/s")}
    # the limiting roofline of the two the north star names = the one that needs the longer time
    bound = max(lines, key=lambda k: lines[k][0] / lines[k][1])
    ipu, peak, unit = lines[bound]
    achieved = units * ipu / k_s
    other = "sfu" if bound == "fp64" else "fp64"
    K = 12 if cfg["approx"] == "TaylorF2_3.5PN" else 28
    if comp:
        kernel = ("fullgrid_tile_kernel<%d ifo, %s, %s, non-uniform axis>, ONE launch over all tables: CUDA events around the "
                  "launch (bj_last_tile_ms)" % (nifo, cfg["approx"], args.sincos))
        note = ("executed work of the tile kernel alone.  The step is no longer one long kernel: `interval` is everything "
                "between the prologue's and the epilogue's events -- classification, the FP64 window (16 % of the nodes at "
                "~2.7x the cost per node), this kernel, the early-exit fix-up launch; the full-grid kernel's own fraction is "
                "in full_grid.roofline")
        units_note = ("walker x node pairs the tile kernel evaluates: %s walkers on the level-0 / level-1 / full-grid tables, "
                      "%d / %d / %d slots each outside the FP64 window" % (comp["counts"], comp["nodes"][0], comp["nodes"][1],
                                                                          comp["n_bins"]))
    else:
        kernel = ("fullgrid_%s_kernel<%d ifo, %s, %s%s>: CUDA events around the launch (bj_last_tile_ms); `interval` adds the FP64 "
                  "window kernel and the early-exit fix-up launch"
                  % ("tile" if tensor_core else "mixed" if args.sincos != "fp64" else "fp64", nifo, cfg["approx"], args.sincos,
                     ", rotated detectors" if rot else ""))
        note = ("what limits this kernel is neither pipe but instruction issue / register-operand bandwidth (DESIGN.md "
                "section 4.1): with one MUFU pair per bin the SFU line no longer measures it; the FP64 line (12 of its 13.9 "
                "FMAs per unit run as DMMA.8x8x4) is the larger of the two the north star names")
        units_note = "walker x bins outside the FP64 window"
    interval = None
    if interval_s:
        interval = {"ms": interval_s * 1e3, "units": interval_units,
                    "frac_if_all_units_ran_at_tile_cost": interval_units * ipu / interval_s / peak}
    return {"bound": bound, "interval": interval, "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": unit,
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "instr_per_unit": ipu,
            "units_per_launch": units, "units": units_note, "kernel_ms": k_s * 1e3, "kernel": kernel,
            "peak_source": "measured live by bj_measure_peak (register-resident DFMA / MUFU chains); "
                           "MEASURED_PEAKS.json has no FP64 or SFU figure",
            "note": note,
            other: {"instr_per_unit": lines[other][0], "achieved": units * lines[other][0] / k_s / 1e9,
                    "peak": lines[other][1] / 1e9, "unit": lines[other][2],
                    "frac": units * lines[other][0] / k_s / lines[other][1]},
            "tensor": ({"achieved": units * 2 * K / k_s / 1e12, "peak": 2 * peaks["dfma"] / 1e12,
                        "unit": "TFLOP/s (FP64, DMMA)", "frac": units * K / k_s / peaks["dfma"]}
                       if tensor_core else None),
            "hbm": {"achieved": alg_bytes / k_s / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": alg_bytes / k_s / 1e9 / peaks["hbm"], "algorithmic_bytes": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in pk else "fallback"}}


def port_parity(port_like, got, X, names, const, n):
    from oracle import harness
    with np.errstate(all="ignore"):
        ref = harness.eval_batch(port_like, X[:n], names, const)
    ref = np.where(np.isfinite(ref), ref, -np.inf)
    fin = np.isfinite(ref)
    if not np.array_equal(np.isfinite(got[:n]), fin):
        return float("inf")
    tol = 1e-6 * np.abs(ref[fin]) + 1e-4
    return float(np.max(np.abs(got[:n][fin] - ref[fin]) / tol)) if fin.any() else 0.0


# ---------------------------------------------------------------------------------------------------------
# secondary configs
# ---------------------------------------------------------------------------------------------------------
def secondary_c1(c, peaks):
    """64 walkers x 4017 bins x 2 detectors on ONE GPU (rank 0): launch-latency bound."""
    import ctypes as C
    import torch
    from bajes_b200 import GWLikelihood, engine, synthetic as syn
    from oracle import harness
    if c.rank != 0:
        return None
    cfg = syn.CONFIGS["C1"]
    like = GWLikelihood(*syn.build_network(cfg, syn.product_classes()), cfg["srate"], cfg["seglen"], cfg["approx"],
                        device=c.local, sincos=c.args.sincos)
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    W, nb = cfg["n_walkers"], like.engine.n_bins
    X, _ = syn.draw_walkers(cfg, W, 5, "wide")
    pm = like.param_map(names, const)
    xd = torch.from_numpy(X).to(c.dev)
    out = torch.empty(W, dtype=torch.float64, device=c.dev)
    st = torch.cuda.current_stream(c.dev).cuda_stream
    step_ms, _ = timed(lambda: like.engine.loglike_device(xd.data_ptr(), W, len(names), pm, out.data_ptr(), 0, st), c.dev, 50, 5)
    kms = like.engine.last_kernel_ms()
    xp, op = torch.from_numpy(X).pin_memory(), torch.empty(W, dtype=torch.float64).pin_memory()
    lib = engine.load_library()
    host_ms = wall(lambda: lib.bj_loglike(like.engine._h, C.cast(xp.data_ptr(), C.POINTER(C.c_double)), W, len(names),
                                          C.byref(pm.c), C.cast(op.data_ptr(), C.POINTER(C.c_double))), 50, 5)
    parity = port_parity(harness.port_likelihood(cfg), out.cpu().numpy(), X, names, const, W)
    _, sfu = instr_per_unit(cfg["approx"], len(cfg["ifos"]), c.args.sincos)
    k_s = kms[1] * 1e-3
    like.engine.close()
    return {"workload": "C1: %s, H1/L1, 4 s, 20-1024 Hz (%d bins), %d walkers, one GPU" % (cfg["approx"], nb, W),
            "value": W * nb / (step_ms * 1e-3), "unit": UNIT, "evals_per_s": W / (step_ms * 1e-3), "ms_per_step": step_ms,
            "kernel_ms": [round(float(x), 5) for x in kms],
            "e2e": {"value": W * nb / (host_ms * 1e-3), "unit": UNIT, "ms_per_step": host_ms, "evals_per_s": W / (host_ms * 1e-3),
                    "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": W * 8, "path": "bj_loglike (host pointers)"},
            "roofline": {"bound": "launch latency (5 launches, 64 walkers fill 2 of 148 SMs per chunk row)",
                         "sfu": {"achieved": W * nb * sfu / k_s / 1e9, "peak": peaks["mufu"] / 1e9, "unit": "GMUFU/s",
                                 "frac": W * nb * sfu / k_s / peaks["mufu"]}},
            "parity_max_err_over_tol": parity, "parity_points": W, "n_gpus": 1}


def secondary_c3(c, peaks):
    """ROQ, 65 536 points sharded over the ranks; HBM-bound spline gather."""
    import torch
    from bajes_b200 import GWLikelihood, roq as roqmod, synthetic as syn
    from bajes_b200.ensemble import BoxPrior
    from oracle import harness
    points, n_lin, n_quad, tc_points = 65536, 400, 200, 10000
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=8.0, seed=3003)
    ifos, datas, dets, noises, f = syn.build_network(cfg, syn.product_classes())
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    basis = syn.make_roq_basis(cfg, n_lin=n_lin, n_quad=n_quad, seed=7)
    prior = BoxPrior(names, [[-1, 1]] * len(names), {})
    prior.bounds[names.index("time_shift")] = [-0.05, 0.05]
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        syn.write_jenpyroq(tmp, basis, cfg)
        roq = roqmod.build_roq_dict(tmp, tc_points, ifos, datas, noises, f, cfg["f_min"], cfg["f_max"], cfg["seglen"],
                                    prior, cfg["approx"], device=c.dev)
    setup_s = time.perf_counter() - t0
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"], roq=roq, device=c.local)
    nl, nq, nifo = len(roq["mask_omega"]), len(roq["mask_psi"]), len(ifos)
    pm = like.param_map(names, const)
    st = torch.cuda.current_stream(c.dev).cuda_stream
    ev = make_evaluator(c, like, names, const, points) if c.world > 1 else None
    from bajes_b200.dist import shard_bounds
    lo, hi = shard_bounds(points, c.world)
    n_loc = hi[c.rank] - lo[c.rank]
    res, parity, boxes = {}, None, {}
    # device-resident timings first (collective: max over ranks), then the end-to-end phase (sampler rank vs serve())
    for box in ("narrow", "wide"):
        X, _ = syn.draw_walkers(cfg, points, 9, box)
        boxes[box] = X
        xd = torch.from_numpy(np.ascontiguousarray(X[lo[c.rank]:hi[c.rank]])).to(c.dev)
        out = torch.empty(n_loc, dtype=torch.float64, device=c.dev)
        step_ms, _ = timed(lambda: like.engine.loglike_device(xd.data_ptr(), n_loc, len(names), pm, out.data_ptr(), 0, st),
                           c.dev, 5, 2)
        k_ms = like.engine.last_kernel_ms()[1]
        step_ms, k_ms = max_over_ranks(c, step_ms), max_over_ranks(c, k_ms)
        alg = n_loc * nifo * 4 * nl * 16
        res[box] = {"value": points / (step_ms * 1e-3), "ms_per_step": step_ms, "kernel_ms": k_ms, "e2e": None,
                    "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": alg / (k_ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                                 "algorithmic_bytes_per_point": nifo * 4 * nl * 16}}
        if box == "narrow" and c.rank == 0:
            port_like = harness.port_likelihood(cfg, roq=roq, net=syn.build_network(cfg, harness.port_classes()))
            parity = port_parity(port_like, out.cpu().numpy(), X[lo[0]:hi[0]], names, const, 48)
    if c.rank == 0:
        import ctypes as C
        from bajes_b200 import engine
        lib = engine.load_library()
        op = torch.empty(points, dtype=torch.float64).pin_memory()
        for box, X in boxes.items():
            xp = torch.from_numpy(X).pin_memory()         # pinned host input, as in every end-to-end arm of this file
            one = lambda: lib.bj_loglike(like.engine._h, C.cast(xp.data_ptr(), C.POINTER(C.c_double)), points, len(names),
                                         C.byref(pm.c), C.cast(op.data_ptr(), C.POINTER(C.c_double)))
            host_ms = wall(one if c.world == 1 else (lambda: ev.evaluate(xp)), 3, 1)
            res[box]["e2e"] = {"value": points / (host_ms * 1e-3), "unit": "points/s", "ms_per_step": host_ms,
                               "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": points * 8}
        if ev is not None:
            ev.shutdown()
    elif ev is not None:
        ev.serve()
    like.engine.close()
    if c.rank != 0:
        return None
    return {"workload": "C3: ROQ %s, H1/L1/V1, 8 s, N_L = %d / N_Q = %d nodes, %d tc points (spline table %.0f MB), %d points "
                        "sharded over %d GPU(s)" % (cfg["approx"], nl, nq, tc_points, tc_points * nl * 16 * nifo / 1e6, points, c.world),
            "unit": "points/s", "setup_s": setup_s, "narrow_box": res["narrow"], "wide_box": res["wide"],
            "value": res["wide"]["value"], "e2e": res["wide"]["e2e"], "roofline": res["wide"]["roofline"],
            "parity_max_err_over_tol": parity, "parity_points": 48, "n_gpus": c.world,
            "note": "headline numbers of this entry are the WIDE box (time shifts over the whole tc grid: every point gathers its "
                    "own four spline rows from HBM); the narrow box keeps part of the table in L2"}


def secondary_c4(c, peaks):
    """Relative binning, 256 s, 2^20 points sharded over the ranks."""
    import torch
    from bajes_b200 import synthetic as syn
    from bajes_b200.binning import GWBinningLikelihood
    from bajes_b200.dist import shard_bounds
    from oracle import bajes_port as port, harness
    cfg = syn.CONFIGS["C4"]
    points = 1 << 20
    ifos, datas, dets, noises, f = syn.build_network(cfg, syn.product_classes())
    fid = syn.injection_params(cfg)
    t0 = time.perf_counter()
    like = GWBinningLikelihood(ifos, datas, dets, noises, dict(fid), f, cfg["srate"], cfg["seglen"], cfg["approx"], device=c.local)
    setup_s = time.perf_counter() - t0
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    X, _ = syn.draw_walkers(cfg, points, 11, "narrow")
    lo, hi = shard_bounds(points, c.world)
    n_loc = hi[c.rank] - lo[c.rank]
    pm = like.param_map(names, const)
    st = torch.cuda.current_stream(c.dev).cuda_stream
    xd = torch.from_numpy(np.ascontiguousarray(X[lo[c.rank]:hi[c.rank]])).to(c.dev)
    out = torch.empty(n_loc, dtype=torch.float64, device=c.dev)
    step_ms, _ = timed(lambda: like.engine.loglike_device(xd.data_ptr(), n_loc, len(names), pm, out.data_ptr(), 0, st), c.dev, 5, 2)
    kms = like.engine.last_kernel_ms()
    step_ms, k_ms, p_ms = max_over_ranks(c, step_ms), max_over_ranks(c, kms[1]), max_over_ranks(c, kms[0])
    host_ms = None
    if c.world == 1:
        import ctypes as C
        from bajes_b200 import engine
        lib = engine.load_library()
        xp, op = torch.from_numpy(X).pin_memory(), torch.empty(points, dtype=torch.float64).pin_memory()
        host_ms = wall(lambda: lib.bj_loglike(like.engine._h, C.cast(xp.data_ptr(), C.POINTER(C.c_double)), points, len(names),
                                              C.byref(pm.c), C.cast(op.data_ptr(), C.POINTER(C.c_double))), 3, 1)
    else:
        ev = make_evaluator(c, like, names, const, points)
        if c.rank == 0:
            xp = torch.from_numpy(X).pin_memory()        # pinned like the single-GPU arm: no staging copy
            host_ms = wall(lambda: ev.evaluate(xp), 3, 1)
            ev.shutdown()
        else:
            ev.serve()
    parity = None
    if c.rank == 0:
        net_p = syn.build_network(cfg, harness.port_classes())
        ref = port.GWBinningLikelihoodPort(net_p[0], net_p[1], net_p[2], net_p[3], dict(fid), net_p[4], cfg["srate"],
                                           cfg["seglen"], cfg["approx"])
        parity = port_parity(ref, out.cpu().numpy(), X[lo[0]:hi[0]], names, const, 64)
    nbin, nifo = int(like.Nbin), len(ifos)
    fp64_per_point = (nbin + 1) * (24 + 10 * nifo) + 300
    like.engine.close()
    if c.rank != 0:
        return None
    return {"workload": "C4: relative binning %s, H1/L1/V1, 256 s, %d bins (chi = 1, eps = 0.5), 2^20 points sharded over %d "
                        "GPU(s) (%d per GPU)" % (cfg["approx"], nbin, c.world, n_loc),
            "unit": "points/s", "setup_s": setup_s, "value": points / (step_ms * 1e-3), "ms_per_step": step_ms,
            "kernel_ms": {"prologue": p_ms, "relbin": k_ms},
            "e2e": {"value": points / (host_ms * 1e-3), "unit": "points/s", "ms_per_step": host_ms,
                    "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": points * 8,
                    "path": "bj_loglike (pinned host pointers)" if c.world == 1 else "ShardedEvaluator.evaluate (scatter + peer stores)",
                    "note": "109 MB of parameters cross PCIe once per call: the end-to-end figure is a PCIe number"},
            "roofline": {"bound": "fp64", "achieved": n_loc * fp64_per_point / (k_ms * 1e-3) / 1e9, "peak": peaks["dfma"] / 1e9,
                         "unit": "GDFMA/s", "frac": n_loc * fp64_per_point / (k_ms * 1e-3) / peaks["dfma"],
                         "instr_per_point": fp64_per_point, "traffic": None,
                         "note": "libdevice sincospi per bin edge and detector counts as ~40 more FP64 instructions that the "
                                 "algorithmic figure does not include"},
            "parity_max_err_over_tol": parity, "parity_points": 64,
            "parity_note": "against the NumPy restatement; bins and summary data are pinned to the reference's own "
                           "setup_bins / compute_sdat (tests/golden/relbin_setup.npz), its log_like is dead code upstream",
            "n_gpus": c.world}


def secondary_c5(c, like, cfg, names, const):
    """Ensemble sampler (stretch move, the update bajes' emcee wrapper drives) on the C2 data: 4096 walkers IN ALL, every
    likelihood evaluation through pool.map(posterior.log_post, points) -> BatchedPool -> ShardedEvaluator."""
    import torch
    from bajes_b200 import Posterior, synthetic as syn
    from bajes_b200.dist import ShardedLikelihood
    from bajes_b200.ensemble import BoxPrior, EnsembleSampler
    from bajes_b200.pool import BatchedPool
    walkers, steps = 4096, 20
    p0, _ = syn.draw_walkers(cfg, walkers, 123, "narrow")
    lo, hi = p0.min(axis=0), p0.max(axis=0)
    pad = 0.5 * (hi - lo)
    bounds = np.stack([lo - pad, hi + pad], axis=1)
    bounds[names.index("q"), 0] = max(1.0, bounds[names.index("q"), 0])
    bounds[names.index("cos_iota")] = np.clip(bounds[names.index("cos_iota")], -1, 1)
    for k in ("lambda1", "lambda2"):
        bounds[names.index(k), 0] = max(0.0, bounds[names.index(k), 0])
    prior = BoxPrior(names, bounds, const)
    ev = make_evaluator(c, like, names, const, walkers) if c.world > 1 else None
    if c.rank != 0:
        ev.serve()
        return None

    def run(evaluator):
        target = ShardedLikelihood(like, evaluator) if evaluator is not None else like
        post = Posterior(target, prior)
        pool = BatchedPool(post, processes=c.world)
        s = EnsembleSampler(post, walkers, pool, seed=7)
        s.initialize(p0)
        torch.cuda.synchronize(c.dev)
        t0 = time.perf_counter()
        trace = s.run(steps)
        torch.cuda.synchronize(c.dev)
        return s, trace, time.perf_counter() - t0, pool

    s, trace, wall_s, pool = run(ev)

    # bajes' own sampler kernel: parallel-tempered ensemble (restatement of _PTMCMC pinned bit for bit to the unmodified
    # reference, bajes_b200/ptmcmc.py), 4 temperatures x 1024 walkers = 4096 points per sampler step, two batched calls
    def run_pt(evaluator, focus_after_warmup=False):
        from bajes_b200.ptmcmc import PTEnsembleSampler
        target = ShardedLikelihood(like, evaluator) if evaluator is not None else like
        ntemps, nw, pt_steps = 4, 1024, 20
        pt = PTEnsembleSampler(nw, Posterior(target, BoxPrior(names, bounds, const, periodic=[1 if n in ("psi", "phi_ref") else 0
                                                                                              for n in names])),
                               ntemps=ntemps, random=np.random.default_rng(11), rng="vector")
        start = p0[:ntemps * nw].reshape(ntemps, nw, len(names)).copy()
        pt.sample(iterations=1, p0=start, store=False)
        if focus_after_warmup:
            # what the sampler's refocus hook does every few hundred iterations (a focus table takes ~1 s to build)
            j = int(np.argmax(pt._loglikelihood0[0]))
            like.focus(pt._p0[0, j], names=names, const=const)
        torch.cuda.synchronize(c.dev)
        t0 = time.perf_counter()
        pt.sample(iterations=pt_steps, store=False)
        torch.cuda.synchronize(c.dev)
        return pt, time.perf_counter() - t0, pt_steps

    pt, pt_wall, pt_steps = run_pt(ev)
    same = pt_same = None
    if ev is not None:
        ev.shutdown()
        s1, _, _, _ = run(None)                    # the same seeded chain on this GPU alone
        same = bool(np.array_equal(s1.pos, s.pos) and np.array_equal(s1.logp, s.logp))
        pt1, _, _ = run_pt(None)
        pt_same = bool(np.array_equal(pt1._p0, pt._p0) and np.array_equal(pt1._loglikelihood0, pt._loglikelihood0))
    # the same sampler on this GPU alone with a focus table around its best cold point (untimed, as its refocus hook would
    # build one every few hundred iterations)
    ptf, ptf_wall, _ = run_pt(None, focus_after_warmup=True)
    ptf_counts = like.engine.last_level_counts()
    like.unfocus()
    pt_eval = pt.ntemps * pt.nwalkers * pt_steps
    ptmcmc = {"workload": "parallel-tempered ensemble (bajes_b200.ptmcmc.PTEnsembleSampler = bajes' _PTMCMC kernel, "
                          "bajes/inf/sampler/ptmcmc.py:184-470, stretch move), %d temperatures x %d walkers, %d steps, one batched "
                          "log_likeprior call per half-ensemble of all temperatures" % (pt.ntemps, pt.nwalkers, pt_steps),
              "value": pt_eval / pt_wall, "unit": "likelihood evaluations/s (end to end)", "s_per_sampler_step": pt_wall / pt_steps,
              "batched_calls": pt.n_batched_calls, "betas": [float(b) for b in pt.betas],
              "acceptance_per_temperature": [float(v) for v in pt.acceptance_fraction.mean(axis=1)],
              "tswap_acceptance": [float(v) for v in pt.tswap_acceptance_fraction],
              "lnL_max_cold_chain": float(np.max(pt._loglikelihood0[0])), "chain_bit_identical_to_one_gpu": pt_same,
              "with_a_focus_table_on_one_gpu": {"value": pt_eval / ptf_wall, "s_per_sampler_step": ptf_wall / pt_steps,
                                                  "last_call_walkers_level0_level1_fullgrid_focus": list(ptf_counts),
                                                  "lnL_max_cold_chain": float(np.max(ptf._loglikelihood0[0]))},
              "pinned_by": "tests/test_ptmcmc.py: bit-identical to the unmodified reference sampler (build container) and to "
                           "tests/golden/ptmcmc_trace.npz (its recorded run) in rng='reference' mode; this run draws in arrays"}
    idx = np.arange(0, walkers, 64)
    fresh = like.log_like_batch(s.pos[idx], names, const) + prior.log_prior(None)
    nb = like.engine.n_bins
    n_eval = walkers * steps
    return {"workload": "C5: stretch-move ensemble, %d walkers in all, %d steps on the C2 data (128 s BNS, %d bins x 3 detectors), "
                        "likelihood through pool.map(posterior.log_post) -> BatchedPool -> %s"
                        % (walkers, steps, nb, "ShardedEvaluator over %d GPUs" % c.world if c.world > 1 else "bj_loglike"),
            "unit": "likelihood evaluations/s (end to end, host in / host out every call)",
            "value": n_eval / wall_s, "walker_bins_per_s": n_eval * nb / wall_s, "s_per_sampler_step": wall_s / steps,
            "batched_calls": pool.n_batched_calls, "acceptance": s.acceptance,
            "logpost_max_first_last": [float(trace[0, 0]), float(trace[-1, 0])],
            "e2e": {"value": n_eval / wall_s, "unit": "evaluations/s", "ms_per_step": 1e3 * wall_s / steps,
                    "h2d_bytes_per_step": int(walkers * len(names) * 8), "d2h_bytes_per_step": walkers * 8},
            "roofline": None,
            "stored_logp_bit_identical_to_fresh_eval": bool(np.array_equal(fresh, s.logp[idx])),
            "chain_bit_identical_to_one_gpu": same, "parity_max_err_over_tol": None,
            "parity_note": "the likelihood is the headline's (parity there); here the check is that the chain is bit-identical "
                           "for any number of GPUs and that stored log-posteriors equal fresh evaluations",
            "n_gpus": c.world, "ptmcmc": ptmcmc,
            "note": "bajes' sampler packages are not installable offline on the GPU box.  The UNMODIFIED reference ptmcmc driven "
                    "through the same adapter runs in the build container (tests/test_sampler_adapters.py) and its recorded call "
                    "trace is replayed on the GPU (tests/test_gpu_parity.py); `ptmcmc` is its kernel restated in this repo"}


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from bajes_b200 import GWLikelihood, engine, synthetic as syn
    from bajes_b200.dist import DeviceShardedLikelihood

    c = Ctx()
    c.args = args
    c.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = rank = int(os.environ.get("RANK", "0"))
    c.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    c.dist = dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    c.dev = dev = torch.device("cuda", local)

    cfg = syn.CONFIGS[args.config]
    W = args.walkers or cfg["n_walkers"]
    nb = cfg_bins(cfg)
    nifo = len(cfg["ifos"])

    # ---- set-up (untimed): synthetic injection built with the product classes, likelihood on this GPU ------
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    like = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], device=local, sincos=args.sincos,
                        compress=False if args.no_compress else None)
    assert like.engine.n_bins == nb
    cinfo = [like.engine.compress_info(lv) for lv in range(2)]
    compressed = cinfo[0]["nodes"] > 0
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    Xall, _ = syn.draw_walkers(cfg, W * world, 5, "narrow")
    Xmine = np.ascontiguousarray(Xall[rank * W:(rank + 1) * W])
    sh = DeviceShardedLikelihood(like, names, const)
    x_dev = torch.from_numpy(Xmine).to(dev)
    out_dev = torch.empty(W, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    x_pin = torch.from_numpy(Xmine).pin_memory()
    out_pin = torch.empty(W, dtype=torch.float64).pin_memory()

    def step_device():
        sh.evaluate_local(x_dev, out_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing ---------------------------------------------------------------------------
    clocks = Clocks(local)
    clocks.start()                                 # every rank watches its own GPU (in-process NVML)
    barrier()
    # W warm-up steps as asked, and then as many more as it takes to keep the GPU busy for 0.3 s: a step is ~0.3 ms now,
    # three of them do not bring an idle GPU up to its boost clock (seen as one rank of four at 1.7x the others' step)
    n_warm, t_warm0 = 0, time.perf_counter()
    while n_warm < max(args.warmup, 3) or time.perf_counter() - t_warm0 < 0.3:
        step_device()
        n_warm += 1
        if n_warm % 16 == 0:
            torch.cuda.synchronize(dev)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms, tile_ms = [], []
    barrier()
    clocks.mark()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.zero_()                              # L2 flush, outside the timed events
        a.record()
        step_device()                              # prologue + FP64 window + fused kernel (+ fix-up) + epilogue
        b.record()
        b.synchronize()
        kern_ms.append(like.engine.last_kernel_ms()[1])
        tile_ms.append(like.engine.last_tile_ms())
        clocks.sample()                            # between two timed steps
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([float(np.sum(step_ms))], dtype=torch.float64, device=dev)
    per_rank = torch.tensor([float(np.mean(step_ms)), float(np.mean(kern_ms))], dtype=torch.float64, device=dev)
    per_rank_all = per_rank.clone().unsqueeze(0)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        per_rank_all = torch.empty((world, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(per_rank_all, per_rank)
    per_rank_all = per_rank_all.cpu().numpy().round(4).tolist()
    total_s = float(total_ms.item()) * 1e-3
    level_counts = like.engine.last_level_counts() if compressed else None

    # ---- the same step with every bin evaluated (compressed axis off), device resident + host pointers (rank-local) ----
    full_grid = None
    if compressed:
        like_full = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], device=local, sincos=args.sincos,
                                 compress=False)
        shf = DeviceShardedLikelihood(like_full, names, const)
        outf_dev = torch.empty(W, dtype=torch.float64, device=dev)
        fk, ftile = [], []

        def step_full():
            shf.evaluate_local(x_dev, outf_dev)

        for _ in range(3):
            step_full()
        fms = []
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_full()
            b.record()
            b.synchronize()
            fms.append(a.elapsed_time(b))
            fk.append(like_full.engine.last_kernel_ms()[1])
            ftile.append(like_full.engine.last_tile_ms())
        full_ms = max_over_ranks(c, float(np.mean(fms)))
        d = (out_dev - outf_dev).abs() / (1e-6 * outf_dev.abs() + 1e-4)
        d = torch.where(torch.isfinite(outf_dev), d, torch.zeros_like(d))
        assert bool(torch.equal(torch.isfinite(out_dev), torch.isfinite(outf_dev)))
        full_grid = {"value": W * world * nb / (full_ms * 1e-3), "unit": UNIT, "ms_per_step": full_ms,
                     "kernel_ms": float(np.mean(fk)), "evals_per_s": W * world / (full_ms * 1e-3),
                     "compressed_vs_full_grid_max_err_over_tol": float(d.max().item()),
                     "dtype": DTYPE}
        fpm = like_full.param_map(names, const)
        flib = engine.load_library()

        def step_full_host():
            rc = flib.bj_loglike(like_full.engine._h, C.cast(x_pin.data_ptr(), C.POINTER(C.c_double)), W, len(names),
                                 C.byref(fpm.c), C.cast(out_pin.data_ptr(), C.POINTER(C.c_double)))
            assert rc == 0, flib.bj_last_error()

        if world == 1:
            hms = wall(step_full_host, args.steps, 3)
            full_grid["e2e"] = {"value": W * nb / (hms * 1e-3), "unit": UNIT, "ms_per_step": hms,
                                "h2d_bytes_per_step": int(Xmine.nbytes), "d2h_bytes_per_step": int(W * 8),
                                "path": "bj_loglike (host pointers)"}
        full_info = like_full.engine.last_launch_info()
        full_kernel_s = float(np.mean(fk)) * 1e-3
        full_tile_s = float(np.mean(ftile)) * 1e-3
        full_tab = like_full.engine.table_info(2)
        like_full.engine.close()

    # ---- the same step with a FOCUS table around the injection (what a converged sampler gets) ------
    focused = None
    if compressed and not args.no_focus:
        finfo = like.focus(syn.injection_params(cfg))
        outc_dev = torch.empty(W, dtype=torch.float64, device=dev)

        def step_focus():
            sh.evaluate_local(x_dev, outc_dev)

        for _ in range(3):
            step_focus()
        cms, ck = [], []
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_focus()
            b.record()
            b.synchronize()
            cms.append(a.elapsed_time(b))
            ck.append(like.engine.last_kernel_ms()[1])
        focus_ms = max_over_ranks(c, float(np.mean(cms)))
        fcounts = like.engine.last_level_counts()
        d = (outc_dev - outf_dev).abs() / (1e-6 * outf_dev.abs() + 1e-4)
        d = torch.where(torch.isfinite(outf_dev), d, torch.zeros_like(d))
        focused = {"value": W * world * nb / (focus_ms * 1e-3), "unit": UNIT, "ms_per_step": focus_ms,
                   "kernel_ms": float(np.mean(ck)), "evals_per_s": W * world / (focus_ms * 1e-3),
                   "nodes": finfo["nodes"], "runs": finfo["blocks"],
                   "walkers_per_table_level0_level1_fullgrid_focus": list(fcounts),
                   "focused_vs_full_grid_max_err_over_tol": float(d.max().item()),
                   "dtype": "f64 throughout (data weights heterodyned with the fiducial point's phase, long double)",
                   "note": "GWLikelihood.focus(point): the band collapses into a few runs of 64 Chebyshev nodes for walkers "
                           "whose phase evolution stays within a bound of the fiducial's (here: the injection; the batch is the "
                           "posterior-scale box of SURVEY 8d); the others are routed as in the headline.  The headline itself "
                           "assumes nothing about where the sampler is"}
        if world == 1:
            def step_focus_host():
                rc = lib_f.bj_loglike(like.engine._h, C.cast(x_pin.data_ptr(), C.POINTER(C.c_double)), W, len(names),
                                      C.byref(pm_f.c), C.cast(out_pin.data_ptr(), C.POINTER(C.c_double)))
                assert rc == 0

            lib_f, pm_f = engine.load_library(), like.param_map(names, const)
            hms = wall(step_focus_host, args.steps, 3)
            focused["e2e"] = {"value": W * nb / (hms * 1e-3), "unit": UNIT, "ms_per_step": hms,
                              "h2d_bytes_per_step": int(Xmine.nbytes), "d2h_bytes_per_step": int(W * 8),
                              "path": "bj_loglike (host pointers)"}
        like.unfocus()

    # ---- end-to-end through the public API, host buffers in / out inside the timed region ----------------------
    pmap = like.param_map(names, const)
    lib = engine.load_library()

    def step_host():
        rc = lib.bj_loglike(like.engine._h, C.cast(x_pin.data_ptr(), C.POINTER(C.c_double)), W, len(names),
                            C.byref(pmap.c), C.cast(out_pin.data_ptr(), C.POINTER(C.c_double)))
        assert rc == 0, lib.bj_last_error()

    strong = None
    if world == 1:
        for _ in range(3):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        assert np.array_equal(out_pin.numpy(), out_dev.cpu().numpy()), "host and device entry points disagree"
        h2d, d2h = int(Xmine.nbytes), int(W * 8)
        e2e_path = "bj_loglike (host pointers)"
        peer = None
    else:
        she = make_evaluator(c, like, names, const, W * world)
        peer = she.peer is not None
        e2e_s = strong_e2e_s = 0.0
        # strong scaling, device resident: the same 4096 walkers in all, W / world per rank
        Ws = W // world
        xs_dev = x_dev[:Ws].contiguous()
        outs_dev = torch.empty(Ws, dtype=torch.float64, device=dev)
        strong_ms, _ = timed(lambda: sh.evaluate_local(xs_dev, outs_dev), dev, args.steps, 3)
        strong_ms = max_over_ranks(c, strong_ms)
        if rank == 0:
            xall_pin = torch.from_numpy(Xall).pin_memory()      # pinned host input, like the N = 1 arm
            for _ in range(3):
                got = she.evaluate(xall_pin)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                got = she.evaluate(xall_pin)
            e2e_s = time.perf_counter() - t0
            Xs = np.ascontiguousarray(Xall[:Ws * world])
            xs_pin = torch.from_numpy(Xs).pin_memory()
            for _ in range(3):
                got_s = she.evaluate(xs_pin)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                got_s = she.evaluate(xs_pin)
            strong_e2e_s = time.perf_counter() - t0
            she.shutdown()
            # EVERY rank's slice against this GPU alone: walker sharding must not change a bit
            alone = np.concatenate([like.log_like_batch(Xall[r * W:(r + 1) * W], names, const) for r in range(world)])
            assert np.array_equal(got, alone), "sharded results differ from the single-GPU evaluation of all %d rows" % len(alone)
            assert np.array_equal(got_s, alone[:Ws * world]), "strong-scaling results differ from the single-GPU evaluation"
            strong = {"walkers_total": Ws * world, "walkers_per_gpu": Ws,
                      "value": Ws * world * nb / (strong_ms * 1e-3), "ms_per_step": strong_ms,
                      "e2e": {"value": Ws * world * nb * args.steps / strong_e2e_s, "ms_per_step": 1e3 * strong_e2e_s / args.steps,
                              "h2d_bytes_per_step": int(Xs.nbytes), "d2h_bytes_per_step": int(Ws * world * 8)},
                      "note": "fixed total batch: what a sampler with a 4096-walker ensemble gets from N GPUs"}
        else:
            she.serve()
        h2d, d2h = int(Xall.nbytes), int(W * world * 8)
        e2e_path = ("ShardedEvaluator.evaluate, no collective: host X -> ONE H2D copy into the sampler rank's HBM -> every rank's "
                    "prologue reads its rows, and every rank's epilogue stores its logL, over NVLink peer memory (flag-ordered) -> D2H") \
            if peer else "ShardedEvaluator.evaluate: host X -> H2D -> NCCL scatter -> per-rank kernels -> NCCL gather -> D2H (no peer access)"
    clk = clocks.stop()
    if world > 1:
        mine = torch.tensor([float(clk.get("sm_mhz") or 0.0)], dtype=torch.float64, device=dev)
        allc = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allc, mine)
        clk["per_rank_sm_mhz"] = allc.cpu().numpy().tolist()
    clk["warmup_steps_run"] = n_warm
    if rank != 0:
        clk = None

    # ---- peaks and the secondary configs (collective: every rank takes part) -------------------------------------
    peaks = {"dfma": engine.measure_peak("dfma", local), "mufu": engine.measure_peak("mufu", local)}
    pk = {}
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peaks["hbm"] = float(pk.get("hbm_gbs", 6650.0))
    want = {"C1", "C3", "C4", "C5"} if args.secondary == "all" else set() if args.secondary == "none" else set(args.secondary.split(","))
    secondary = {}
    for key, fn in (("C1", lambda: secondary_c1(c, peaks)), ("C3", lambda: secondary_c3(c, peaks)),
                    ("C4", lambda: secondary_c4(c, peaks)), ("C5", lambda: secondary_c5(c, like, cfg, names, const))):
        if key in want and args.config == "C2":
            t0 = time.perf_counter()
            r = fn()
            if r is not None:
                r["wall_s_incl_setup"] = time.perf_counter() - t0
                secondary[key] = r

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------
    # the tile kernel alone: its own CUDA events (bj_last_tile_ms) and the units IT evaluates (table slots minus the FP64
    # window's); the whole interval between prologue and epilogue is reported beside it
    info = like.engine.last_launch_info()
    tensor_core = info[2] == 256           # fullgrid_tile_kernel: 256 threads = 64 walkers per CTA
    rot = tensor_core and nifo > 1 and not os.environ.get("BAJES_B200_NO_ROT")
    k_s = float(np.mean(kern_ms)) * 1e-3
    t_s = float(np.mean(tile_ms)) * 1e-3
    if compressed:
        ft = full_tab
        full_grid["roofline"] = build_roofline(cfg, args, nifo, W, W * (ft["slots"] - ft["window_slots"]), full_info,
                                               full_tile_s, peaks, pk, True, rot, args.config + ":" + args.sincos, None,
                                               interval_s=full_kernel_s, interval_units=W * nb)
        tabs = [like.engine.table_info(t) for t in range(3)]
        fast = [t["slots"] - t["window_slots"] for t in tabs]
        units = sum(level_counts[lv] * fast[lv] for lv in range(3))
        all_units = sum(level_counts[lv] * tabs[lv]["slots"] for lv in range(3))
        roofline = build_roofline(cfg, args, nifo, W, units, info, t_s, peaks, pk, tensor_core, False,
                                  args.config + ":" + args.sincos + ":compressed",
                                  {"counts": list(level_counts), "nodes": fast, "n_bins": fast[2]},
                                  interval_s=k_s, interval_units=all_units)
    else:
        tab = like.engine.table_info(2)
        roofline = build_roofline(cfg, args, nifo, W, W * (tab["slots"] - tab["window_slots"]), info, t_s, peaks, pk,
                                  tensor_core, rot, args.config + ":" + args.sincos, None, interval_s=k_s, interval_units=W * nb)
    value = W * world * nb * args.steps / total_s
    e2e_value = W * world * nb * args.steps / e2e_s

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = args.cpu_points or cpu_points_auto(cfg, cores)
        arm = CpuArm(args.config, cores)
        arm.run(Xmine[:cores])
        got, dt = arm.run(Xmine[:n])
        arm.close()
        cpu = {"value": n * nb / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of the %d walkers x %d bins x %d detectors, oracle/bajes_port.py over a %d-process pool, "
                         "%.1f s" % (n, W, nb, nifo, cores, dt)}
        ref = got
        mine = out_dev.cpu().numpy()[:n]
        tol = 1e-6 * np.abs(ref) + 1e-4
        cpu["parity_max_err_over_tol"] = float(np.max(np.abs(mine - ref) / tol))

    wl = workload(cfg, args, W)
    if compressed:
        wl["frequency_axis"] = ("compressed (default): %d bins -> %d Chebyshev nodes for walkers with chirp mass >= %.2f and "
                                "|delay + time_shift| <= %.2f s (level 1: %d nodes up to %.2f s; outside: every bin); all %d "
                                "walkers of this batch: %s on level 0 / level 1 / full grid.  `value` counts the reference's "
                                "bins; full_grid is the same step with every bin evaluated"
                                % (nb, cinfo[0]["nodes"], cinfo[0]["mchirp_min"], cinfo[0]["tau_max"], cinfo[1]["nodes"],
                                   cinfo[1]["tau_max"], W, list(level_counts)))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.sincos == "fp64" else DTYPE_COMPRESSED if compressed else DTYPE, "data": "synthetic",
            "config": wl, "evals_per_s": value / nb, "full_grid": full_grid, "focused": focused,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "path": e2e_path},
            "strong_scaling": strong,
            "gpu_launches": (LAUNCHES_PER_STEP_COMPRESSED if compressed else LAUNCHES_PER_STEP) * args.steps, "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
            "parity_max_err_over_tol": None if cpu is None else cpu["parity_max_err_over_tol"],
            "sincos": args.sincos, "wall_s_timed_region": t_wall,
            "per_rank_ms": {"step_and_fused_kernel": per_rank_all},
            "kernel_source_sha": source_sha(), "secondary": secondary}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
