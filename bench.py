#!/usr/bin/env python
"""
bench.py -- headline benchmark of the batched GW likelihood (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2]

A "step" is ONE batched ``GWLikelihood.log_like`` over the configured walker batch (config C2: TaylorF2 3.5PN
+ 6PN tides BNS, H1/L1/V1, 128 s, 20-4096 Hz = 521 729 bins, 4096 walkers per GPU) -- prologue, fused
full-grid kernel, epilogue.

  value     walker*bins/s over all ranks, inputs resident in HBM, timed per step with CUDA events on the
            launching stream (L2 flushed between steps, outside the timed events), max over ranks
  e2e       same metric through the host-pointer C-ABI entry (``bj_loglike``): pinned host X -> H2D -> kernels
            -> D2H logL inside the timed region
  roofline  the fused kernel against the slower of the FP64-issue and SFU-issue rooflines (the HBM figure is
            listed too, it is far from binding): achieved = algorithmic instructions per walker*bin (DESIGN.md
            section 4) x units / kernel time; peaks = DFMA and MUFU issue rates measured live on this GPU
  cpu_baseline  the oracle port (NumPy restatement of the reference algorithm, oracle/bajes_port.py) on the
            host cores, multiprocessing over all of them like bajes does, on a bounded sample of the SAME
            walkers (the reference itself is Python and cannot travel to the GPU box)

``--impl reference`` times that CPU implementation alone and prints the same JSON line.
With N > 1 (torchrun) the walker batch is sharded: every rank evaluates its own 4096 walkers (weak scaling, no
data-path collective); the timed step is each rank's own work, max over ranks.  The NCCL all-gather of the per-walker
logL to the sampler rank runs between the timed events and is part of ``e2e`` (ShardedEvaluator.evaluate: host X ->
NCCL broadcast -> per-rank slices -> NCCL all-gather -> host).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "likelihood evals/sec (walkers x freq bins / s)"
UNIT = "walker*bins/s"

# Algorithmic instruction counts per walker*bin of the fused kernel:
#   FP64, Horner kernels : phase Horner (13 for 3.5PN + 6PN tides, 27 for the 5.5PN family) + one time-delay ramp FMA
#         per detector (+ 2 FMAs per detector per 8 bins for the FP32 -> FP64 flush)
#   FP64, tensor-core kernel : K = 12 / 28 FMAs of the phase GEMM (executed as DMMA.8x8x4) + per 8 bins and detector
#         2 flush FMAs and 1 FP64 ramp anchor (uniform grids; else one ramp FMA per bin and detector)
#   SFU : one MUFU.SIN + one MUFU.COS per detector ("mufu" mode; none in "poly32")
def instr_per_unit(approx, n_ifo, sincos, tensor_core=False):
    phase = 13 if approx == "TaylorF2_3.5PN" else 27
    if sincos == "fp64":
        fp64 = phase + 7 + 5 * n_ifo
        return fp64, 0
    if tensor_core:
        fp64 = (12 if approx == "TaylorF2_3.5PN" else 28) + 3.0 * n_ifo / 8.0
    else:
        fp64 = phase + n_ifo + 2.0 * n_ifo / 8.0
    sfu = 2 * n_ifo if sincos == "mufu" else 0
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
# clocks sampler (nvidia-smi, 200 ms) -- runs during the timed region
# ---------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.first = 0

    def start(self):
        """Launch the sampler and wait until it delivers: nvidia-smi attaches to every GPU of the box when it starts
        (seconds on an 8-GPU node, and kernel launches of all ranks stall meanwhile), so that must be over before the
        timed region begins."""
        if os.environ.get("BENCH_NO_CLOCKS"):
            return
        try:
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

    def mark(self):
        """Start of the timed region: only samples from here on (and the one just before) are reported."""
        self.first = max(0, len(self.rows) - 1)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
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
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm)}


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
            "sharding": "walkers, dp%d, no data-path collective (logL all-gather untimed here, inside e2e)" % args.gpus if args.gpus > 1 else "walkers, dp1", "l2": "flushed between timed steps (256 MiB memset, untimed)"}


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from bajes_b200 import GWLikelihood, engine, synthetic as syn
    from bajes_b200.dist import DeviceShardedLikelihood

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    cfg = syn.CONFIGS[args.config]
    W = args.walkers or cfg["n_walkers"]
    nb = cfg_bins(cfg)
    nifo = len(cfg["ifos"])

    # ---- set-up (untimed): synthetic injection built with the product classes, likelihood on this GPU ------
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    like = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], device=local, sincos=args.sincos)
    assert like.engine.n_bins == nb
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
    if rank == 0:
        clocks.start()
    barrier()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    barrier()
    clocks.mark()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.zero_()                              # L2 flush, outside the timed events
        a.record()
        step_device()                              # prologue + fused kernel (+ fix-up) + epilogue on this rank's walkers
        b.record()
        if world > 1:
            # The path shards by walkers with NO data-path collective: the timed step is each rank's own work.  The
            # logL all-gather to the sampler rank is a caller-side step; it runs here, untimed (8 x 32 KB), and it is
            # INSIDE the end-to-end number below.  (Inside the events it measured NCCL's small-message latency jitter
            # at 8 ranks -- 0.05 to 4 ms per call on this pool, all ranks' local work being done within 20 us.)
            sh.gather(out_dev)
        b.synchronize()
        kern_ms.append(like.engine.last_kernel_ms()[1])
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

    # ---- end-to-end through the public API, host buffers in / out inside the timed region ----------------------
    #   N = 1: the host-pointer C-ABI entry bj_loglike (pinned X -> H2D -> prologue/kernel/epilogue -> D2H logL)
    #   N > 1: ShardedEvaluator.evaluate on the sampler rank (host X -> device, NCCL broadcast, every rank evaluates
    #          its slice, NCCL all-gather, device -> host), the other ranks in serve()
    pmap = like.param_map(names, const)
    import ctypes as C
    lib = engine.load_library()

    def step_host():
        rc = lib.bj_loglike(like.engine._h, C.cast(x_pin.data_ptr(), C.POINTER(C.c_double)), W, len(names),
                            C.byref(pmap.c), C.cast(out_pin.data_ptr(), C.POINTER(C.c_double)))
        assert rc == 0, lib.bj_last_error()

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
    else:
        from bajes_b200.dist import ShardedEvaluator

        def eval_dev(xs, outs):
            like.engine.loglike_device(xs.data_ptr(), xs.shape[0], len(names), pmap, outs.data_ptr(), 0,
                                       torch.cuda.current_stream(dev).cuda_stream)

        she = ShardedEvaluator(lambda Xs: like.log_like_batch(Xs, names, const), ndim=len(names), device=dev,
                               local_eval_device=eval_dev)
        e2e_s = 0.0
        if rank == 0:
            for _ in range(3):
                got = she.evaluate(Xall)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                got = she.evaluate(Xall)
            e2e_s = time.perf_counter() - t0
            she.shutdown()
            assert np.array_equal(got[:W], out_dev.cpu().numpy()), "sharded and local results disagree"
        else:
            she.serve()
        h2d, d2h = int(Xall.nbytes), int(W * world * 8)
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------
    units_per_launch = W * nb
    info = like.engine.last_launch_info()
    tensor_core = info[2] == 256           # fullgrid_tile_kernel: 256 threads = 64 walkers per CTA
    fp64_pu, sfu_pu = instr_per_unit(cfg["approx"], nifo, args.sincos, tensor_core)
    k_s = float(np.mean(kern_ms)) * 1e-3
    peak_dfma = engine.measure_peak("dfma", local)
    peak_mufu = engine.measure_peak("mufu", local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n_wtiles = (W + 63) // 64 if tensor_core else (W + 127) // 128
    # algorithmic bytes: every walker tile streams the static table once; coefficient block; partial sums
    alg_bytes = n_wtiles * nb * info[6] * 8 + W * 50 * 8 + info[5] * 2 * nifo * W * 8
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.config + ":" + args.sincos)
        except Exception:
            traffic = None
    lines = {"fp64": (fp64_pu, peak_dfma, "GDFMA/s"), "sfu": (sfu_pu, peak_mufu, "GMUFU/s")}
    # the limiting roofline of the two the north star names = the one that needs the longer time
    bound = max(lines, key=lambda k: lines[k][0] / lines[k][1])
    ipu, peak, unit = lines[bound]
    achieved = units_per_launch * ipu / k_s
    other = "sfu" if bound == "fp64" else "fp64"
    roofline = {"bound": bound, "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": unit,
                "frac": achieved / peak, "traffic": traffic, "instr_per_unit": ipu,
                "units_per_launch": units_per_launch, "kernel_ms": k_s * 1e3,
                "kernel": "fullgrid_%s_kernel<%d ifo, %s, %s>" % (
                    "tile" if tensor_core else "mixed" if args.sincos != "fp64" else "fp64", nifo, cfg["approx"], args.sincos),
                "peak_source": "measured live by bj_measure_peak (register-resident DFMA / MUFU chains); "
                               "MEASURED_PEAKS.json has no FP64 or SFU figure",
                other: {"instr_per_unit": lines[other][0], "achieved": units_per_launch * lines[other][0] / k_s / 1e9,
                        "peak": lines[other][1] / 1e9, "unit": lines[other][2],
                        "frac": units_per_launch * lines[other][0] / k_s / lines[other][1]},
                # the phase GEMM alone, in the units of a tensor-core roofline: K FMAs per walker*bin on DMMA.8x8x4 against
                # the FP64 rate measured on this GPU (DMMA and DFMA share it: scripts/ubench_dmma.cu)
                "tensor": ({"achieved": units_per_launch * 2 * (12 if cfg["approx"] == "TaylorF2_3.5PN" else 28) / k_s / 1e12,
                            "peak": 2 * peak_dfma / 1e12, "unit": "TFLOP/s (FP64, DMMA)",
                            "frac": units_per_launch * (12 if cfg["approx"] == "TaylorF2_3.5PN" else 28) / k_s / peak_dfma}
                           if tensor_core else None),
                "hbm": {"achieved": alg_bytes / k_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / k_s / 1e9 / hbm_peak, "algorithmic_bytes": alg_bytes,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}

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

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload(cfg, args, W), "evals_per_s": value / nb,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "path": "bj_loglike (host pointers)" if world == 1 else
                            "ShardedEvaluator.evaluate: host X -> NCCL broadcast -> per-rank slices -> NCCL all-gather -> host"},
            "gpu_launches": 4 * args.steps, "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
            "sincos": args.sincos, "wall_s_timed_region": t_wall,
            "per_rank_ms": {"step_and_fused_kernel": per_rank_all}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
