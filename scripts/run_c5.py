#!/usr/bin/env python
"""
BASELINE config 5: end-to-end ensemble-sampler run on the 128 s BNS injection with the batched GPU log_like.

    python scripts/run_c5.py --walkers 4096 --steps 20                       # one GPU
    torchrun --nproc-per-node N scripts/run_c5.py --walkers 4096 --steps 20  # N GPUs (walkers sharded)

Rank 0 owns the sampler (stretch-move ensemble, the update bajes' emcee wrapper drives) and reaches the likelihood
ONLY through ``BatchedPool.map(posterior.log_post, points)``, i.e. the unchanged sampler contract; the other ranks sit
in ``ShardedEvaluator.serve()``.  Prints one JSON line: wall-clock per sampler step, likelihood evaluations per
second end to end (host -> device -> host every call), acceptance, logL trace.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--sincos", default="mufu")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from bajes_b200 import GWLikelihood, Posterior, synthetic as syn
    from bajes_b200.dist import ShardedEvaluator, ShardedLikelihood
    from bajes_b200.ensemble import BoxPrior, EnsembleSampler
    from bajes_b200.pool import BatchedPool

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = syn.CONFIGS[args.config]
    cl = syn.product_classes()
    like = GWLikelihood(*syn.build_network(cfg, cl), cfg["srate"], cfg["seglen"], cfg["approx"], device=local,
                        sincos=args.sincos)
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    inj = syn.INJECTIONS[cfg["source"]]
    # prior box: a few posterior widths around the injection (SURVEY.md section 8d "narrow")
    p0, _ = syn.draw_walkers(cfg, args.walkers, 123, "narrow")
    lo, hi = p0.min(axis=0), p0.max(axis=0)
    pad = 0.5 * (hi - lo)
    bounds = np.stack([lo - pad, hi + pad], axis=1)
    bounds[names.index("q"), 0] = max(1.0, bounds[names.index("q"), 0])
    bounds[names.index("cos_iota")] = np.clip(bounds[names.index("cos_iota")], -1, 1)
    bounds[names.index("lambda1"), 0] = max(0.0, bounds[names.index("lambda1"), 0])
    bounds[names.index("lambda2"), 0] = max(0.0, bounds[names.index("lambda2"), 0])
    prior = BoxPrior(names, bounds, const)

    pmap = like.param_map(names, const)

    def eval_device(x_dev, out_dev):
        like.engine.loglike_device(x_dev.data_ptr(), x_dev.shape[0], len(names), pmap, out_dev.data_ptr(), 0,
                                   torch.cuda.current_stream(dev).cuda_stream)

    ev = ShardedEvaluator(lambda Xs: like.log_like_batch(Xs, names, const), ndim=len(names), device=dev,
                          local_eval_device=eval_device if world > 1 else None)
    if rank != 0:
        ev.serve()
        dist.destroy_process_group()
        return

    post = Posterior(ShardedLikelihood(like, ev), prior)
    pool = BatchedPool(post, processes=world)
    sampler = EnsembleSampler(post, args.walkers, pool, seed=7)
    t0 = time.perf_counter()
    sampler.initialize(p0)
    t_init = time.perf_counter() - t0
    logl_inj = like.log_like(syn.injection_params(cfg))
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    trace = sampler.run(args.steps)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    ev.shutdown()

    nb = like.engine.n_bins
    n_eval = args.walkers * args.steps            # two half-ensembles per step
    # spot check: the sampler's stored log-posterior equals a fresh batched evaluation + prior
    idx = np.arange(0, args.walkers, max(1, args.walkers // 64))
    fresh = like.log_like_batch(sampler.pos[idx], names, const) + prior.log_prior(None)
    consistent = bool(np.array_equal(fresh, sampler.logp[idx]))
    line = {"config": args.config, "n_gpus": world, "walkers": args.walkers, "steps": args.steps,
            "sampler": "stretch-move ensemble via BatchedPool.map(posterior.log_post)",
            "s_per_step": wall / args.steps, "evals_per_s": n_eval / wall, "walker_bins_per_s": n_eval * nb / wall,
            "batched_calls": pool.n_batched_calls, "init_s": t_init, "acceptance": sampler.acceptance,
            "logL_injection": logl_inj, "logpost_max_first_last": [float(trace[0, 0]), float(trace[-1, 0])],
            "logpost_median_first_last": [float(trace[0, 1]), float(trace[-1, 1])],
            "stored_logp_bit_identical_to_fresh_eval": consistent, "sincos": args.sincos}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
