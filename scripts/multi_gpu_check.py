#!/usr/bin/env python
"""
Multi-GPU correctness check (run under torchrun, one rank per GPU; tests/test_gpu_multi.py spawns it):

  * walker sharding through ``ShardedEvaluator`` -- one scatter out, logL stored by every rank's epilogue into the
    sampler rank's HBM (peer memory) -- returns, for EVERY rank's slice, the bits a single GPU produces;
  * the same with the gather fall-back (``peer=False``) and across a capacity growth;
  * the frequency-axis split (``FrequencySplitEvaluator``: one all-reduce of the partial sums) agrees with the
    single-GPU values to rounding.

Prints one JSON line on rank 0 and exits non-zero on any mismatch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from bajes_b200 import GWLikelihood, synthetic as syn
    from bajes_b200.dist import FrequencySplitEvaluator, ShardedEvaluator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    cfg = syn.CONFIGS["C2_small"]
    like = GWLikelihood(*syn.build_network(cfg, syn.product_classes()), cfg["srate"], cfg["seglen"], cfg["approx"], device=local)
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    pmap = like.param_map(names, const)

    def eval_dev(xs, outs):
        like.engine.loglike_device(xs.data_ptr(), xs.shape[0], len(names), pmap, outs.data_ptr(), 0,
                                   torch.cuda.current_stream(dev).cuda_stream)

    report = {"world": world}
    ok = True
    for peer in (None, False):
        ev = ShardedEvaluator(lambda Xs: like.log_like_batch(Xs, names, const), ndim=len(names), device=dev,
                              local_eval_device=eval_dev, capacity=512, peer=peer)
        tag = "peer" if ev.peer is not None else "gather"
        if rank == 0:
            res = []
            for n, seed in ((1, 1), (257, 2), (512, 3), (3001, 4), (64, 5)):     # 3001 > capacity: collective growth
                Xa, _ = syn.draw_walkers(cfg, n - n // 2, seed, "narrow")
                Xb, _ = syn.draw_walkers(cfg, n // 2, seed + 100, "wide")
                X = np.concatenate([Xa, Xb]) if n > 1 else Xa
                got = ev.evaluate(X)
                alone = like.log_like_batch(X, names, const)
                res.append(bool(np.array_equal(got, alone)))
            ev.shutdown()
            report[tag] = res
            ok = ok and all(res)
        else:
            ev.serve()
        dist.barrier()
    # frequency split
    fs = FrequencySplitEvaluator(like, names, const)
    if rank == 0:
        X, _ = syn.draw_walkers(cfg, 37, 9, "narrow")
        got = fs.evaluate(X)
        alone = like.log_like_batch(X, names, const)
        fs.shutdown()
        err = np.abs(got - alone) / (1e-6 * np.abs(alone) + 1e-4)
        report["freq_split_max_err_over_tol"] = float(err.max())
        ok = ok and float(err.max()) < 1e-3
    else:
        fs.serve()
    dist.barrier()
    if rank == 0:
        report["ok"] = ok
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
