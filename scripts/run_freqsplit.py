#!/usr/bin/env python
"""Frequency-axis split on N GPUs (torchrun): a small batch evaluated with every rank taking a share of the frequency
chunks + one NCCL all-reduce of the per-walker partial sums, compared with the single-GPU result."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from bajes_b200 import GWLikelihood, synthetic as syn
from bajes_b200.dist import FrequencySplitEvaluator

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = syn.CONFIGS["C2"]
like = GWLikelihood(*syn.build_network(cfg, syn.product_classes()), cfg["srate"], cfg["seglen"], cfg["approx"], device=local)
names, const = syn.WALKER_NAMES, syn.constants(cfg)
X, _ = syn.draw_walkers(cfg, W, 5, "narrow")
ev = FrequencySplitEvaluator(like, names, const)
if rank != 0:
    ev.serve()
    dist.destroy_process_group()
    sys.exit(0)
single = like.log_like_batch(X, names, const)
for _ in range(3):
    got = ev.evaluate(X)
torch.cuda.synchronize(dev)
t0 = time.perf_counter()
K = 20
for _ in range(K):
    got = ev.evaluate(X)
t_split = (time.perf_counter() - t0) / K
t0 = time.perf_counter()
for _ in range(K):
    like.log_like_batch(X, names, const)
t_single = (time.perf_counter() - t0) / K
ev.shutdown()
print(json.dumps({"n_gpus": world, "walkers": W, "n_chunks": like.engine.num_chunks(), "ms_split_e2e": 1e3 * t_split,
                  "ms_single_gpu_e2e": 1e3 * t_single, "max_rel_diff_vs_single_gpu": float(np.max(np.abs(got - single) / (np.abs(single) + 1.0)))}))
if world > 1:
    dist.destroy_process_group()
