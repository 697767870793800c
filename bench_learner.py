#!/usr/bin/env python
"""Secondary benchmark: end-to-end training steps of the batched Energy learner — predictor
forward, one batched IPOfunc forward + backward, predictor backward, gradient all-reduce, optimiser step — at the
reference's batch of 24 days (Interior/intopt_energy_mlp.py:312) and at larger batches, eager and as ONE CUDA graph.
Synthetic features / prices of the experiment's shape (48 slots per day, 8 features per slot).  One process per GPU
(torchrun for N > 1): every rank sees the same global minibatch and solves its shard; the predictor gradients are
all-reduced over NCCL inside the timed step.  Prints one JSON line per batch size (rank 0).

    python bench_learner.py [--batches 24 256 1024] [--steps 50]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_learner.py --gpus N"""
import argparse
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from neuripsintopt_b200.data import icon
from neuripsintopt_b200.energy_learner import IntoptEnergyLearner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[24, 256, 1024])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(0)
    for B in args.batches:
        X = torch.tensor(rng.standard_normal((B, 48, 8)), dtype=torch.float32, device=dev)
        y = torch.tensor(rng.uniform(20.0, 120.0, (B, 48)), dtype=torch.float32, device=dev)
        out = dict(metric="energy_learner_training_steps_per_s", batch_days=B, n_gpus=world)
        for mode in ("eager", "graph"):
            torch.manual_seed(0)
            learner = IntoptEnergyLearner(icon.load1_day01(), input_size=8, thr=0.1, damping=1e-6, lr=1e-2, device=dev)
            with torch.no_grad():  # start from positive prices, like a pre-trained predictor would give
                learner.model.net.weight.fill_(0.0)
                learner.model.net.bias.fill_(60.0)
            if mode == "eager":
                step = lambda: learner.training_step(X, y, rank=rank, world=world)  # noqa: E731
            else:
                g = learner.compile_step(B, 8, rank=rank, world=world)
                step = lambda: float(g(X, y))  # noqa: E731  (the loss read back: one device->host sync per step)
            losses = [step() for _ in range(args.warmup)]
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                losses.append(step())
            e1.record()
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            out[mode] = dict(ms_per_step=ms, steps_per_s=1e3 / ms, days_per_s=B * 1e3 / ms, first_loss=losses[0],
                             last_loss=losses[-1])
        if rank == 0:
            print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
