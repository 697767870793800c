#!/usr/bin/env python
"""Secondary benchmark: end-to-end training steps of the batched Energy learner — predictor
forward, one batched IPOfunc forward + backward, predictor backward, optimiser step — at the reference's batch of
24 days (Interior/intopt_energy_mlp.py:312) and at larger batches.  Synthetic features / prices of the experiment's
shape (48 slots per day, 8 features per slot).  Prints one JSON line per batch size.

    python bench_learner.py [--batches 24 256 2368] [--steps 20]"""
import argparse
import json

import numpy as np
import torch

from neuripsintopt_b200.data import icon
from neuripsintopt_b200.energy_learner import IntoptEnergyLearner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[24, 256, 2368])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    for B in args.batches:
        torch.manual_seed(0)
        learner = IntoptEnergyLearner(icon.load1_day01(), input_size=8, thr=0.1, damping=1e-6, lr=1e-3)
        with torch.no_grad():  # start from positive prices, like a pre-trained predictor would give
            learner.model.net.weight.fill_(0.0)
            learner.model.net.bias.fill_(60.0)
        X = torch.tensor(rng.standard_normal((B, 48, 8)), dtype=torch.float32, device="cuda")
        y = torch.tensor(rng.uniform(20.0, 120.0, (B, 48)), dtype=torch.float32, device="cuda")
        losses = [learner.training_step(X, y) for _ in range(args.warmup)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            losses.append(learner.training_step(X, y))  # returns the loss: one device->host read per step
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps(dict(metric="energy_learner_training_steps_per_s", batch_days=B, steps_per_s=1e3 / ms,
                              days_per_s=B * 1e3 / ms, ms_per_step=ms, first_loss=losses[0], last_loss=losses[-1])),
              flush=True)


if __name__ == "__main__":
    main()
