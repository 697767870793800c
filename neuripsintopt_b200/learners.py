"""Batched counterparts of the reference's other two IntOpt training loops (SURVEY.md §8 f-1 widened to the
Twitter shortest-path and Building-knapsack drivers).  Like `energy_learner.IntoptEnergyLearner`, each replaces
the per-instance Python loop around `IPOfunc` by ONE batched call, sharded over ranks; the predictor is whatever
`nn.Module` the caller passes (the reference's are `make_fc`, shortespath/shortespath.py:128-137, and the LSTM
`PriceNet`, experiments/Building Knapsack/optimizer.py:48-78).

* shortest path (`intopt.fit`, shortespath/shortespath.py:377-449): the reference solves one LP per training
  instance (source, destination) with the SAME predicted edge costs and steps the optimiser after every instance;
  here a batch of instances goes through one call — `b` of shape (B, n_eq), one shared `c` — and the gradients of
  the instances add up before one optimiser step.
* knapsack (`Intopt.fit`, experiments/Building Knapsack/optimizer.py:359-412): every minibatch of items is its own
  LP whose equality row is the minibatch's cost vector (`IPOfunc(G, h, A, b, ...)` with the cost row in the FIRST,
  i.e. equality, slot — SURVEY §3.5); here K minibatches are solved at once through the per-instance-values path
  (`A` of shape (K, 1, items)).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn, optim

from . import parallel


def make_fc(num_layers, num_features, num_targets=1, activation_fn=nn.ReLU, intermediate_size=50):
    """The reference's edge-cost predictor (shortespath/shortespath.py:128-137): ReLU-terminated MLP."""
    layers = [nn.Linear(num_features, intermediate_size), activation_fn()]
    for _ in range(num_layers - 2):
        layers += [nn.Linear(intermediate_size, intermediate_size), activation_fn()]
    layers += [nn.Linear(intermediate_size, num_targets), nn.ReLU()]
    return nn.Sequential(*layers)


class IntoptShortestPathLearner:
    """`intopt` of shortespath/shortespath.py with a batched training step.

    A: node-arc incidence matrix after the reference's redundancy removal (n_eq x n_edges); `removed_rows` are the
    row indices that removal dropped from the full matrix, so that right-hand sides can be built from (source,
    destination) pairs exactly like shortespath.py:407-413."""

    def __init__(self, A_rr, n_nodes, removed_rows, net, thr=0.1, max_iter=100, damping=1e-2, optimizer=optim.Adam,
                 device="cuda", **hyperparams):
        self.device = torch.device(device)
        self.A = torch.as_tensor(np.asarray(A_rr), dtype=torch.float32)
        self.n_nodes = int(n_nodes)
        self.keep = np.setdiff1d(np.arange(self.n_nodes), np.asarray(removed_rows))
        self.net = net.to(self.device)
        parallel.broadcast_parameters(self.net.parameters())
        self.optimizer = optimizer(self.net.parameters(), **hyperparams)
        self.opts = dict(bounds=[(0.0, None)], max_iter=max_iter, thr=thr, damping=damping, method=1,
                         check_finite=False)

    def rhs(self, pairs):
        """(source, destination) pairs -> (B, n_eq) right-hand sides (+1 at the source, -1 at the destination)."""
        b = np.zeros((len(pairs), self.n_nodes), dtype=np.float32)
        for i, (s, d) in enumerate(pairs):
            b[i, s] = 1.0
            b[i, d] = -1.0
        return b[:, self.keep]

    def solve(self, c_pred, pairs):
        """x (B, n_edges) for the shared predicted costs `c_pred` (n_edges,) and a batch of instances."""
        from .ip_model_whole import IPOfunc
        f = IPOfunc(self.A, torch.as_tensor(self.rhs(pairs)), torch.Tensor(), torch.Tensor(), **self.opts)
        return f(c_pred)

    def training_step(self, X, y, pairs, rank=0, world=1, group=None):
        """One optimiser step on a batch of instances; every rank passes the same `pairs` and solves its shard."""
        self.net.train()
        self.optimizer.zero_grad()
        X = torch.as_tensor(X, dtype=torch.float32, device=self.device)
        y = torch.as_tensor(y, dtype=torch.float32, device=self.device)
        lo, hi = parallel.shard_bounds(len(pairs), rank, world)
        c_pred = self.net(X).squeeze(-1)
        if hi > lo:
            x = self.solve(c_pred, list(pairs)[lo:hi])
            loss = (y * x).mean(dim=1).sum()  # sum over instances of the reference's per-instance loss (:426)
        else:  # more ranks than instances: nothing to solve here, but the collective below needs every rank
            loss = parallel.zero_loss(self.net.parameters())
        loss.backward()
        parallel.allreduce_gradients(self.net.parameters(), group=group)
        self.optimizer.step()
        return float(loss.detach())


class IntoptKnapsackLearner:
    """`Intopt` of experiments/Building Knapsack/optimizer.py with K minibatches per call.

    `predict(batch)` -> predicted prices (K, items) is supplied by the caller (upstream: the LSTM `PriceNet`)."""

    def __init__(self, model, budget=20.0, thr=0.1, damping=0.1, max_iter=1000, optimizer=optim.Adam, device="cuda",
                 **hyperparams):
        self.device = torch.device(device)
        self.model = model.to(self.device)
        parallel.broadcast_parameters(self.model.parameters())
        self.optimizer = optimizer(self.model.parameters(), **hyperparams)
        self.budget = float(budget)
        self.opts = dict(bounds=[(0.0, 1.0)], max_iter=max_iter, thr=thr, damping=damping, check_finite=False)

    def solve(self, prices, costs):
        """x (K, items): each row solves  max price.x  s.t.  cost.x = items * budget, 0 <= x <= 1  (the cost row sits
        in the equality slot, as upstream's positional call does, optimizer.py:404)."""
        from .ip_model_whole import IPOfunc
        K, items = costs.shape
        A = torch.as_tensor(costs, dtype=torch.float32).reshape(K, 1, items)
        h = torch.tensor([items * self.budget], dtype=torch.float32)
        return IPOfunc(A, h, torch.Tensor(), torch.Tensor(), **self.opts)(-prices)

    def training_step(self, features, y, costs, rank=0, world=1, group=None):
        """features: whatever `model` consumes, leading dimension K (minibatches); y, costs: (K, items)."""
        self.model.train()
        self.optimizer.zero_grad()
        y = torch.as_tensor(y, dtype=torch.float32, device=self.device)
        lo, hi = parallel.shard_bounds(y.shape[0], rank, world)
        prices = self.model(features)[lo:hi]
        if hi > lo:
            x = self.solve(prices, np.asarray(costs)[lo:hi])
            loss = -(x * y[lo:hi]).mean(dim=1).sum()  # (:406) per minibatch, summed
        else:
            loss = parallel.zero_loss(self.model.parameters())
        loss.backward()
        parallel.allreduce_gradients(self.model.parameters(), group=group)
        self.optimizer.step()
        return float(loss.detach())
