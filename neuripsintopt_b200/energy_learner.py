"""Batched Energy-scheduling learner: the reference's `intopt_energy.fit` loop
(Interior/intopt_energy_mlp.py:355-500) with the 24-instance Python loop replaced by ONE batched
`IPOfunc` call, sharded over ranks.

Reference step, per minibatch (:407-453): for each of the 24 days j: `c_pred = F @ model(X_j)`,
`x = IPOfunc(A,b,G,h,...)(c_pred)`, `loss = (x*c_true).mean()`, `loss.backward()`; then `optimizer.step()`.
Here: `c_pred = model(X).view(B,48) @ F^T` (B days at once), `x = IPOfunc(...)(c_pred)`,
`loss = (x*c_true).mean(dim=1).sum()`, one backward, gradient all-reduce, `optimizer.step()` — the same
accumulated gradient.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn, optim

from . import parallel
from .data import icon


class MultilayerRegression(nn.Module):
    """Same architecture as the reference predictor (Interior/intopt_energy_mlp.py:34-52): linear layers with
    Dropout in between and no non-linearity (the ReLUs are commented out upstream)."""

    def __init__(self, input_size, hidden_size, target_size=1, num_layers=1):
        super().__init__()
        if num_layers > 1:
            layers = [nn.Linear(input_size, hidden_size), nn.Dropout()]
            for _ in range(num_layers - 2):
                layers += [nn.Linear(hidden_size, hidden_size), nn.Dropout()]
            layers.append(nn.Linear(hidden_size, target_size))
            self.net = nn.Sequential(*layers)
        else:
            self.net = nn.Linear(input_size, target_size)

    def forward(self, x):
        return self.net(x)


class IntoptEnergyLearner:
    """`intopt_energy` (Interior/intopt_energy_mlp.py:310-500) with a batched, data-parallel training step.

    `lp_layer` maps a (B, n) cost batch to (B, n) solutions; by default it is the B200 `IPOfunc` built from the
    ICON instance `param` with the experiment's options (experiments/EnergyScheduling/exp_run.py:61-63)."""

    def __init__(self, param, input_size, hidden_size=1, num_layers=1, n_items=48, thr=0.1, damping=1e-6,
                 max_iter=None, optimizer=optim.Adam, device="cuda", lp_layer=None, **hyperparams):
        A, b, G, h, F = icon.make_matrix_intopt(**param)
        self.n_items = n_items
        self.device = torch.device(device)
        self.F = torch.as_tensor(F.toarray(), dtype=torch.float32, device=self.device)  # (n, 48)
        self.model = MultilayerRegression(input_size, hidden_size, 1, num_layers).to(self.device)
        self.optimizer = optimizer(self.model.parameters(), **hyperparams)
        if lp_layer is None:
            from .ip_model_whole import IPOfunc
            tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)  # noqa: E731
            lp_layer = IPOfunc(tt(A), tt(b), tt(G), tt(h), pc=True, max_iter=max_iter, bounds=[(0.0, None)],
                               thr=thr, method=1, damping=damping, check_finite=False)
        self.lp_layer = lp_layer

    def costs(self, X_days):
        """X_days: (B, n_items, features) -> predicted per-variable costs (B, n)."""
        B = X_days.shape[0]
        price = self.model(X_days.reshape(B * self.n_items, -1)).view(B, self.n_items)
        return price @ self.F.t()

    def training_step(self, X_days, y_days, rank=0, world=1, group=None):
        """One minibatch of whole days.  Every rank passes the SAME global batch and solves its own shard."""
        self.model.train()
        self.optimizer.zero_grad()
        X_loc = parallel.shard(torch.as_tensor(X_days, dtype=torch.float32, device=self.device), rank, world)
        y_loc = parallel.shard(torch.as_tensor(y_days, dtype=torch.float32, device=self.device), rank, world)
        c_true = y_loc @ self.F.t()
        c_pred = self.costs(X_loc)
        x = self.lp_layer(c_pred)
        loss = (x * c_true).mean(dim=1).sum()  # sum over days of the reference's per-day loss (:426)
        loss.backward()
        parallel.allreduce_gradients(self.model.parameters(), group=group)
        self.optimizer.step()
        return float(loss.detach())


def days(X, y, n_items=48):
    """Flat (T, features), (T,) series -> whole days (T // n_items, n_items, ...), like the reference's slicing
    `n_start = (batchsize*shuffled_batches[i] + j)*n_items` (:411-414)."""
    T = (len(y) // n_items) * n_items
    return np.asarray(X[:T]).reshape(-1, n_items, X.shape[1]), np.asarray(y[:T]).reshape(-1, n_items)
