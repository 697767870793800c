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
        self.param = param
        self.n_items = n_items
        self.device = torch.device(device)
        self.F = torch.as_tensor(F.toarray(), dtype=torch.float32, device=self.device)  # (n, 48)
        self.model = MultilayerRegression(input_size, hidden_size, 1, num_layers).to(self.device)
        parallel.broadcast_parameters(self.model.parameters())
        if optimizer is optim.Adam and self.device.type == "cuda":
            hyperparams.setdefault("capturable", True)  # state on the device: the step can live inside a CUDA graph
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
        if X_loc.shape[0]:
            c_true = y_loc @ self.F.t()
            c_pred = self.costs(X_loc)
            x = self.lp_layer(c_pred)
            loss = (x * c_true).mean(dim=1).sum()  # sum over days of the reference's per-day loss (:426)
        else:  # more ranks than days
            loss = parallel.zero_loss(self.model.parameters())
        loss.backward()
        parallel.allreduce_gradients(self.model.parameters(), group=group)
        self.optimizer.step()
        return float(loss.detach())

    # ---- the same step as ONE CUDA graph --------------------------------------------------------------------
    def compile_step(self, batch_days, n_features, rank=0, world=1, group=None):
        """Capture one training step for a FIXED minibatch size into a CUDA graph and return `step(X_days, y_days)`.

        A step at the reference's batch of 24 days is launch- and Python-bound (a dozen small predictor / optimiser
        kernels around the two solver launches): captured once, every later step is one graph launch — predictor
        forward, the batched `IPOfunc` forward + backward (the library only enqueues kernels on the capturing stream),
        predictor backward and the optimiser step.  With several ranks the step is two graphs (forward + backward, then
        the optimiser) with the flat gradient all-reduce issued eagerly in between, so that no collective lives inside
        a capture.  The optimiser must be
        capturable (`torch.optim.Adam(..., capturable=True)` — done by `__init__` for Adam on CUDA).  The returned
        callable copies its inputs into the graph's static buffers, replays, and returns the loss (one device->host
        read).  The eager `training_step` stays the reference for results: both produce the same numbers."""
        dev = self.device
        lo, hi = parallel.shard_bounds(batch_days, rank, world)
        Xs = torch.zeros(hi - lo, self.n_items, n_features, dtype=torch.float32, device=dev)
        ys = torch.zeros(hi - lo, self.n_items, dtype=torch.float32, device=dev)
        loss_buf = torch.zeros((), dtype=torch.float32, device=dev)

        distributed = world > 1

        def fwd_bwd():
            c_true = ys @ self.F.t()
            x = self.lp_layer(self.costs(Xs))
            loss = (x * c_true).mean(dim=1).sum()
            loss.backward()
            return loss

        def body():
            loss = fwd_bwd()
            parallel.allreduce_gradients(self.model.parameters(), group=group)
            self.optimizer.step()
            return loss

        self.model.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        params0 = [p.detach().clone() for p in self.model.parameters()]
        had_state = len(self.optimizer.state) > 0
        opt0 = self.optimizer.state_dict() if had_state else None
        with torch.cuda.stream(side):  # warm-up off the default stream (allocator, cold-start factor, NCCL, autograd)
            Xs.normal_()
            ys.uniform_(20.0, 120.0)
            for _ in range(3):
                self.optimizer.zero_grad(set_to_none=True)
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():  # the warm-up steps must not count: restore parameters and optimiser state
            for p, p0 in zip(self.model.parameters(), params0):
                p.copy_(p0)
            if had_state:
                self.optimizer.load_state_dict(opt0)
            else:
                # a fresh optimiser: keep the state tensors the warm-up created (a lazy initialisation INSIDE the capture
                # would be replayed, i.e. reset the moments at every step) and zero them in place
                for st in self.optimizer.state.values():
                    for v in st.values():
                        if isinstance(v, torch.Tensor):
                            v.zero_()
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        if not distributed:
            with torch.cuda.graph(graph):
                loss_buf.copy_(body().detach())
            self._graph = (graph,)  # keep alive
        else:
            with torch.cuda.graph(graph):
                loss_buf.copy_(fwd_bwd().detach())
            graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_opt, pool=graph.pool()):
                self.optimizer.step()
            self._graph = (graph, graph_opt)

        def step(X_days, y_days):
            Xs.copy_(torch.as_tensor(X_days, dtype=torch.float32)[lo:hi], non_blocking=True)
            ys.copy_(torch.as_tensor(y_days, dtype=torch.float32)[lo:hi], non_blocking=True)
            graph.replay()
            if distributed:
                parallel.allreduce_gradients(self.model.parameters(), group=group)
                self._graph[1].replay()
            return loss_buf

        return step

    # ---- the epoch / minibatch driver of the reference (Interior/intopt_energy_mlp.py:355-500) -------------
    def fit(self, X, y, X_validation=None, y_validation=None, epochs=1, batchsize=24, doScale=True, seed=None, shuffle=True,
            validate_every=7, validation_relax=True, rank=0, world=1, group=None, log=None):
        """`intopt_energy.fit` (:355-500): standardise the features (:370-372), then for every epoch visit the
        `n_batches = T // (batchsize * n_items)` minibatches of whole days in a shuffled order (:388-406); a minibatch is
        ONE batched training step here instead of the reference's loop over its `batchsize` days (:409-453: the order
        inside a minibatch only permutes a sum of gradients).  Every `validate_every` minibatches, and at the end of an
        epoch (:464), the regret on the validation days is recorded if validation data was given (:468-485; HiGHS
        instead of Gurobi, evaluate.EnergySchedule).  Returns a list of dicts, one per minibatch.

        `seed`: the reference shuffles with the global numpy RNG (:404); a seed makes the order reproducible
        (`shuffle=False` keeps the minibatches in storage order).  All ranks
        must use the same seed (they see the same global minibatch and take their own shard of it)."""
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if doScale:
            self.scaler = StandardScaler().fit(X)
            X = self.scaler.transform(X)
        else:
            self.scaler = None
        n_items = self.n_items
        Xd, yd = days(X, y, n_items)
        n_batches = X.shape[0] // (batchsize * n_items)  # :386
        order = list(range(n_batches))
        rng = np.random.RandomState(seed) if seed is not None else np.random
        history = []
        for e in range(epochs):
            if shuffle:
                rng.shuffle(order)  # :404
            for i in range(n_batches):
                lo = batchsize * order[i]
                loss = self.training_step(Xd[lo:lo + batchsize], yd[lo:lo + batchsize], rank=rank, world=world, group=group)
                rec = {"epoch": e, "batch": i, "first_day": lo, "loss": loss / batchsize}
                if X_validation is not None and ((i + 1) % validate_every == 0 or (i + 1) % n_batches == 0):
                    rec.update(self.validation_result(X_validation, y_validation, relax=validation_relax))
                history.append(rec)
                if log is not None:
                    log(rec)
        return history

    def predict(self, X, doScale=True):
        """Predicted prices per slot (:504-506, predict_func): the features go through the scaler fitted by `fit`."""
        X = np.asarray(X, dtype=np.float64)
        if doScale and getattr(self, "scaler", None) is not None:
            X = self.scaler.transform(X)
        self.model.eval()
        with torch.no_grad():
            out = self.model(torch.as_tensor(X, dtype=torch.float32, device=self.device)).squeeze(-1)
        self.model.train()
        return out.cpu().numpy().astype(np.float64)

    def validation_result(self, X_validation, y_validation, relax=True, doScale=True, max_days=None):
        """Regret and MSE on whole validation days (:500-502, validation_func): schedule with the predicted prices, pay the
        true ones, against the schedule that knew them (evaluate.EnergySchedule: the same program on HiGHS)."""
        from .evaluate import EnergySchedule
        if not hasattr(self, "_schedules"):
            self._schedules = {}
        if relax not in self._schedules:
            self._schedules[relax] = EnergySchedule(self.param, relax=relax)
        sched = self._schedules[relax]
        y_pred = self.predict(X_validation, doScale=doScale)
        y_true = np.asarray(y_validation, dtype=np.float64)
        nd = len(y_true) // self.n_items
        if max_days is not None:
            nd = min(nd, max_days)
        regret = [sched.regret(y_pred[d * self.n_items:(d + 1) * self.n_items], y_true[d * self.n_items:(d + 1) * self.n_items])
                  for d in range(nd)]
        T = nd * self.n_items
        return {"regret": float(np.mean(regret)), "mse": float(np.mean((y_pred[:T] - y_true[:T]) ** 2))}


class StandardScaler:
    """Zero mean / unit (population) variance per column, like sklearn.preprocessing.StandardScaler, which the reference
    fits on the training features (:370-372); constant columns are left unscaled."""

    def fit(self, X):
        X = np.asarray(X, dtype=np.float64)
        self.mean_ = X.mean(axis=0)
        sd = X.std(axis=0)
        self.scale_ = np.where(sd == 0.0, 1.0, sd)
        return self

    def transform(self, X):
        return (np.asarray(X, dtype=np.float64) - self.mean_) / self.scale_


def days(X, y, n_items=48):
    """Flat (T, features), (T,) series -> whole days (T // n_items, n_items, ...), like the reference's slicing
    `n_start = (batchsize*shuffled_batches[i] + j)*n_items` (:411-414)."""
    T = (len(y) // n_items) * n_items
    return np.asarray(X[:T]).reshape(-1, n_items, X.shape[1]), np.asarray(y[:T]).reshape(-1, n_items)
