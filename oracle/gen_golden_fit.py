"""TEST INFRASTRUCTURE — generates tests/golden/energy_fit.npz: the first optimiser steps of the reference's Energy
training loop (Interior/intopt_energy_mlp.py:355-453), run here with the UNTOUCHED reference IPOfunc
(oracle/ref_shim.py), its own `make_matrix_intopt` / `data_reading` / `MultilayerRegression` (lifted from the driver
file without importing it: the module imports gurobipy) on real days of Data.npz.

The loop body below is the reference's, statement for statement (:407-453), minus logging / time limits; the
minibatch order is fixed (the reference shuffles, which only permutes which days come first).

    python oracle/gen_golden_fit.py     (build container only: needs /root/reference; ~1 min)
"""
import ast
import os
import sys

import numpy as np
import torch
from torch import nn, optim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

STEPS, BATCH, N_ITEMS = 3, 24, 48


def lift(relpath, names):
    path = os.path.join(ref_shim.REF_ROOT, relpath)
    tree = ast.parse(open(path).read())
    ns = {"np": np, "torch": torch, "nn": nn}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def main():
    W = ref_shim.load("whole")
    make_matrix_intopt, MultilayerRegression = lift("Interior/intopt_energy_mlp.py", ["make_matrix_intopt", "MultilayerRegression"])
    (data_reading,) = lift("EnergyCost/ICON.py", ["data_reading"])
    param = data_reading(os.path.join(ref_shim.REF_ROOT, "EnergyCost", "load1", "day01.txt"))
    z = np.load(os.path.join(ref_shim.REF_ROOT, "Data.npz"))
    T = STEPS * BATCH * N_ITEMS
    X_raw = np.array(z["X_1gtrain"][:T, 1:], dtype=np.float64)  # exp_run.py:36 drops the day-id column
    y = np.maximum(np.array(z["y_train"][:T], dtype=np.float64), 0.0)  # get_energy.py:27-28
    from sklearn import preprocessing
    scaler = preprocessing.StandardScaler().fit(X_raw)  # :370-372
    X = scaler.transform(X_raw)

    torch.manual_seed(0)
    model = MultilayerRegression(input_size=X.shape[1], hidden_size=1, target_size=1, num_layers=1)
    w0 = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    optimizer = optim.Adam(model.parameters(), lr=0.7)  # exp_run.py:61-63
    A, b, G, h, F = make_matrix_intopt(**param)
    losses = np.zeros((STEPS, BATCH))
    weights, biases, grads_w, grads_b = [], [], [], []
    for i in range(STEPS):
        optimizer.zero_grad()
        for j in range(BATCH):
            n_start = (BATCH * i + j) * N_ITEMS  # :411
            n_stop = n_start + N_ITEMS
            X_tensor = torch.tensor(X[n_start:n_stop, :], dtype=torch.float)
            c_true = torch.mm(F, torch.tensor(y[n_start:n_stop], dtype=torch.float).unsqueeze(1)).squeeze()
            c_pred = torch.mm(F, model(X_tensor)).squeeze()
            x = W.IPOfunc(A, b, G, h, pc=True, max_iter=None, bounds=[(0., None)], smoothing=False, thr=0.1, method=1,
                          mu0=None, damping=1e-6)(c_pred)  # :422-425 with the experiment's options
            loss = (x * c_true).mean()
            loss.backward()
            losses[i, j] = float(loss)
        grads_w.append(model.state_dict() and [p.grad.detach().numpy().copy() for p in model.parameters()][0])
        grads_b.append([p.grad.detach().numpy().copy() for p in model.parameters()][1])
        optimizer.step()  # :453
        ps = [p.detach().numpy().copy() for p in model.parameters()]
        weights.append(ps[0])
        biases.append(ps[1])
        print("step %d: mean loss %.6f  weight %s bias %s" % (i, losses[i].mean(), ps[0].ravel()[:3], ps[1]), flush=True)
    out = os.path.join(ROOT, "tests", "golden", "energy_fit.npz")
    np.savez_compressed(out, X_raw=X_raw.astype(np.float32), y=y.astype(np.float32), w0=w0[list(w0)[0]], b0=w0[list(w0)[1]],
                        losses=losses, weights=np.stack(weights), biases=np.stack(biases), grads_w=np.stack(grads_w),
                        grads_b=np.stack(grads_b), lr=0.7, thr=0.1, damping=1e-6, steps=STEPS, batch=BATCH)
    print("wrote", out)


if __name__ == "__main__":
    main()
