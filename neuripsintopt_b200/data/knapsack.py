"""Building-knapsack data: the two pickled frames, the experiment's train/test split and the
per-minibatch arrays the reference's `MyCustomDataset` serves.

  * `load_frames` / `split` follow experiments/Building Knapsack/ModelRun.py:14-27: 310 of the 372 properties
    (`sample(n=310, random_state=10)`) train, the rest test, both sorted by serial number; the economic series
    of the same serials, sorted by (Sl, Lag descending);
  * `arrays` follows optimizer.py:81-102: 7 numeric property features (columns 6..12), the ZIP-code category
    code, sale price (column 13, the predicted quantity), construction cost (column 14, the knapsack weight) and
    the `lookback` x 19 economic window of each property;
  * `minibatches` cuts them into knapsack instances of `items` properties (batch_size=31, ModelRun.py:58), each
    one LP for `IntoptKnapsackLearner` (costs (K, items) -> `A` of shape (K, 1, items)).
"""
from __future__ import annotations

import numpy as np


def load_frames(property_pkl, economic_pkl):
    import pandas as pd
    return pd.read_pickle(property_pkl), pd.read_pickle(economic_pkl)


def split(prop, econ, n_train=310, random_state=10):
    """-> (train_prop, train_econ, test_prop, test_econ)."""
    train_prop = prop.sample(n=n_train, random_state=random_state)
    test_prop = prop.loc[~prop.index.isin(train_prop.index)]
    out = []
    for part in (train_prop, test_prop):
        part = part.sort_values(["Sl"], ascending=[True])
        e = econ[econ.Sl.isin(part.Sl.unique().tolist())].sort_values(["Sl", "Lag"], ascending=[True, False])
        out += [part, e]
    return tuple(out)


def arrays(prop, econ, lookback=5, upstream_indexing=False):
    """-> dict(x_features (P,7), x_cat (P,), x_ts (P,lookback,19), y (P,), cost (P,)).

    `upstream_indexing=True` reproduces what the reference's dataset actually serves: it allocates one window
    per economic ROW, fills only every `lookback`-th of them (optimizer.py:91-96) and then indexes that array by
    PROPERTY (:101), so property i sees rows i..i+lookback of the series when i is a multiple of `lookback` and
    zeros otherwise.  The default pairs every property with its own window, which is what the layout intends."""
    ts = econ.iloc[:, 2:].values.astype(np.float32)
    if upstream_indexing:
        win = np.zeros((len(ts), lookback, ts.shape[1]), dtype=np.float32)
        for i in range(0, len(ts), lookback):
            win[i] = ts[i:i + lookback]
        x_ts = win[:len(prop)]
    else:
        x_ts = ts.reshape(-1, lookback, ts.shape[1])
    return {
        "x_features": prop.iloc[:, 6:13].values.astype(np.float32),
        "x_cat": prop.iloc[:, 5].cat.codes.values.astype(np.int64),
        "x_ts": x_ts,
        "y": prop.iloc[:, 13].values.astype(np.float32),
        "cost": prop.iloc[:, 14].values.astype(np.float32),
    }


def minibatches(arr, items=31):
    """Whole minibatches only: every array reshaped to (K, items, ...)."""
    K = len(arr["y"]) // items
    return {k: v[:K * items].reshape((K, items) + v.shape[1:]) for k, v in arr.items()}
