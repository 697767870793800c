"""Energy price data of the ICON scheduling experiment: the (X, y) loader the reference drivers call
(`get_energy`, get_energy.py:16-45) and the experiment's train / validation / test split
(experiments/EnergyScheduling/exp_run.py:29-39).

Upstream reads `prices2013.dat`, which the repository does not ship; it ships the loader's output as `Data.npz`
(keys X_1gtrain, y_train, X_1gtest, y_test) — except that the stored prices are NOT clamped (y_train min -335,
SURVEY C.11) while `get_energy` clamps negative prices to zero (get_energy.py:27-28).  Both sources are accepted
here and both come out clamped, i.e. as `get_energy` returns them.
"""
from __future__ import annotations

import numpy as np

GROUP = 48  # half-hour slots per day (get_energy.py:21)


def get_energy(fname, trainTestRatio=0.70):
    """(X_1gtrain, y_train, X_1gtest, y_test); X has the group (day) id as its first column (get_energy.py:23)."""
    if str(fname).endswith(".npz"):
        z = np.load(fname)
        X_tr, y_tr, X_te, y_te = (np.array(z[k], dtype=np.float64) for k in ("X_1gtrain", "y_train", "X_1gtest", "y_test"))
        return X_tr, np.maximum(y_tr, 0.0), X_te, np.maximum(y_te, 0.0)
    df = get_energy_pandas(fname)
    length = df["groupID"].nunique()
    X1g = df.loc[:, df.columns != "SMPEP2"].values.astype(np.float64)
    y = np.maximum(df.loc[:, "SMPEP2"].values.astype(np.float64), 0.0)  # :27-28
    train_len = int(trainTestRatio * length)  # ordered split per complete day (:32-38)
    cut = GROUP * train_len
    return X1g[:cut], y[:cut], X1g[cut:], y[cut:]


def get_energy_pandas(fname):
    """The raw table as the reference prepares it (get_energy.py:62-95): string / leaking / incomplete columns dropped,
    zero CO2 intensities interpolated, days with missing values removed, a day id in front."""
    import pandas as pd

    df = pd.read_csv(fname, sep=r"\s+", quotechar='"')
    df = df.drop(columns=["#DateTime", "Holiday", "ActualWindProduction", "SystemLoadEP2", "ORKTemperature", "ORKWindspeed"])
    df.loc[df["CO2Intensity"] == 0, "CO2Intensity"] = np.nan
    df["CO2Intensity"] = df["CO2Intensity"].interpolate()
    bad = [i for i in range(0, len(df), GROUP) if df.iloc[i:i + GROUP].isnull().any(axis=1).any()]
    df = df.drop(index=[df.index[j] for i in bad for j in range(i, min(i + GROUP, len(df)))])
    df = df.drop(columns=["Day", "Year", "PeriodOfDay"])
    df.insert(0, "groupID", [g for g in range(len(df) // GROUP) for _ in range(GROUP)][:len(df)])
    return df


def experiment_split(X_1gtrain, y_train, X_1gtest, y_test, n_validation=2880):
    """exp_run.py:29-39: the first 60 test days become the validation set and the day-id column is dropped.
    Returns (X_train, y_train, X_validation, y_validation, X_test, y_test)."""
    return (X_1gtrain[:, 1:], y_train, X_1gtest[:n_validation, 1:], y_test[:n_validation],
            X_1gtest[n_validation:, 1:], y_test[n_validation:])


def synthetic_days(n_days, n_features=8, seed=0):
    """Synthetic (X, y) of the experiment's shape: prices U[20, 120], standard-normal features."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n_days * GROUP, n_features)), rng.uniform(20.0, 120.0, n_days * GROUP)
