"""Loader for the fixtures written by oracle/gen_golden.py (outputs of the untouched reference)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dense(z, prefix):
    shape = tuple(int(v) for v in z[prefix + "_shape"])
    M = np.zeros(shape, dtype=np.float64)
    M[z[prefix + "_row"], z[prefix + "_col"]] = z[prefix + "_val"]
    return M


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def cases(z):
    out = []
    for ci in range(int(z["n_cases"])):
        t = "case%d_" % ci
        kw = json.loads(str(z[t + "kw"]))
        if "bounds" in kw:
            kw["bounds"] = [tuple(b) for b in kw["bounds"]]
        out.append(dict(variant=str(z[t + "variant"]), kw=kw, c=z[t + "c"], g=z[t + "g"], x=z[t + "x"],
                        grad=z[t + "grad"], iters=z[t + "iters"], state=z[t + "state"]))
    return out


def problem_inputs(name, z, i):
    """User-form (A, b, G, h) of instance i of fixture `name`, as the reference drivers pass them."""
    if name == "energy_load1":
        return dict(A=dense(z, "A"), b=z["b"], G=dense(z, "G"), h=z["h"])
    if name == "twitter_sp":
        A = np.delete(dense(z, "Afull"), z["removed_rows"], axis=0)
        return dict(A=A, b=z["b_rr"][i], G=None, h=None)
    if name == "knapsack":
        return dict(A=z["costs"][i][None, :], b=z["budget_rhs"], G=None, h=None)
    if name.startswith("dense"):
        return dict(A=z["A"], b=z["b"], G=None, h=None)
    raise KeyError(name)


def split_state(s, m, n):
    return dict(x=s[:n], y=s[n:n + m], t=s[n + m:2 * n + m], tau=s[2 * n + m], kappa=s[2 * n + m + 1])


ALL = ["knapsack", "energy_load1", "twitter_sp", "dense_m16", "dense_m64"]
