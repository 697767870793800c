"""Diagnostic (not a test): quantifies converged-mode (thr = 0, tol 1e-8) parity for every golden fixture.
For each converged case: GPU vs oracle (x, objective, iterations, gradient at the GPU's own iterate), the oracle's own
noise floor (the same LP with its constraint rows permuted: identical mathematics, different rounding), and the
reference's float32 output vs the oracle.  Prints one line per case; profiles/r02_converged_parity.txt is this output."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import hsd_numpy as O
from tests import golden_util as GU
from tests.test_gpu_parity import _rel2, _relinf, _run_gpu


def permuted(inp, seed=0):
    rng = np.random.default_rng(seed)
    out = dict(inp)
    if inp.get("A") is not None and np.ndim(inp["A"]) == 2 and len(inp["A"]) > 1:
        p = rng.permutation(len(inp["A"]))
        out["A"], out["b"] = np.asarray(inp["A"])[p], np.asarray(inp["b"])[p]
    if inp.get("G") is not None and len(inp["G"]) > 1:
        p = rng.permutation(len(inp["G"]))
        out["G"], out["h"] = np.asarray(inp["G"])[p], np.asarray(inp["h"])[p]
    return out


def main():
    print("%-14s %-6s %4s | %9s %9s %5s %9s | %9s %9s | %9s" % ("fixture", "var", "inst", "x gpu-orc", "obj gpu-o", "d_it", "grad@iter",
                                                                  "x orc-prm", "obj o-prm", "x ref32-o"))
    worst = {}
    for name in GU.ALL:
        z = GU.load(name)
        for case in GU.cases(z):
            kw = dict(case["kw"])
            if (kw.get("thr") or 0.0) > 0:
                continue
            variant, damping = case["variant"], kw.get("damping", 1e-2)
            for i in range(len(case["c"])):
                inp = GU.problem_inputs(name, z, i)
                xs, gs, status, iters, end = _run_gpu(variant, kw, inp, case["c"][i:i + 1] if np.ndim(case["c"]) > 1 else case["c"],
                                                      case["g"][i:i + 1] if np.ndim(case["g"]) > 1 else case["g"])
                okw = dict(variant=variant, bounds=kw.get("bounds"), thr=kw.get("thr"), max_iter=kw.get("max_iter", 1000), damping=damping)
                r = O.ipofunc(case["c"][i], case["g"][i], **okw, **inp)
                rp = O.ipofunc(case["c"][i], case["g"][i], **okw, **permuted(inp))
                c64 = case["c"][i].astype(np.float64)
                x = xs[0]
                sf = r["sf"]
                st = dict(x=end["x"][0], y=end["y"][0], t=end["t"][0], tau=float(end["tau"][0]), kappa=float(end["kappa"][0]))
                fn = O.backward_whole if variant == "whole" else O.backward_alter
                dc = fn(sf["A"], sf["b"], sf["c"], st, case["g"][i].astype(np.float64)[sf["var_index"]], damping, sf["nvar"])
                row = dict(x=_relinf(x, r["x"]), obj=abs(c64 @ x - c64 @ r["x"]) / max(1.0, abs(c64 @ r["x"])),
                           dit=int(iters[0]) - r["iters"], g=_rel2(gs[0][sf["var_index"]], dc),
                           xp=_relinf(rp["x"], r["x"]), objp=abs(c64 @ rp["x"] - c64 @ r["x"]) / max(1.0, abs(c64 @ r["x"])),
                           xr=_relinf(case["x"][i].astype(np.float64), r["x"]))
                print("%-14s %-6s %4d | %9.2e %9.2e %5d %9.2e | %9.2e %9.2e | %9.2e" % (name, variant, i, row["x"], row["obj"], row["dit"],
                                                                                         row["g"], row["xp"], row["objp"], row["xr"]), flush=True)
                w = worst.setdefault(name, dict(x=0, obj=0, g=0, xp=0))
                for k in w:
                    w[k] = max(w[k], row[k])
    print()
    for name, w in worst.items():
        print("worst %-14s x gpu-oracle %.2e   objective %.2e   grad at the same iterate %.2e   oracle row-permutation noise on x %.2e"
              % (name, w["x"], w["obj"], w["g"], w["xp"]))



if __name__ == "__main__":
    main()
