"""Diagnostic (not a test): prints GPU-vs-oracle gaps for every golden case.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import hsd_numpy as O
from tests import golden_util as GU
from tests.test_gpu_parity import _run_gpu, _relinf, _rel2

for name in GU.ALL:
    z = GU.load(name)
    for case in GU.cases(z):
        kw = dict(case["kw"]); variant = case["variant"]; thr = kw.get("thr") or 0.0
        damping = kw.get("damping", 1e-2)
        k = len(case["c"])
        if name == "knapsack":
            outs = [_run_gpu(variant, kw, GU.problem_inputs(name, z, i), case["c"][i], case["g"][i]) for i in range(k)]
            xs = np.stack([o[0] for o in outs]); gs = np.stack([o[1] for o in outs])
            status = np.concatenate([o[2] for o in outs]); iters = np.concatenate([o[3] for o in outs])
            end = {kk: np.concatenate([o[4][kk] for o in outs]) for kk in outs[0][4]}
        else:
            inp = GU.problem_inputs(name, z, 0)
            if name == "twitter_sp":
                inp = dict(inp, b=z["b_rr"][:k])
            xs, gs, status, iters, end = _run_gpu(variant, kw, inp, case["c"], case["g"])
        for i in range(k):
            inp_i = GU.problem_inputs(name, z, i)
            r = O.ipofunc(case["c"][i], case["g"][i], variant=variant, bounds=kw.get("bounds"), thr=kw.get("thr"),
                          max_iter=kw.get("max_iter", 1000), damping=damping, **inp_i)
            sf = r["sf"]
            c64 = case["c"][i].astype(np.float64)
            st = dict(x=end["x"][i], y=end["y"][i], t=end["t"][i], tau=float(end["tau"][i]), kappa=float(end["kappa"][i]))
            fn = O.backward_whole if variant == "whole" else O.backward_alter
            dc = fn(sf["A"], sf["b"], sf["c"], st, case["g"][i].astype(np.float64)[sf["var_index"]], damping, sf["nvar"])
            print("%-13s %-5s thr=%-6g i=%d it gpu/orc/ref %d/%d/%d st %d/%d | x-orc %.1e x-ref32 %.1e obj %.1e | g-orc %.1e g@gpustate %.1e g-ref32 %.1e" % (
                name, variant, thr, i, iters[i], r["iters"], case["iters"][i], status[i], r["status"],
                _relinf(xs[i], r["x"]), _relinf(xs[i], case["x"][i].astype(np.float64)),
                abs(c64 @ xs[i] - c64 @ r["x"]) / max(1, abs(c64 @ r["x"])),
                _rel2(gs[i], r["grad"]), _rel2(gs[i][sf["var_index"]], dc), _rel2(gs[i], case["grad"][i].astype(np.float64))))
