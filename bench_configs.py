#!/usr/bin/env python
"""Secondary benchmark: batched IPOfunc fwd+bwd throughput on every BASELINE.json config shape
(bench.py stays the contract benchmark on the Energy LP).  Prints one JSON line per config.

    python bench_configs.py [--batch B] [--steps K]

Configs: knapsack (32x62, per-instance cost row => one problem per instance, shared structure cache),
energy (172x1582), twitter (221x2852, per-instance b, shared c), dense m in {16, 64} (n = 2m).
Inputs come from tests/golden/*.npz (problem data) with synthetic cost batches."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, ROOT)
import torch
from tests import golden_util as GU


def time_steps(fn, steps, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1024); ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--thr", type=float, default=0.1)
    ap.add_argument("--no-dense", action="store_true", help="skip the dense (config 5) lines")
    args = ap.parse_args()
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    from neuripsintopt_b200 import ipofunc as _ipo
    tt = lambda v: torch.Tensor() if v is None else torch.as_tensor(np.asarray(v), dtype=torch.float32)
    from neuripsintopt_b200 import _capi
    rng = np.random.default_rng(0)
    B = args.batch
    out = []
    peak = _capi.measure_fp64_peak(5) * 1e-3  # TFLOP/s, DFMA micro-kernel on this device (as in bench.py)
    for name in ("energy_load1", "twitter_sp", "dense_m16", "dense_m64", "knapsack"):
        z = GU.load(name); case = GU.cases(z)[0]; kw = dict(case["kw"]); kw["thr"] = args.thr
        inp = GU.problem_inputs(name, z, 0)
        n_x = case["c"].shape[1]
        c0 = case["c"][0].astype(np.float64)
        if name == "twitter_sp":
            pairs = z["train_pairs"]; nb = z["b_rr"].shape[1]
            bb = np.tile(z["b_rr"], (B // len(z["b_rr"]) + 1, 1))[:B]
            c = torch.tensor((c0 * np.abs(1 + 0.3 * rng.standard_normal((B, n_x)))).astype(np.float32), device="cuda")
            f = IPOfunc(tt(inp["A"]), tt(bb), torch.Tensor(), torch.Tensor(), check_finite=False, **kw)
        elif name == "knapsack":
            # one cost row per instance (the minibatch's cost vector is the equality row, optimizer.py:396-405): batched
            # through the per-instance-values path, A of shape (B, 1, 31)
            costs = rng.uniform(20, 1000, (B, 1, n_x)).astype(np.float32)
            c = torch.tensor(-(rng.standard_normal((B, n_x)) + 0.3 * rng.standard_normal((B, n_x))).astype(np.float32), device="cuda")
            f = IPOfunc(tt(costs), tt(z["budget_rhs"]), torch.Tensor(), torch.Tensor(), check_finite=False, **kw)
        else:
            scale = np.abs(c0).mean()
            c = torch.tensor((c0 + 0.05 * scale * rng.standard_normal((B, n_x))).astype(np.float32), device="cuda")
            f = IPOfunc(tt(inp["A"]), tt(inp["b"]), tt(inp["G"]), tt(inp["h"]), check_finite=False, **kw)
        g = torch.ones(B, n_x, device="cuda") / n_x
        def step():
            cc = c.detach().requires_grad_(True); f(cc).backward(g)
        step(); torch.cuda.synchronize()
        status, iters = IPOfunc.status()
        ms = time_steps(step, args.steps)
        info = [v.info for v in _ipo._PROBLEM_CACHE.values() if hasattr(v, "info")][-1]
        # algorithmic flops per LP: F_it = 2S + m^3/3 + 6m^2 + 20 nnz + 30 n per iteration,
        # F_bw = 2S + m^3/3 + 4m^2 + 12 nnz + 20 n for the (whole) backward
        m_, n_, nnz_, S_ = info["m"], info["n"], info["nnz"], info["asm_terms"]
        f_it = 2 * S_ + m_ ** 3 / 3 + 6 * m_ ** 2 + 20 * nnz_ + 30 * n_
        f_bw = 2 * S_ + m_ ** 3 / 3 + 4 * m_ ** 2 + 12 * nnz_ + 20 * n_
        flops = float(iters.float().sum()) * f_it + B * f_bw
        tf = flops / (ms * 1e-3) / 1e12
        out.append(dict(config=name, shape="%dx%d nnz %d" % (info["m"], info["n"], info["nnz"]), batch=B,
                        lp_per_s=B / ms * 1e3, ms_per_step=ms, thr=args.thr, mean_iters=float(iters.float().mean()),
                        roofline=dict(bound="fp64", achieved=tf, peak=peak, unit="TFLOP/s", frac=tf / peak,
                                      flops_per_iter=f_it, flops_backward=f_bw),
                        bad_status=int((status != 0).sum()), smem_bytes=info["smem_bytes"], ctas_per_sm=info["ctas_per_sm"],
                        envelope=info["envelope"]))
        print(json.dumps(out[-1]))
    # config 5: dense LPs through the lock-step dense path (batched DMMA assembly, tiled Cholesky); roofline against the
    # MEASURED FP64 tensor-core peak.  Batches chosen so that a step stays within seconds; bench_dense.py runs the sweep.
    if not args.no_dense:
        import bench_dense
        dmma = _capi.measure_dmma_peak(5) * 1e-3
        for m_, B_ in ((256, 2048), (512, 512), (1024, 256)):
            line = bench_dense.measure(m_, B_, args.thr if args.thr > 0 else None, steps=3, parity=1, dmma=dmma, dfma=peak)
            line["roofline"] = dict(bound="fp64 tensor (DMMA)", achieved=line["achieved_tflops"], peak=dmma, unit="TFLOP/s",
                                    frac=line["roofline_frac_dmma"])
            print(json.dumps(line))


if __name__ == "__main__":
    main()
