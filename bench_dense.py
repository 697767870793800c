#!/usr/bin/env python
"""Dense LP sweep (BASELINE.json configs[4]): batched IPOfunc forward+backward on synthetic
feasible-and-bounded dense LPs (m x 2m) through the lock-step dense path, roofline fraction against the MEASURED
FP64 tensor-core (DMMA) peak, parity of the first instances against the CPU oracle.  One JSON line per shape.

    python bench_dense.py [--shapes 256x512,1024x256,...]  (m x batch PER GPU)  [--thr 0.1] [--steps 5]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_dense.py --gpus N ...

With N > 1 every rank solves its own batch of the same problem (weak scaling, no data-path collective); the line
reports the whole-job LP/s over the max-over-ranks time.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def gen(m, B, seed=None):
    rng = np.random.default_rng(m)
    n = 2 * m
    A = rng.standard_normal((m, n)).astype(np.float32)
    x0 = rng.uniform(0.1, 1.0, n)
    b = (A.astype(np.float64) @ x0).astype(np.float32)
    if seed is not None:
        rng = np.random.default_rng(seed)  # per-rank costs for the same problem
    y0 = rng.standard_normal((B, m))
    s0 = rng.uniform(0.1, 1.0, (B, n))
    c = (y0 @ A.astype(np.float64) + s0).astype(np.float32)
    return A, b, c


def contract_flops(m, n, iters_sum, B):
    """F_it = 2S + m^3/3 + 6 m^2 + 20 nnz + 30 n, F_bw = 2S + m^3/3 + 4 m^2 + 12 nnz + 20 n (whole)."""
    S = n * m * (m + 1) / 2.0
    nnz = m * n
    f_it = 2 * S + m ** 3 / 3.0 + 6 * m * m + 20 * nnz + 30 * n
    f_bw = 2 * S + m ** 3 / 3.0 + 4 * m * m + 12 * nnz + 20 * n
    return iters_sum * f_it + B * f_bw


def measure(m, B, thr, steps=5, parity=1, dmma=None, dfma=None, rank=0, world=1):
    import torch.distributed as dist
    from neuripsintopt_b200 import _capi, ipofunc
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    n = 2 * m
    A, b, c = gen(m, B, seed=None if world == 1 else 7000 + rank)
    ipofunc.clear_cache()
    t0 = time.time()
    f = IPOfunc(torch.tensor(A), torch.tensor(b), torch.Tensor(), torch.Tensor(), bounds=[(0.0, None)], thr=thr,
                damping=1e-6, check_finite=False)
    dev = torch.device("cuda", torch.cuda.current_device())
    ct = torch.tensor(c, device=dev, requires_grad=True)
    g = torch.ones_like(ct)
    x = f(ct)
    x.backward(g)
    torch.cuda.synchronize()
    t_first = time.time() - t0
    status, iters = IPOfunc.status()
    info = [v.info for v in ipofunc._PROBLEM_CACHE.values() if hasattr(v, "info")][-1]
    ex = eg = None
    if parity and rank == 0:
        from oracle import hsd_numpy as O  # checker only
        ex = eg = 0.0
        for i in range(parity):
            r = O.ipofunc(c[i], np.ones(n, np.float32), A=A, b=b, bounds=[(0.0, None)], thr=thr, damping=1e-6)
            xi = x[i].detach().double().cpu().numpy()
            ex = max(ex, float(np.abs(xi - r["x"]).max() / max(1.0, np.abs(r["x"]).max())))
            eg = max(eg, float(np.linalg.norm(ct.grad[i].double().cpu().numpy() - r["grad"]) / np.linalg.norm(r["grad"])))
            if thr:
                assert int(iters[i]) == r["iters"], (int(iters[i]), r["iters"])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tfs, tbs = [], []
    for _ in range(steps):
        cc = ct.detach().requires_grad_(True)
        if world > 1:
            dist.barrier()
        ev[0].record()
        xx = f(cc)
        ev[1].record()
        xx.backward(g)
        ev[2].record()
        ev[2].synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tfs.append(float(t[0]))
        tbs.append(float(t[1]))
    tf, tb = float(np.median(tfs)), float(np.median(tbs))
    ms = tf + tb
    it = torch.tensor([float(iters.sum()), float((status != 0).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(it)
    fl = contract_flops(m, n, float(it[0]), B * world)
    tfl = fl / (ms * 1e-3) * 1e-12
    if dmma is None:
        dmma = _capi.measure_dmma_peak(5) * 1e-3
    if dfma is None:
        dfma = _capi.measure_fp64_peak(5) * 1e-3
    return {"config": "dense m=%d n=%d" % (m, n), "batch_per_gpu": B, "n_gpus": world, "thr": thr or 0.0, "path": info["path"],
            "lp_per_s": round(B * world / ms * 1e3, 1), "ms_fwd": round(tf, 3), "ms_bwd": round(tb, 3),
            "iters_mean": round(float(it[0]) / (B * world), 2), "bad_status": int(it[1]),
            "contract_gflop": round(fl * 1e-9, 2), "achieved_tflops": round(tfl, 3),
            "dmma_peak_tflops_per_gpu": round(dmma, 2), "dfma_peak_tflops_per_gpu": round(dfma, 2),
            "roofline_frac_dmma": round(tfl / (dmma * world), 4), "roofline_frac_dfma": round(tfl / (dfma * world), 4),
            "x_rel_vs_oracle": ex, "grad_rel_vs_oracle": eg, "first_call_s": round(t_first, 2),
            "ms_fwd_all": [round(v, 2) for v in tfs], "ms_bwd_all": [round(v, 2) for v in tbs]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="256x1024,512x512,1024x256")
    ap.add_argument("--thr", type=float, default=0.1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--parity", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    import torch.distributed as dist
    from neuripsintopt_b200 import _capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dmma = _capi.measure_dmma_peak(5) * 1e-3
    dfma = _capi.measure_fp64_peak(5) * 1e-3
    thr = args.thr if args.thr > 0 else None
    for shape in args.shapes.split(","):
        m, B = (int(v) for v in shape.split("x"))
        line = measure(m, B, thr, args.steps, args.parity, dmma, dfma, rank, world)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
