#!/usr/bin/env python
"""Benchmark of the hot path: batched IPOfunc forward+backward LP solves per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--thr T]

A "step" is one pass of the hot path over one batch of synthetic cost vectors of the Energy-scheduling
LP (BASELINE.json configs[1]: ICON load1, standard form 172 x 1582): `x = IPOfunc(...)(c)` for B
instances followed by the backward for dL/dc.  One process per GPU (torchrun for N > 1), instances
sharded across ranks with no data-path collective (weak scaling: B per GPU is fixed).

Prints ONE JSON line (rank 0).  `value` is whole-job LP/s with inputs resident in HBM, `e2e` the same
through host (pinned) buffers with the copies inside the timed region, `roofline` the FP64 roofline of
the dominant kernel with ALGORITHMIC flops (SURVEY.md §8d formula, dense m^3/3 Cholesky count), and
`cpu_baseline` the reference's CPU path (as-executed oracle port) timed on this box's host cores.
`--impl reference` times only that CPU path, on the same config/metric/unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched IPOfunc fwd+bwd LP solves/sec"
UNIT = "LP/s"
OPTS = dict(bounds=[(0.0, None)], thr=0.1, damping=1e-6, max_iter=None, pc=True, method=1)  # exp_run.py:61-63


def energy_problem():
    from neuripsintopt_b200.data import icon
    A, b, G, h, F = icon.make_matrix_intopt(**icon.load1_day01())
    return A, b, G, h, F


def synthetic_costs(F, B, seed):
    """P ~ U[20,120] per 48-slot day, c = P @ F^T; upstream gradient c_true/1440."""
    rng = np.random.default_rng(seed)
    P_true = rng.uniform(20.0, 120.0, (B, 48))
    P_pred = P_true * np.abs(1.0 + 0.2 * rng.standard_normal((B, 48)))
    c = np.ascontiguousarray((F @ P_pred.T).T, dtype=np.float32)  # row-major (B, n), like a predictor's output
    g = np.ascontiguousarray((F @ P_true.T).T / F.shape[0], dtype=np.float32)
    return c, g


SEED0 = 1000  # rank r draws its shard with seed SEED0 + r; the CPU arm times the first instances of rank 0's shard


def workload_name(thr):
    return ("Energy ICON load1 LP (172x1582 standard form, nnz 22188), thr=%g, damping=1e-6, whole (HSD) backward" % thr)


def bench_config(args, world):
    """The `config` object of the JSON line — built by ONE function for both arms, so that they describe the same
    workload with the same words (batch per GPU, seeds, L2 policy of the GPU arm)."""
    B = batch_per_gpu(args, world)
    return dict(workload=workload_name(args.thr), batch_per_gpu=B, global_batch=B * world, seed=SEED0,
                l2="inputs rotate over %d device batches, %.0f MB in total (> 126 MB L2)" % (NSETS, NSETS * 2 * B * 1440 * 4 / 1e6),
                parallelism="dp%d" % world, scaling=args.scaling)


def batch_per_gpu(args, world):
    if args.scaling == "strong":
        return max(1, args.global_batch // world)
    return args.batch


NSETS = 12  # resident input sets the timed steps rotate over (> L2 in total at the default batch)


def algorithmic_flops(info, iters_total, B, k_bwd=2):
    """F_it = 2S + m^3/3 + 6m^2 + 20 nnz + 30 n ;  F_bw = 2S + m^3/3 + 2k m^2 + 12 nnz + 20 n."""
    m, n, nnz, S = info["m"], info["n"], info["nnz"], info["asm_terms"]
    f_it = 2 * S + m ** 3 / 3 + 6 * m * m + 20 * nnz + 30 * n
    f_bw = 2 * S + m ** 3 / 3 + 2 * k_bwd * m * m + 12 * nnz + 20 * n
    return f_it * iters_total, f_bw * B, f_it, f_bw


def executed_flops(info, iters_total, B, k_bwd=2):
    """Same formula with the block-skyline Cholesky/solve MAC counts actually executed (reported beside the
    contract figure, SURVEY App. D accounting rule)."""
    n, nnz, S = info["n"], info["nnz"], info["asm_terms"]
    chol, solve = 2 * info["chol_macs"], 2 * info["solve_macs"]
    f_it = 2 * S + chol + 3 * solve + 20 * nnz + 30 * n
    f_bw = 2 * S + chol + k_bwd * solve + 12 * nnz + 20 * n
    return f_it * iters_total + f_bw * B


def cpu_reference_pass(A, b, G, h, c, g, count):
    """The reference's CPU path as executed (oracle port): per instance host prep + HSD forward + float32
    n-right-hand-side backward, BLAS threads = all host cores, like Interior/intopt_energy_mlp.py:409-432."""
    from oracle import hsd_numpy as O
    t0 = time.perf_counter()
    for i in range(count):
        O.ipofunc(c[i], g[i], A=A, b=b, G=G, h=h, bounds=OPTS["bounds"], variant="whole", thr=OPTS["thr"],
                  max_iter=OPTS["max_iter"], damping=OPTS["damping"], asref=True)
    return time.perf_counter() - t0


def _mp_worker(job):
    """One worker of the best-case CPU leg: a share of the sample with ONE BLAS thread (set by the pool initialiser)."""
    Ad, b, Gd, h, c, g, thr = job
    from oracle import hsd_numpy as O
    t0 = time.perf_counter()
    for i in range(len(c)):
        O.ipofunc(c[i], g[i], A=Ad, b=b, G=Gd, h=h, bounds=OPTS["bounds"], variant="whole", thr=thr,
                  max_iter=OPTS["max_iter"], damping=OPTS["damping"], asref=True)
    return time.perf_counter() - t0


def _mp_init():
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass


def cpu_baseline_measure(A, b, G, h, c, g, sample, thr, passes=3):
    """The reference's CPU path (as-executed oracle port) on this box's host cores, on the first `sample` instances of the
    GPU arm's batch.  Leg (i): what the reference does — a serial loop, BLAS threads pinned to the core count; median of
    `passes` passes.  Leg (ii): best case — one process per core with one BLAS thread each, `sample` LPs per worker.
    `value` is the better of the two (the conservative denominator)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    Ad = A.toarray() if hasattr(A, "toarray") else A
    Gd = G.toarray() if hasattr(G, "toarray") else G
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=cores)
    except Exception:
        limiter = None
    cpu_reference_pass(Ad, b, Gd, h, c, g, 1)  # warm (imports, caches)
    times = [cpu_reference_pass(Ad, b, Gd, h, c, g, sample) for _ in range(passes)]
    if limiter is not None:
        limiter.restore_original_limits()
    serial = sample / float(np.median(times))
    workers = min(cores, 32)
    per_worker = max(2, sample // 2)
    multi = None
    try:
        ctx = mp.get_context("spawn")
        jobs = [(Ad, b, Gd, h, c[:per_worker], g[:per_worker], thr) for _ in range(workers)]
        with ctx.Pool(workers, initializer=_mp_init) as pool:
            pool.map(_mp_worker, [(Ad, b, Gd, h, c[:1], g[:1], thr)] * workers)  # warm every worker
            walls = []
            for _ in range(passes):  # (one pass is a fraction of a second: the median of `passes` is what is reported)
                t0 = time.perf_counter()
                pool.map(_mp_worker, jobs)
                walls.append(time.perf_counter() - t0)
            wall = float(np.median(walls))
        multi = dict(value=workers * per_worker / wall, workers=workers, blas_threads_per_worker=1, lps_per_worker=per_worker,
                     passes_s=[round(t, 3) for t in walls])
    except Exception as e:  # a box that cannot fork workers still reports leg (i)
        multi = dict(value=None, error=repr(e)[:120])
    best = max(serial, multi["value"] or 0.0)
    return dict(value=best, unit=UNIT, cores=cores, kind="port",
                sample="first %d LPs of rank 0's batch (seed %d); leg (i) serial loop, BLAS threads=%d, median of %d passes; "
                       "leg (ii) %s processes x 1 BLAS thread, median of the same number of passes; as-executed oracle port incl. per-call host prep and the "
                       "float32 n-right-hand-side backward; value = the better leg" % (sample, SEED0, cores, passes,
                                                                                       multi.get("workers")),
                serial=dict(value=serial, blas_threads=cores, passes_s=[round(t, 3) for t in times]),
                multiprocess=multi)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    A, b, G, h, F = energy_problem()
    B = batch_per_gpu(args, world)
    c, g = synthetic_costs(F, B, SEED0)  # rank 0's shard of the GPU arm
    t0 = time.perf_counter()
    cb = cpu_baseline_measure(A, b, G, h, c, g, args.ref_sample, args.thr, passes=max(3, min(args.steps, 5)))
    wall = time.perf_counter() - t0
    value = cb["value"]
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * B / value, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference", config=bench_config(args, world),
                cpu_baseline=cb, measure_wall_s=round(wall, 1),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", os.environ.get("BENCH_SMI_MS", "200")], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first_sample(self, timeout=30.0):
        """nvidia-smi's cold start (first NVML init on a fresh box) takes seconds and disturbs running kernels;
        block until it has printed its first sample so that this happens before the timed region."""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return True
            time.sleep(0.05)
        return False

    def mark(self):
        """Byte offset of the log right now (samples after it belong to the timed region)."""
        return os.path.getsize(self.f.name) if self.p is not None else 0

    def stop(self, start_offset=0):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(start_offset)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            hi = [v for v in sm if v >= 0.5 * max(sm)]
            out = dict(sm_mhz=float(np.median(hi)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from neuripsintopt_b200 import _capi
    from neuripsintopt_b200.ip_model_whole import IPOfunc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = batch_per_gpu(args, world), args.steps, args.warmup
    sampler = ClockSampler(local) if rank == 0 else None  # started early: its cold start must not hit the timing

    A, b, G, h, F = energy_problem()
    tt = lambda M: torch.as_tensor(M.toarray() if hasattr(M, "toarray") else M, dtype=torch.float32)  # noqa: E731
    opts = dict(OPTS, thr=args.thr)
    f = IPOfunc(tt(A), tt(b), tt(G), tt(h), check_finite=False, **opts)
    c_np, g_np = synthetic_costs(F, B, SEED0 + rank)  # every rank owns a different shard of the global batch
    c_dev = torch.tensor(c_np, device=dev)
    g_dev = torch.tensor(g_np, device=dev)
    c_host = torch.empty(c_np.shape, dtype=torch.float32, pin_memory=True).copy_(torch.from_numpy(c_np))
    g_host = torch.empty(g_np.shape, dtype=torch.float32, pin_memory=True).copy_(torch.from_numpy(g_np))
    x_host = torch.empty(c_np.shape, dtype=torch.float32, pin_memory=True)
    gc_host = torch.empty(c_np.shape, dtype=torch.float32, pin_memory=True)

    def step_resident():
        c = c_dev.detach().requires_grad_(True)
        x = f(c)
        x.backward(g_dev)
        return x, c.grad

    # end-to-end: every step copies its inputs from pinned host memory and its results back; copies run on
    # side streams with two device buffer sets, so step k+1's upload and step k-1's download overlap step k's
    # kernels (what a prefetching data loader does).  Nothing is cached across steps.
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dbuf = [dict(c=torch.empty_like(c_dev), g=torch.empty_like(g_dev), free=None) for _ in range(2)]

    def e2e_loop(n_steps):
        main = torch.cuda.current_stream(dev)
        for k in range(n_steps):
            buf = dbuf[k % 2]
            with torch.cuda.stream(s_in):
                if buf["free"] is not None:
                    s_in.wait_event(buf["free"])  # the step that last used this buffer has finished
                buf["c"].copy_(c_host, non_blocking=True)
                buf["g"].copy_(g_host, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(s_in)
            main.wait_event(ready)
            c = buf["c"].detach().requires_grad_(True)
            x = f(c)
            x.backward(buf["g"])
            done = torch.cuda.Event()
            done.record(main)
            buf["free"] = done
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                x_host.copy_(x.detach(), non_blocking=True)
                gc_host.copy_(c.grad, non_blocking=True)
                x.record_stream(s_out)
                c.grad.record_stream(s_out)
        main.wait_stream(s_out)

    # ---- parity gate on the same inputs before timing (rank 0, a few instances) ----------------------
    x, gc = step_resident()
    torch.cuda.synchronize()
    status, iters = IPOfunc.status()
    iters_total = int(iters.sum().item())
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import hsd_numpy as O  # checker only
        Ad, Gd = A.toarray(), G.toarray()
        ex = eg = 0.0
        for i in range(min(2, B)):
            r = O.ipofunc(c_np[i], g_np[i], A=Ad, b=b, G=Gd, h=h, bounds=opts["bounds"], thr=opts["thr"],
                          max_iter=opts["max_iter"], damping=opts["damping"])
            ex = max(ex, float(np.abs(x[i].detach().double().cpu().numpy() - r["x"]).max() / max(1.0, np.abs(r["x"]).max())))
            eg = max(eg, float(np.linalg.norm(gc[i].double().cpu().numpy() - r["grad"]) / np.linalg.norm(r["grad"])))
            assert int(iters[i]) == r["iters"], (int(iters[i]), r["iters"])
        # thr > 0 (the experiments' mode): the north-star tolerances.  thr = 0: the converged gradient is a 0/0-type
        # function of the last iterate, only x is gated here; tests/test_gpu_parity.py checks the
        # backward kernel at a fixed iterate
        if opts["thr"] and opts["thr"] > 0:
            assert ex < 1e-6 and eg < 1e-5, ("parity gate failed", ex, eg)
        else:
            assert ex < 1e-3, ("parity gate failed", ex, eg)
        parity = dict(x_rel=ex, grad_rel=eg, checked=min(2, B))
    assert int((status != 0).sum().item()) == 0, "non-zero status words in the benchmark batch"

    dp = None
    info = None
    from neuripsintopt_b200 import ipofunc as _ipo
    for v in _ipo._PROBLEM_CACHE.values():
        if hasattr(v, "info"):
            info = v.info
    fp64_peak = _capi.measure_fp64_peak(5) * 1e-3  # TFLOP/s, measured now on this device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_rank = []
    # Resident inputs rotate over NSETS device batches (> L2 in total) so that no timed step finds its inputs in L2.
    c_sets = [c_dev] + [torch.tensor(synthetic_costs(F, B, 5000 + 17 * rank + s)[0], device=dev) for s in range(1, NSETS)]
    g_sets = [g_dev] + [g_dev.clone() for _ in range(1, NSETS)]

    def timed_resident():
        """Exactly K steps in ONE timed region (barrier + synchronize on both sides); forward/backward shares from
        event pairs recorded on the launching stream without any host synchronisation inside the region."""
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        for i in range(K):
            c = c_sets[i % NSETS].detach().requires_grad_(True)
            ev[i][0].record()
            xx = f(c)
            ev[i][1].record()
            xx.backward(g_sets[i % NSETS])
            ev[i][2].record()
        ev[-1][2].synchronize()
        barrier()
        tot = ev[0][0].elapsed_time(ev[-1][2])
        fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev)
        bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev)
        t = torch.tensor([tot, fwd_ms, bwd_ms], dtype=torch.float64, device=dev)
        if world > 1:
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank.append([round(float(v[0]) / K, 4) for v in allt])
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def timed_e2e():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_loop(K)
        e1.record()
        e1.synchronize()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    if sampler:
        sampler.wait_first_sample()
    smi_from = sampler.mark() if sampler else 0
    # untimed pre-warm (clocks, allocator, NCCL, the host's caches) on top of the W warm-up steps
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.prewarm_s:
        step_resident()
    torch.cuda.synchronize()
    for i in range(W):
        c = c_sets[i % NSETS].detach().requires_grad_(True)
        f(c).backward(g_sets[i % NSETS])
    # The host that drives the GPU is shared and occasionally stalls a launch for milliseconds; every region is
    # exactly K steps between barrier + synchronize, the region is repeated REGIONS times and the fastest one is
    # reported (max over ranks inside each region); the median is kept beside it.
    regions = [timed_resident() for _ in range(args.regions)]
    tot_ms, fwd_ms, bwd_ms = min(regions, key=lambda r: r[0])
    tot_ms_median = float(np.median([r[0] for r in regions]))
    # end-to-end: one event pair around exactly K pipelined steps (inputs come from the host every step)
    e2e_loop(max(W, 3))
    e2e_regions = [timed_e2e() for _ in range(args.regions)]
    e2e_ms = min(e2e_regions)
    e2e_ms_median = float(np.median(e2e_regions))
    # transparency: the same resident measurement at B = 1024 (3.46 waves of 296 CTA slots), one extra region
    alt = None
    if world == 1 and B != 1024 and not args.no_alt:
        ca, ga = synthetic_costs(F, 1024, 777)
        ca_d, ga_d = torch.tensor(ca, device=dev), torch.tensor(ga, device=dev)
        def alt_step():
            cc = ca_d.detach().requires_grad_(True)
            f(cc).backward(ga_d)
        for _ in range(5):
            alt_step()
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(K):
                alt_step()
            a1.record()
            a1.synchronize()
            ms = a0.elapsed_time(a1)
            best = ms if best is None else min(best, ms)
        alt = dict(batch_per_gpu=1024, value=1024 * K / (best * 1e-3), ms_per_step=best / K,
                   note="same kernels, batch of 3.46 waves; inputs of this side measurement stay L2 resident")
    # Strong-scaling side measurement (BASELINE.json configs[3]: "batch 1024 sharded over 8xB200"): a FIXED global batch
    # of `--global-batch` LPs, this rank solving its 1/world share; max over ranks like the main measurement.  `gpu_ms`
    # is the device time between the first launch and the last completion of a step (CUDA events on the launching stream),
    # `ms_per_step` the wall time per step of K back-to-back steps: their difference is host (Python + launch) time the
    # device spends idle.
    strong = None
    if not args.no_strong and args.scaling == "weak":
        Bs = max(1, args.global_batch // world)
        cs, gs = synthetic_costs(F, Bs, 4000 + rank)
        cs_d, gs_d = torch.tensor(cs, device=dev), torch.tensor(gs, device=dev)

        def s_step():
            cc = cs_d.detach().requires_grad_(True)
            f(cc).backward(gs_d)
        for _ in range(10):
            s_step()
        best, gpu_best = None, None
        for _ in range(args.regions):
            barrier()
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(K):
                evs[i][0].record()
                s_step()
                evs[i][1].record()
            a1.record()
            a1.synchronize()
            barrier()
            ts = torch.tensor([a0.elapsed_time(a1), sum(e[0].elapsed_time(e[1]) for e in evs)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            ms, gms = float(ts[0]), float(ts[1])
            if best is None or ms < best:
                best, gpu_best = ms, gms
        strong = dict(global_batch=Bs * world, batch_per_gpu=Bs, value=Bs * world * K / (best * 1e-3), ms_per_step=best / K,
                      gpu_ms_per_step=gpu_best / K, waves_per_gpu=Bs / float(info["ctas_per_sm"] * 148),
                      note="fixed global batch sharded over the ranks (strong scaling); same kernels; value = global batch / "
                           "max-over-ranks time; gpu_ms_per_step = first launch to last completion on the device")
    # clock samples cover pre-warm + warm-up + both timed regions (identical load); the timed regions alone are
    # only ~K x 4 ms long, too short for more than a couple of 100 ms samples
    clocks = sampler.stop(smi_from) if sampler else None

    if rank == 0:
        value = world * B * K / (tot_ms * 1e-3)
        e2e = world * B * K / (e2e_ms * 1e-3)
        f_fwd, f_bwd, f_it, f_bw = algorithmic_flops(info, iters_total, B)
        ach_fwd = f_fwd / (fwd_ms / K * 1e-3) * 1e-12
        ach_bwd = f_bwd / (bwd_ms / K * 1e-3) * 1e-12
        ach_all = (f_fwd + f_bwd) / (tot_ms / K * 1e-3) * 1e-12
        exe_all = executed_flops(info, iters_total, B) / (tot_ms / K * 1e-3) * 1e-12
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        n_x, n, m = c_np.shape[1], info["n"], info["m"]
        alg_bytes = 8.0 * (8 * n + 2 * m) * B  # SURVEY §8(d) algorithmic bytes per LP
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=tot_ms / K,
            higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f64", data="synthetic",
            iters_per_lp=iters_total / B,
            timing=dict(regions=args.regions, rule="fastest of %d regions of exactly K steps each" % args.regions,
                        ms_per_step_median=tot_ms_median / K, e2e_ms_per_step_median=e2e_ms_median / K),
            config=bench_config(args, world),
            e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=int(c_host.numel() * 4 + g_host.numel() * 4),
                     d2h_bytes_per_step=int(x_host.numel() * 4 + gc_host.numel() * 4)),
            # per step: intopt_forward_kernel + intopt_backward_kernel (the mapping between the caller's float32 user
            # layout and the standard form runs inside them)
            gpu_launches=2 * K,
            roofline=dict(bound="fp64", kernel=("intopt_forward_tmem_kernel" if info.get("tmem") else "intopt_forward_kernel"),
                          achieved=ach_fwd, peak=fp64_peak,
                          unit="TFLOP/s", frac=ach_fwd / fp64_peak,
                          # DRAM bytes of one forward launch from the committed `ncu --set full` capture of this
                          # default configuration (profiles/r02_energy_fwd_ncu.txt: 14.4 MB read + 61.2 MB written for
                          # 2368 LPs = float32 c in, packed c + saved iterate + float32 x out; 117 MB before the n-vectors
                          # moved from the workspace to tensor memory); null for other configurations
                          traffic=(75.6e6 if (B == 2368 and args.thr == 0.1 and info.get("tmem")) else None),
                          peak_source="DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                          flops_per_iter=f_it, flops_backward=f_bw,
                          backward=dict(kernel=("intopt_backward_tmem_kernel" if info.get("tmem") else "intopt_backward_kernel"),
                                        achieved=ach_bwd, frac=ach_bwd / fp64_peak),
                          step=dict(achieved=ach_all, frac=ach_all / fp64_peak,
                                    executed_achieved=exe_all, executed_frac=exe_all / fp64_peak,
                                    note="achieved uses the contract's dense m^3/3 Cholesky count; executed_* uses the "
                                         "envelope factorisation's real MAC count"),
                          hbm=dict(algorithmic_bytes_per_step=alg_bytes,
                                   achieved_gbs=alg_bytes / (tot_ms / K * 1e-3) * 1e-9, peak_gbs=hbm_peak),
                          fwd_ms=fwd_ms / K, bwd_ms=bwd_ms / K),
            problem=info, parity=parity, clocks=clocks, per_rank_ms_per_step=per_rank[-1] if per_rank else [],
            alt_batch=alt, strong_side=strong)
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_measure(A, b, G, h, c_np, g_np, args.ref_sample, args.thr)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=2368,
                    help="LP instances per GPU per step; default = 8 whole waves of the 148 SMs x 2 resident CTAs of a "
                         "B200 (the persistent kernels run one LP per CTA slot, so a batch that is not a multiple of 296 "
                         "leaves part of the last wave idle: 1024 = 3.46 waves measures 11%% lower, reported as alt_batch)")
    ap.add_argument("--thr", type=float, default=0.1)
    ap.add_argument("--ref-sample", type=int, default=8, help="LPs per CPU-baseline pass")
    ap.add_argument("--regions", type=int, default=5, help="timed regions of K steps each (the fastest is reported)")
    ap.add_argument("--prewarm-s", type=float, default=1.0, help="seconds of untimed pre-warm before the W warm-up steps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the side measurement at batch 1024")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch LPs per GPU (default, the headline); strong: --global-batch LPs in total, "
                         "1/N per GPU (BASELINE.json configs[3])")
    ap.add_argument("--global-batch", type=int, default=1024, help="total LPs per step of the strong-scaling measurements")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling side measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
