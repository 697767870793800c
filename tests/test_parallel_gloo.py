"""CPU tests (gloo, world_size 2) of the data-parallel plumbing: contiguous sharding of the instance batch
and the predictor-gradient all-reduce.  The LP layer itself needs a GPU, so a smooth surrogate layer with the
same (B,n)->(B,n) signature stands in for it here; the GPU test test_gpu_parity.py::
test_batch_position_independence covers the other half (per-instance results do not depend on the shard)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuripsintopt_b200 import parallel
from neuripsintopt_b200.data import icon
from neuripsintopt_b200.energy_learner import IntoptEnergyLearner, days


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 24, 1024):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _surrogate(c):  # smooth stand-in for the LP layer, instance-wise
    return torch.softmax(-c / c.abs().mean(dim=1, keepdim=True), dim=1)


def _make_learner():
    torch.manual_seed(0)
    return IntoptEnergyLearner(icon.load1_day01(), input_size=4, hidden_size=3, num_layers=2, device="cpu",
                               lp_layer=_surrogate, optimizer=torch.optim.SGD, lr=1e-3)


def _data():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((48 * 6, 4)).astype(np.float32)
    y = rng.uniform(20, 120, 48 * 6).astype(np.float32)
    return days(X, y)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lrn = _make_learner()
        lrn.model.eval()  # dropout off so that both runs are comparable
        lrn.model.train = lambda *a, **k: lrn.model  # keep eval mode inside training_step
        Xd, yd = _data()
        loss = lrn.training_step(Xd, yd, rank=rank, world=world)
        flat = torch.cat([p.detach().reshape(-1) for p in lrn.model.parameters()])
        out[rank] = (loss, flat.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_single_process_step():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ref = _make_learner()
    ref.model.eval()
    ref.model.train = lambda *a, **k: ref.model
    Xd, yd = _data()
    loss = ref.training_step(Xd, yd)
    flat = torch.cat([p.detach().reshape(-1) for p in ref.model.parameters()]).numpy()
    assert abs(out[0][0] + out[1][0] - loss) < 1e-4 * abs(loss)
    np.testing.assert_allclose(out[0][1], flat, rtol=1e-5, atol=1e-7)  # same parameters after the step
    np.testing.assert_allclose(out[1][1], flat, rtol=1e-5, atol=1e-7)


def _worker_uneven(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # different initial parameters per rank: the constructor must broadcast rank 0's
        lrn = IntoptEnergyLearner(icon.load1_day01(), input_size=4, hidden_size=3, num_layers=2, device="cpu",
                                  lp_layer=_surrogate, optimizer=torch.optim.SGD, lr=1e-3)
        lrn.model.eval()
        lrn.model.train = lambda *a, **k: lrn.model
        Xd, yd = _data()
        loss = lrn.training_step(Xd[:1], yd[:1], rank=rank, world=world)  # ONE day for two ranks: rank 1's shard is empty
        flat = torch.cat([p.detach().reshape(-1) for p in lrn.model.parameters()])
        out[rank] = (loss, flat.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_more_ranks_than_instances_and_unequal_start():
    """A rank with an empty shard still joins the gradient all-reduce (zero gradients, same buffer layout), and every
    rank starts from rank 0's parameters: after the step all replicas are equal and equal to the one-process step."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_uneven, args=(2, port, out), nprocs=2, join=True)
    torch.manual_seed(100)
    ref = IntoptEnergyLearner(icon.load1_day01(), input_size=4, hidden_size=3, num_layers=2, device="cpu",
                              lp_layer=_surrogate, optimizer=torch.optim.SGD, lr=1e-3)
    ref.model.eval()
    ref.model.train = lambda *a, **k: ref.model
    Xd, yd = _data()
    loss = ref.training_step(Xd[:1], yd[:1])
    flat = torch.cat([p.detach().reshape(-1) for p in ref.model.parameters()]).numpy()
    assert out[1][0] == 0.0 and abs(out[0][0] - loss) < 1e-4 * abs(loss)
    np.testing.assert_allclose(out[0][1], flat, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out[1][1], flat, rtol=1e-5, atol=1e-7)


def test_icon_builder_matches_reference_matrices():
    """make_matrix_intopt (sparse, index-range based) reproduces the reference's dense torch builder output
    stored in the golden fixture (Interior/intopt_energy_mlp.py:124-168)."""
    from tests import golden_util as GU
    z = GU.load("energy_load1")
    A, b, G, h, F = icon.make_matrix_intopt(**icon.load1_day01())
    np.testing.assert_array_equal(A.toarray(), GU.dense(z, "A"))
    np.testing.assert_array_equal(G.toarray(), GU.dense(z, "G"))
    np.testing.assert_array_equal(F.toarray(), GU.dense(z, "F"))
    np.testing.assert_array_equal(b, z["b"])
    np.testing.assert_array_equal(h, z["h"])


def test_icon_text_roundtrip_and_synthesis(tmp_path):
    p = icon.load1_day01()
    fn = tmp_path / "inst.txt"
    icon.write_icon(p, fn)
    assert icon.data_reading(fn) == p
    s = icon.synthesize(5, 20, seed=3)
    A, b, G, h, F = icon.make_matrix_intopt(**s)
    assert A.shape == (60, 4800) and G.shape == (240, 4800) and F.shape == (4800, 48)
