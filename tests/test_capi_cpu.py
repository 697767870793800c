"""CPU tests of the C-ABI library: it loads, exports every symbol include/intopt_b200.h declares, the
host-only structure analysis works, and compute entry points fail loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from neuripsintopt_b200 import _capi
from tests import golden_util as GU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        _capi.build()
    return _capi.lib()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "intopt_b200.h")).read()
    declared = set(re.findall(r"\b(intopt_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.intopt_version() == 5
    # the profiling entry points live in their own header and their own library (never in the product .so)
    prof = open(os.path.join(ROOT, "include", "intopt_b200_profile.h")).read()
    assert set(re.findall(r"\b(intopt_[a-z0-9_]+)\s*\(", prof)) == set(_capi.PROFILE_EXPORTS)
    for name in _capi.PROFILE_EXPORTS:
        assert not hasattr(lib, name)


def test_dense_problems_are_routed_to_the_lock_step_path(lib):
    # dense A with m >= 128 takes the lock-step dense path (no gather lists, no -5 refusal up to m = 4096)
    rng = np.random.default_rng(0)
    for m, n, path in ((64, 128, 0), (96, 192, 0), (128, 256, 1), (192, 384, 1), (1024, 2048, 1)):
        A = sp.csr_matrix(rng.standard_normal((m, n)))
        info, perm = _capi.analyze(A)
        assert info["path"] == path and (info["m"], info["n"], info["nnz"]) == (m, n, m * n)
        assert info["asm_terms"] == n * m * (m + 1) // 2
        if path:
            assert perm.tolist() == list(range(m))


def test_structure_analysis_on_reference_shapes(lib):
    # Energy 172x1582 nnz 22188 S=217133; Twitter 221x2852 nnz 5570 S=8288; Knapsack S=124
    expect = {"energy_load1": (172, 1582, 22188, 217133), "twitter_sp": (221, 2852, 5570, 8288),
              "knapsack": (32, 62, 93, 124)}
    for name, (m, n, nnz, S) in expect.items():
        z = GU.load(name)
        A = sp.csr_matrix(GU.dense(z, "stdA"))
        info, perm = _capi.analyze(A)
        assert (info["m"], info["n"], info["nnz"], info["asm_terms"]) == (m, n, nnz, S)
        assert sorted(perm.tolist()) == list(range(m))
        mb = (m + 7) // 8
        assert info["envelope"] <= 64 * mb * (mb + 1) // 2 and info["factor_in_smem"] == 1
        assert info["scalar_envelope"] <= info["dense_lower"]
        assert info["smem_bytes"] <= 232448 - 256
    # the block-skyline ordering must beat dense storage clearly on the structured Energy problem
    e, _ = _capi.analyze(sp.csr_matrix(GU.dense(GU.load("energy_load1"), "stdA")))
    assert e["envelope"] < 0.65 * e["dense_lower"] and e["chol_macs"] < 0.4 * 172 ** 3 / 6


def test_malformed_input_is_rejected(lib):
    info = _capi.IntoptInfo()
    rp = np.array([0, 2, 1], dtype=np.int32)
    ci = np.array([0, 1], dtype=np.int32)
    val = np.ones(2)
    rc = lib.intopt_analyze(2, 2, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, ctypes.byref(info), None)
    assert rc != 0 and lib.intopt_last_error()


def test_oversized_problem_is_refused_by_the_gather_list_path(lib, monkeypatch):
    """On the persistent-CTA path dense 1024 x 2048 would need 1.07e9 gather-list terms: refused (-5) by the column-count
    check, in milliseconds (by default such a problem takes the lock-step dense path instead)."""
    monkeypatch.setenv("INTOPT_PATH", "cta")
    m, n = 1024, 2048
    rp = (np.arange(m + 1, dtype=np.int64) * n).astype(np.int32)
    ci = np.tile(np.arange(n, dtype=np.int32), m)
    val = np.ones(m * n)
    info = _capi.IntoptInfo()
    rc = lib.intopt_analyze(m, n, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, ctypes.byref(info), None)
    assert rc == -5 and b"too large" in lib.intopt_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure(lib):
    A = sp.csr_matrix(np.array([[1.0, 1.0, 0.0], [0.0, 1.0, 1.0]]))
    h = ctypes.c_void_p()
    rp, ci, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    rc = lib.intopt_problem_create(ctypes.byref(h), 2, 3, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, None, 0)
    assert rc != 0 and not h.value
    from neuripsintopt_b200.ip_model_whole import IPOfunc
    f = IPOfunc(torch.tensor(A.toarray(), dtype=torch.float32), torch.ones(2), torch.Tensor(), torch.Tensor())
    with pytest.raises(RuntimeError):
        f(torch.ones(3))
