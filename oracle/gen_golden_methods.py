"""TEST INFRASTRUCTURE — writes tests/golden/methods.npz: dL/dc of the UNTOUCHED reference (through
oracle/ref_shim.py) for `method` = 1, 2, 3 (the saved-tensor slices of ip_model_whole.py:1467-1482 /
ip_model_alter.py:1482-1496), both variants, on instance 0 of the knapsack and Energy fixtures, and for
`presolve=False` on the Energy instance and `pc=False` on both.
Run here (needs /root/reference):  python -m oracle.gen_golden_methods"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ref_shim  # noqa: E402
from tests import golden_util as GU  # noqa: E402


def main():
    out = {}
    for name in ("knapsack", "energy_load1"):
        z = GU.load(name)
        inp = GU.problem_inputs(name, z, 0)
        case = GU.cases(z)[0]
        for variant in ("whole", "alter"):
            for method in (1, 2, 3):
                kw = dict(case["kw"], method=method)
                x, grad, cap = ref_shim.run_ipofunc(variant, case["c"][0], case["g"][0], **inp, **kw)
                out["%s_%s_m%d_grad" % (name, variant, method)] = grad.astype(np.float32)
                out["%s_%s_m%d_x" % (name, variant, method)] = x.astype(np.float32)
    # presolve=False (ip_model_whole.py:1352): x and dL/dc of the reference on the Energy LP (174 x 1584 standard form)
    z = GU.load("energy_load1")
    inp = GU.problem_inputs("energy_load1", z, 0)
    case = GU.cases(z)[0]
    for variant in ("whole", "alter"):
        x, grad, cap = ref_shim.run_ipofunc(variant, case["c"][0], case["g"][0], **inp, **dict(case["kw"], presolve=False))
        out["energy_load1_%s_nopresolve_grad" % variant] = grad.astype(np.float32)
        out["energy_load1_%s_nopresolve_x" % variant] = x.astype(np.float32)
        out["energy_load1_%s_nopresolve_iters" % variant] = np.int32(cap["iters"])
    # pc=False (:1392-1394, :846-847): one direction per iteration, centring beta * mean(x t)
    for name in ("knapsack", "energy_load1"):
        z = GU.load(name)
        inp = GU.problem_inputs(name, z, 0)
        case = GU.cases(z)[0]
        x, grad, cap = ref_shim.run_ipofunc("whole", case["c"][0], case["g"][0], **inp, **dict(case["kw"], pc=False))
        out["%s_whole_plain_grad" % name] = grad.astype(np.float32)
        out["%s_whole_plain_x" % name] = x.astype(np.float32)
        out["%s_whole_plain_iters" % name] = np.int32(cap["iters"])
    np.savez_compressed(os.path.join(GU.GOLDEN, "methods.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
