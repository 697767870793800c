"""SASS opcode histograms of the product objects (run after csrc/build.sh): python profiles/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
units = ["intopt_dense", "intopt_inst_fwd", "intopt_inst_fwd_tmem", "intopt_inst_bwd"]
keep = {"DADD", "DFMA", "DMMA", "DMUL", "SHFL", "LDS", "STS", "LDG", "STG", "LD", "ST", "MUFU", "BAR", "UTMALDG", "UBLKCP",
        "SYNCS", "LDGSTS", "ATOMG", "LDTM", "STTM", "UTCATOMSWS", "FENCE", "LDL", "STL"}
print("# SASS opcode histograms of the product objects (cuobjdump -sass build/obj/*.o, sm_100a, this tree); counts over all "
      "kernel instantiations of the unit\n")
for u in units:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(root, "build", "obj", u + ".o")], capture_output=True, text=True).stdout
    c = collections.Counter()
    for line in txt.splitlines():
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            c[m.group(1).split(".")[0]] += 1
    print("## %s.cu" % u)
    for op, n in sorted(((o, n) for o, n in c.items() if o in keep), key=lambda t: -t[1]):
        print("%7d %s" % (n, op))
    print()
print("# dense path: DMMA = FP64 tensor-core MMA (mma.sync m8n8k4.f64), UTMALDG = TMA tensor load (cp.async.bulk.tensor), "
      "UBLKCP = bulk async copy, SYNCS = mbarrier ops, LDGSTS = cp.async.  persistent-CTA path: DFMA, DMMA for the trailing "
      "updates of the block-skyline Cholesky (one warp per 8x8 block pair), UBLKCP + SYNCS for the cold-start factor prefetch; "
      "intopt_inst_fwd_tmem.cu (the kernels the Energy / Twitter shapes run): LDTM / STTM = tcgen05.ld / tcgen05.st "
      "(element-wise n-vectors in tensor memory), UTCATOMSWS = tcgen05.alloc / dealloc; LD / ST (generic) come from the "
      "non-ALLSMEM variants only (the ALLSMEM kernels' schedule lookups compile to LDS).")
