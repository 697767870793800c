"""Summarise `ncu --page source --print-source cuda,sass --csv` by source line: samples, instructions,
dominant stall reasons.  Usage: python profiles/ncu_source_summary.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
H = None; fpath = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; continue
    if H is None or len(r) < len(H) - 5: continue
    if r[2] != "-": continue  # keep the per-source-line rows (address '-'), skip SASS rows
    try: line = int(r[0])
    except ValueError: continue
    d = dict(zip(H[4:], r[4:]))
    smp = float(d.get("# Samples") or 0); ins = float(d.get("Instructions Executed") or 0)
    stalls = {k: float(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
    key = (fpath, line)
    a = agg.setdefault(key, dict(src=r[1], smp=0.0, ins=0.0, stalls=collections.Counter()))
    a["smp"] += smp; a["ins"] += ins; a["stalls"].update(stalls)
tot = sum(a["smp"] for a in agg.values()); toti = sum(a["ins"] for a in agg.values())
print("total samples %d, instructions %.3g" % (tot, toti))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    st = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(a["smp"], 1)) for k, v in a["stalls"].most_common(3))
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s   [%s]" % (100 * a["smp"] / tot, 100 * a["ins"] / toti, f, l, a["src"].strip()[:90], st))
