#!/usr/bin/env python
"""Summarises `ncu --page source --csv` output: per SASS instruction the stall samples by reason, and totals per
code region (regions are cut wherever the executed-instruction count changes by > 4x, which separates the warp roles).
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_source_summary.py src.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    def f(n):
        try:
            return float(r[col[n]])
        except ValueError:
            return 0.0
    ins.append(dict(addr=r[col["Address"]], sass=r[col["Source"]].strip(), samples=f("# Samples"), execd=f("Instructions Executed"),
                    thr=f("Avg. Predicated-On Threads Executed"), wf=f("L1 Wavefronts Shared"), stalls={n: f(n) for n in stall_cols}))
tot = sum(i["samples"] for i in ins)
print("instructions", len(ins), "samples", tot)
agg = {}
for i in ins:
    for n, v in i["stalls"].items():
        agg[n] = agg.get(n, 0) + v
print("stall totals:", ", ".join(f"{n[6:]} {v / tot * 100:.1f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
# regions
regions = []
cur = None
for k, i in enumerate(ins):
    e = max(i["execd"], 1.0)
    if cur is None or not (cur["e"] / 4 <= e <= cur["e"] * 4):
        cur = dict(start=k, e=e, samples=0.0, n=0, execd=0.0, stalls={}, wf=0.0)
        regions.append(cur)
    cur["samples"] += i["samples"]; cur["n"] += 1; cur["execd"] += i["execd"]; cur["end"] = k; cur["wf"] += i["wf"]
    for n, v in i["stalls"].items():
        cur["stalls"][n] = cur["stalls"].get(n, 0) + v
print("\nregions with >= 1% of the samples:")
for rg in regions:
    if rg["samples"] < 0.01 * tot:
        continue
    st = ", ".join(f"{n[6:]} {v / rg['samples'] * 100:.0f}%" for n, v in sorted(rg["stalls"].items(), key=lambda kv: -kv[1])[:5] if v > 0)
    print(f"  [{rg['start']:5d}-{rg['end']:5d}] {rg['n']:4d} instr, executed {rg['execd']:.3g} warp-instr, smem wavefronts {rg['wf']:.3g}, samples {rg['samples'] / tot * 100:5.1f}%: {st}")
    print("        first:", ins[rg["start"]]["sass"][:90])
print(f"\ntop {top_n} instructions by samples:")
for i in sorted(ins, key=lambda x: -x["samples"])[:top_n]:
    st = ", ".join(f"{n[6:]} {v:.0f}" for n, v in sorted(i["stalls"].items(), key=lambda kv: -kv[1])[:3] if v > 0)
    k = ins.index(i)
    print(f"  #{k:5d} {i['samples'] / tot * 100:5.2f}%  exec {i['execd']:.3g}  {i['sass'][:70]:70s} {st}")
