#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source
line: stall samples and warp instructions.  Usage: ncu_lines.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
cur_file = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
        continue
    if r[0] != "" and hdr:
        try:
            out.append((cur_file, int(r[0]), r[1].strip(), int(r[i_s] or 0), int(r[i_i] or 0)))
        except ValueError:
            pass
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print(f"total samples {ts}  total warp-inst {ti}")
print("--- by samples")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[3]:7d} {100*o[3]/ts:5.1f}%  inst {o[4]:10d} {100*o[4]/ti:5.1f}%  {o[0]}:{o[1]}  {o[2][:90]}")

# optional region aggregation: ncu_lines.py file.csv topN "name:lo-hi,name:lo-hi" (auv_kernels.cu lines)
if len(sys.argv) > 3:
    regs = []
    for item in sys.argv[3].split(","):
        nm, rg = item.split(":"); lo, hi = rg.split("-"); regs.append((nm, int(lo), int(hi)))
    agg = {nm: [0, 0] for nm, _, _ in regs}; agg["other-files"] = [0, 0]; agg["unassigned"] = [0, 0]
    for f, ln, src, s, i in out:
        if f != "auv_kernels.cu":
            agg["other-files"][0] += s; agg["other-files"][1] += i; continue
        for nm, lo, hi in regs:
            if lo <= ln <= hi:
                agg[nm][0] += s; agg[nm][1] += i; break
        else:
            agg["unassigned"][0] += s; agg["unassigned"][1] += i
    print("--- regions")
    for nm, (s, i) in agg.items():
        print(f"{nm:24s} samples {100*s/ts:5.1f}%   warp-inst {100*i/ti:5.1f}%")
    byfile = {}
    for f, ln, src, s, i in out:
        if f != "auv_kernels.cu":
            k = f"{f}:{ln}"; byfile.setdefault(k, [0, 0, src]); byfile[k][0] += s; byfile[k][1] += i
    print("--- other files top")
    for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0])[:12]:
        print(f"{k:40s} samples {100*v[0]/ts:5.1f}% inst {100*v[1]/ti:5.1f}%  {v[2][:60]}")
