#!/usr/bin/env python
"""Turn gpurun_out/prof_TAG.ncu-rep + launches_TAG.csv (tools/gpu_ncu_r2.sh) into the tracked
profiles/TAG_summary.txt, profiles/TAG_launches.csv and profiles/ncu_traffic.json.
Usage: python tools/make_profile_summary.py TAG"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
run = lambda *a: subprocess.run(a, capture_output=True, text=True, cwd=ROOT).stdout
KERNELS = ("k_vessel_nav", "k_nav_cull", "k_lidar")

out = [f"# {tag}: ncu --set full --clock-control none --import-source on (tools/gpu_ncu_r2.sh): the three step kernels of",
       "# BASELINE config 3 (65536 envs x 180 rays x 16+16 obstacles, 1024-path bank) in STEADY STATE, the 1500th step after reset()",
       "# (vessels spread along their paths, episodes desynchronised).  Per step: k_vessel_nav<1,1,4> (obstacle counter, RKF45",
       "# step, path projection), k_nav_cull<4> (navigation features, reward base, culling -> obstacle records), k_lidar<0,0>",
       "# (ray casting, closeness, reward, done, auto-reset).  Launch times under ncu are cold-cache and serialised; the",
       f"# CUDA-event times of the running step are in profiles/{tag}_bench.json (roofline.kernel_ms)."]
out.append("\n".join(l for l in run("python", "tools/ncu_summary.py", rep).split("\n") if "warp_issue_stalled" not in l))
raw = list(csv.reader(run("ncu", "-i", rep, "--page", "raw", "--csv").splitlines()))
h, units = raw[0], raw[1]
traffic = {"envs": 65536, "rays": 180, "source": f"profiles/{tag}_summary.txt (ncu --set full, one launch each, steady state)"}
for r in raw[2:]:
    name = r[h.index("Kernel Name")]
    out.append(f"## stall reasons (warps per issue-active cycle): {name[:40]}")
    st = []
    for i, k in enumerate(h):
        if "issue_stalled" in k and "per_issue_active" in k:
            try:
                st.append((float(r[i]), k))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True)[:8]:
        out.append("   %.3f %s" % (v, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))

    def val(k):
        i = h.index(k)
        return float(r[i]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(units[i], 1)

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    key = next(k for k in KERNELS if k in name)
    traffic[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                    "inst_executed": val("smsp__inst_executed.sum") if "smsp__inst_executed.sum" in h else None}
# (one entry per kernel: bench.py times every launch of the step by its own CUDA-event interval)
for k in KERNELS:
    f = os.path.join(ROOT, "gpurun_out", f"cs_{tag}_{k}.csv")
    with open(f, "w") as fh:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{k}"],
                       stdout=fh, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(open(f)))
    cur, hdr = None, None
    inst, samp = collections.Counter(), collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif hdr and r[0].isdigit():
            try:
                inst[cur] += int(r[i_i] or 0)
                samp[cur] += int(r[i_s] or 0)
            except ValueError:
                pass
    ti, ts = max(1, sum(inst.values())), max(1, sum(samp.values()))
    out.append(f"\n## {k}: share of warp instructions / stall samples per source file")
    for name, v in inst.most_common(8):
        out.append(f"   {name:30s} inst {100 * v / ti:5.1f}%   samples {100 * samp[name] / ts:5.1f}%")
    out.append(f"## {k}: hot lines")
    out.append("\n".join(run("python", "tools/ncu_lines.py", f"gpurun_out/cs_{tag}_{k}.csv", "12").split("\n")[:15]))
open(os.path.join(ROOT, "profiles", f"{tag}_summary.txt"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
lf = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lf):
    rows = [r for r in csv.reader(open(lf)) if r and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split("(")[0][:48], []).append(int(r[-1]))
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|auv' -c 400, python bench.py --steps 2 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e --preroll-steps 40\n")
        f.write("# (cold-cache, serialised launch times: the kernels' SHARES of a step are comparable with the CUDA-event times, not the absolutes)\n")
        f.write("kernel,launches,mean_ns,min_ns,max_ns\n")
        for k, v in agg.items():
            f.write(f"{k},{len(v)},{sum(v) / len(v):.0f},{min(v)},{max(v)}\n")
        f.write("# raw rows: launch id, kernel, ns\n")
        for r in rows:
            f.write(",".join([r[0], r[4].split("(")[0][:48], r[-1]]) + "\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_summary.txt")).read()[:6000])
