#!/usr/bin/env python
"""Turn gpurun_out/prof_TAG.ncu-rep + launches_TAG.csv (tools/gpu_ncu.sh) into the tracked
profiles/TAG_summary.txt, profiles/TAG_launches.csv and profiles/ncu_traffic.json.
Usage: python tools/make_profile_summary.py TAG"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
src = open(os.path.join(ROOT, "gym_auv_b200", "csrc", "auv_kernels.cu")).read().split("\n")


def line_of(pattern):
    for i, l in enumerate(src, 1):
        if re.search(pattern, l):
            return i
    raise KeyError(pattern)


marks = [("fetch", r"^__device__ __forceinline__ void lidar_fetch"), ("cast_ray_record", r"^__device__ __forceinline__ float cast_ray_record"),
         ("env_head", r"^template <bool COUNT>$"), ("init", r"every ray starts at"), ("round", r"---- round:"),
         ("candidates", r"---- candidate rays"), ("stage_vertices", r"---- stage vertices"), ("count_mode", r"if \(COUNT\) \{  // reference"),
         ("cast_loop", r"---- cast: lanes"), ("closeness", r"---- closeness / collision"), ("obs_out", r"closeness part of the observation"),
         ("pooling", r"optional sector pooling"), ("info", r"the navigation part of the observation \(obs"),
         ("reward_done", r"---- reward \(rewarder"), ("auto_reset", r"auto-reset, bulk"), ("kernel_loop", r"lidar_smem_per_warp\(int")]
lines = [(n, line_of(p)) for n, p in marks]
regions = ",".join(f"{n}:{a}-{(lines[i + 1][1] - 1) if i + 1 < len(lines) else a + 45}" for i, (n, a) in enumerate(lines) if n != "count_mode")
for k in ("lidar", "nav"):
    with open(os.path.join(ROOT, "gpurun_out", f"cs_{tag}_{k}.csv"), "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k",
                        "regex:k_lidar" if k == "lidar" else "regex:k_vessel_nav"], stdout=f, stderr=subprocess.DEVNULL)
run = lambda *a: subprocess.run(a, capture_output=True, text=True, cwd=ROOT).stdout
out = [f"# {tag}: ncu --set full --clock-control none, python bench.py --steps 2 --warmup 3 --chunks 1 (65536 envs, 180 rays,",
       f"# 16+16 obstacles); launch list of the same command: profiles/{tag}_launches.csv.  Two launches per step:",
       "# k_vessel_nav<1,1,4> (moving-obstacle update + RK step + projection + navigation + culling) and k_lidar<0>."]
out.append("\n".join(l for l in run("python", "tools/ncu_summary.py", rep).split("\n") if "warp_issue_stalled" not in l))
raw = list(csv.reader(run("ncu", "-i", rep, "--page", "raw", "--csv").splitlines()))
h, units = raw[0], raw[1]
traffic = {"envs": 65536, "rays": 180, "source": f"profiles/{tag}_summary.txt (ncu --set full, one launch each)"}
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
        return float(r[i]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[units[i]]

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    traffic["k_lidar" if "k_lidar" in name else "k_vessel_nav"] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr}
out.append(f"\n## k_lidar regions: share of stall samples / of warp instructions (auv_kernels.cu line ranges {regions})")
reg = run("python", "tools/ncu_lines.py", f"gpurun_out/cs_{tag}_lidar.csv", "12", regions)
out.append("\n".join(l for l in reg[reg.index("--- regions"):reg.index("--- other files")].split("\n") if not l.startswith("---")))
out.append("## k_vessel_nav hot lines")
out.append("\n".join(run("python", "tools/ncu_lines.py", f"gpurun_out/cs_{tag}_nav.csv", "14").split("\n")[:18]))
open(os.path.join(ROOT, "profiles", f"{tag}_summary.txt"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv"))) if r and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(r[4].split("(")[0][:40], []).append(int(r[-1]))
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|auv', python bench.py --steps 2 --warmup 3 --chunks 1 --no-cpu-baseline --no-e2e\n")
    f.write("# (cold-cache, serialised launch times; k_vessel_nav<0,...> and the first launches belong to the reset-cache build and reset(); k_lidar<1> = counting pass)\n")
    f.write("kernel,launches,mean_ns,min_ns,max_ns\n")
    for k, v in agg.items():
        f.write(f"{k},{len(v)},{sum(v) / len(v):.0f},{min(v)},{max(v)}\n")
    f.write("# raw rows\n")
    for r in rows:
        f.write(",".join([r[0], r[4].split("(")[0][:40], r[-1]]) + "\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_summary.txt")).read()[:3000])
