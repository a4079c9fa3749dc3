#!/usr/bin/env python
"""Turn one GPU-box visit (gpurun_out/: bench.json, bench_ref.json, launches.csv, prof.ncu-rep) into the tracked
summaries under profiles/.   usage: python tools/make_profile_summary.py r02 [config-name]

Writes profiles/<tag>_bench.json, <tag>_bench_ref.json, <tag>_ncu_launches.{csv,md}, <tag>_ncu_step_kernel.md (metrics, stall
breakdown, executed instructions and samples by kernel phase, hottest source lines) and updates profiles/traffic.json
(DRAM bytes and executed warp instructions per launch at the capture's load: what bench.py prints as roofline.traffic /
roofline.issue when the run's load matches)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
config = sys.argv[2] if len(sys.argv) > 2 else "hangzhou"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "pytsc_b200", "csrc", "tsc_b200.cu")

for src, dst in (("bench.json", f"{tag}_bench.json"), ("bench_ref.json", f"{tag}_bench_ref.json"),
                 ("launches.csv", f"{tag}_ncu_launches.csv")):
    if os.path.exists(os.path.join(G, src)) and os.path.getsize(os.path.join(G, src)) > 0:
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# ---- launch list ----
if os.path.exists(os.path.join(G, "launches.csv")):
    rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if r and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:60]
        if "tsc_step_kernel" in r[4]:
            name = r[4].split("(DevScn")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1]) / 1e6
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_ncu_launches.md"), "w") as f:
        f.write(f"# ncu launch list, `bench.py --steps 20 --warmup 5 --no-cpu-baseline` (gpu__time_duration.sum, --clock-control none), "
                f"first {len(rows)} launches (360 of them are the untimed fast-forward)\n\nPer-launch times are cold-cache and serialised: "
                "the step kernel's SHARE is what counts.\n\n| kernel | launches | total ms | share | mean ms |\n|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0]:.4f} |\n")

# ---- full capture ----
rep = os.path.join(G, "prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
want += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
ix = {h: i for i, h in enumerate(hdr)}
lines = ["| metric | " + " | ".join(f"launch {d[0]}" for d in data) + " | unit |", "|---|" + "---|" * (len(data) + 1)]
vals = {}
for w in want:
    if w in ix:
        vals[w] = [d[ix[w]] for d in data]
        lines.append(f"| {w} | " + " | ".join(vals[w]) + f" | {units[ix[w]]} |")

# ---- per source line -> per kernel phase ----
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
shdr = rows[hi]
six = {n: i for i, n in enumerate(shdr)}
per_line = {}
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].rsplit("/", 1)[-1]
        continue
    if r and r[0].isdigit():
        try:
            key = (cur_file, int(r[0]))
            a = per_line.setdefault(key, [r[1][:120], 0, 0])
            a[1] += int(r[six["# Samples"]]); a[2] += int(r[six["Instructions Executed"]])
        except (ValueError, IndexError):
            pass
MARKERS = [("small helpers, templates, division helper", r"// Small device helpers"),
           ("car-following law (no_collision / car_follow)", r"// ---- A.4 car following"),
           ("stop_before / can_yield / reach_steps", r"__device__ double stop_before_speed"),
           ("cross_claimant / can_pass", r"// Which vehicle does the lane-link"),
           ("finish_vehicle (commit, route walk, mover list)", r"// Commit one vehicle's decision"),
           ("slot compaction", r"// ---- stable compaction of the vehicle slots"),
           ("head look-ahead (A.7)", r"// Leader and gap of a head vehicle"),
           ("tick: prologue, handleWaiting (spawn)", r"__device__ void engine_tick"),
           ("tick: head gathering + look-ahead driver", r"//      warp-locally: look-ahead leader \+ gap"),
           ("tick: decisions (car following, intersection rules)", r"// ---- getAction\.  Every decision"),
           ("tick: cross phase", r"// ---- getAction, cross phase"),
           ("tick: list surgery (leave / enter)", r"// ---- updateLocation"),
           ("pytsc layer helpers (round6, windows, masks, controllers)", r"// pytsc layer: phase program"),
           ("retrieve (lane sums, signal stats, rewards, masks, rows, packet)", r"__device__ void retrieve\("),
           ("kernel: stage in / prologue / stage out", r"// The step kernel"),
           ("host code", r"// MetricsParser.mst: maximum spanning forest")]
src_lines = open(SRC).read().split("\n")
starts = []
for label, pat in MARKERS:
    ln = next((i + 1 for i, l in enumerate(src_lines) if re.search(pat, l)), None)
    if ln:
        starts.append((ln, label))
starts.sort()
phase = collections.OrderedDict((label, [0, 0]) for _, label in starts)
other = collections.OrderedDict()
tot_s = sum(a[1] for a in per_line.values()); tot_i = sum(a[2] for a in per_line.values())
for (f, ln), a in per_line.items():
    if f == "tsc_b200.cu":
        label = None
        for s0, lab in starts:
            if ln >= s0:
                label = lab
        if label:
            phase[label][0] += a[1]; phase[label][1] += a[2]
            continue
    o = other.setdefault("header: " + f, [0, 0]); o[0] += a[1]; o[1] += a[2]
ptab = ["| phase (source region of tsc_b200.cu) | warp-state samples | executed warp instructions |", "|---|---|---|"]
for lab, (s_, i_) in sorted(list(phase.items()) + list(other.items()), key=lambda x: -x[1][0]):
    if s_ or i_:
        ptab.append(f"| {lab} | {100 * s_ / max(tot_s, 1):.1f} % | {100 * i_ / max(tot_i, 1):.1f} % |")
hot = []
for ln, a in sorted(per_line.items(), key=lambda x: -x[1][1])[:25]:
    hot.append(f"{ln[0][:14]:14s}:{ln[1]:5d} {100 * a[1] / max(tot_s, 1):5.1f}% smp {100 * a[2] / max(tot_i, 1):5.1f}% inst  {a[0]}")

bench = {}
for name in ("bench_short.json", "bench.json"):
    try:
        bench = json.load(open(os.path.join(G, name)))
        break
    except Exception:
        pass


def mb(x, unit):
    x = float(x.replace(",", ""))
    return x * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]


tr = [mb(a, units[ix["dram__bytes_read.sum"]]) + mb(b, units[ix["dram__bytes_write.sum"]])
      for a, b in zip(vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"])]
inst = [float(x.replace(",", "")) for x in vals["smsp__inst_executed.sum"]]
V = bench.get("config", {}).get("mean_running_vehicles")
tpath = os.path.join(P, "traffic.json")
try:
    traffic = json.load(open(tpath))
    if "dram_bytes_per_launch" in traffic:      # round-1 file: one flat record
        traffic = {}
except Exception:
    traffic = {}
traffic[config] = {"dram_bytes_per_launch": sum(tr) / len(tr), "inst_executed_per_launch": sum(inst) / len(inst),
                   "mean_running_vehicles": V, "source": f"profiles/{tag}_ncu_step_kernel.md", "launches": tr}
json.dump(traffic, open(tpath, "w"), indent=1)
B = bench.get("roofline", {}).get("units_per_launch", 4096)
alg = bench.get("roofline", {}).get("algorithmic_bytes_per_env_step", 0) * B
with open(os.path.join(P, f"{tag}_ncu_step_kernel.md"), "w") as f:
    f.write(f"# ncu --set full, tsc_step_kernel, one launch at the loaded state ({config}, B = {B}, V = {V}, kernel {bench.get('config', {}).get('kernel')}, "
            f"build {bench.get('config', {}).get('build_hash')})\n\n"
            "`ncu --set full --clock-control none --import-source on -k regex:tsc_step -s 372 -c 1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline` "
            "(launch 372 comes after the 360 untimed fast-forward launches).  Cold-cache, serialised replays: use shares and ratios, not absolute times.\n\n"
            + "\n".join(lines) + "\n\n"
            f"DRAM traffic per launch: {', '.join(f'{t / 1e6:.1f} MB' for t in tr)} (= {sum(tr) / len(tr) / B / 1e3:.1f} KB per replica); "
            f"algorithmic bytes per launch (SURVEY 8d formula at this load) = {alg / 1e6:.0f} MB.  "
            "The fused kernel touches HBM once per env-step, not once per tick, so traffic is BELOW the algorithmic figure.\n\n"
            "## Where the warp-state samples and the executed instructions fall (source page)\n\n" + "\n".join(ptab) + "\n\n"
            "## Hottest source lines\n\n```\n" + "\n".join(hot) + "\n```\n")
print(open(os.path.join(P, f"{tag}_ncu_step_kernel.md")).read()[:5000])
