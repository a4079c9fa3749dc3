#!/usr/bin/env python
"""Turn one GPU-box visit (gpurun_out/: bench.json, bench_ref.json, launches.csv, prof.ncu-rep) into the
tracked summaries under profiles/.   usage: python tools/make_profile_summary.py r01b"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

for src, dst in (("bench.json", f"{tag}_bench.json"), ("bench_ref.json", f"{tag}_bench_ref.json"),
                 ("launches.csv", f"{tag}_ncu_launches.csv")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# ---- launch list ----
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
    f.write(f"# ncu launch list, `bench.py --steps 40 --warmup 5 --no-cpu-baseline` (gpu__time_duration.sum, --clock-control none), "
            f"first {len(rows)} launches\n\nPer-launch times are cold-cache and serialised: the step kernel's SHARE is what counts.\n\n"
            "| kernel | launches | total ms | share | mean ms |\n|---|---|---|---|---|\n")
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
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
ix = {h: i for i, h in enumerate(hdr)}
lines = ["| metric | " + " | ".join(f"launch {d[0]}" for d in data) + " | unit |", "|---|" + "---|" * (len(data) + 1)]
vals = {}
for w in want:
    if w in ix:
        vals[w] = [d[ix[w]] for d in data]
        lines.append(f"| {w} | " + " | ".join(vals[w]) + f" | {units[ix[w]]} |")
src = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout
bench = json.load(open(os.path.join(G, "bench.json")))


def mb(x, unit):
    x = float(x.replace(",", ""))
    return x * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]


tr = [mb(a, units[ix["dram__bytes_read.sum"]]) + mb(b, units[ix["dram__bytes_write.sum"]])
      for a, b in zip(vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"])]
json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "source": f"profiles/{tag}_ncu_step_kernel.md", "launches": tr},
          open(os.path.join(P, "traffic.json"), "w"))
alg = bench["roofline"]["algorithmic_bytes_per_env_step"] * bench["roofline"]["units_per_launch"]
with open(os.path.join(P, f"{tag}_ncu_step_kernel.md"), "w") as f:
    f.write(f"# ncu --set full, tsc_step_kernel, launch(es) from step 450 of `bench.py --steps 452..500 --warmup 5 --no-cpu-baseline` "
            f"(Hangzhou 4x4, B=4096, kernel {bench['config']['kernel']})\n\n"
            "Cold-cache, serialised replays: use shares and ratios, not absolute times.\n\n" + "\n".join(lines) + "\n\n"
            f"DRAM traffic per launch: {', '.join(f'{t / 1e6:.1f} MB' for t in tr)} (= {sum(tr) / len(tr) / 4096 / 1e3:.1f} KB per replica); "
            f"algorithmic bytes per launch (SURVEY 8d formula, V = {bench['config']['mean_running_vehicles']:.0f}) = {alg / 1e6:.0f} MB.  "
            "The fused kernel touches HBM once per env-step, not once per tick, so traffic is BELOW the algorithmic figure.\n\n"
            "Per-source-line sampling (tools/ncu_lines.py):\n\n```\n" + src + "```\n")
print(open(os.path.join(P, f"{tag}_ncu_launches.md")).read())
print(open(os.path.join(P, f"{tag}_ncu_step_kernel.md")).read()[:6000])
