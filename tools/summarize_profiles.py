"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01
    python tools/summarize_profiles.py r01 --only cleanup_step_kernel c3_step_kernel   # extra captures only
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")


def launches():
    path = os.path.join(go, f"{tag}_launches.csv")
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        agg.setdefault(row["Kernel Name"].split("(")[0][:90], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    dst = os.path.join(out_dir, f"{tag}_launches.csv")
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k}\",{len(v)},{sum(v):.1f},{sum(v)/len(v):.1f},{sum(v)/tot:.4f}\n")
    return dst


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def raw(name):
    rep = os.path.join(go, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None, None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    dst = os.path.join(out_dir, f"{tag}_{name}_metrics.csv")
    with open(dst, "w") as f:
        f.write("metric,value,unit\n")
        f.write(f"kernel,\"{d.get('Kernel Name', ('?',))[0]}\",\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k},{d[k][0]},{d[k][1]}\n")
    return d, dst


def lines(name, top=40):
    rep = os.path.join(go, f"{tag}_{name}.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    iex, ism = hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg, smp = {}, {}
    for r in rows[hi + 1:]:
        if len(r) > iex and r[0] != "" and r[2] == "-":
            try:
                k = (int(r[0]), r[1][:100])
                agg[k] = agg.get(k, 0) + int(r[iex])
                smp[k] = smp.get(k, 0) + int(r[ism])
            except ValueError:
                pass
    tot, ts = sum(agg.values()), max(1, sum(smp.values()))
    dst = os.path.join(out_dir, f"{tag}_{name}_source_lines.csv")
    with open(dst, "w") as f:
        f.write("pct_warp_instructions,pct_stall_samples,line,source\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
            f.write(f"{100*v/tot:.2f},{100*smp[k]/ts:.2f},{k[0]},\"{k[1].replace(chr(34), chr(39))}\"\n")
    return tot, dst


if "--only" in sys.argv:
    for extra in sys.argv[sys.argv.index("--only") + 1:]:
        print(raw(extra)[1])
        print(lines(extra))
    sys.exit(0)

print(launches())
d, p = raw("step_kernel")
print(p)
tot, p2 = lines("step_kernel")
print(p2, "warp instructions:", tot)
ds, ps = raw("sample_kernel")
print(ps)
# per-shape DRAM traffic (bench.py `roofline.traffic` / `frac_dram`): profiles/traffic.json, keyed by bench.shape_key
SHAPES = {"step_kernel": "treasurehunt:21:1:v2:n1048576:t25:c32:philox",
          "sparse_kernel": "treasurehunt:21:1:v2:n1048576:t25:c32:sparse",
          "c2_kernel": "treasurehunt:21:1:v2:n4096:t100:c32:philox",
          "c3_kernel": "treasurehunt:21:8:v2:n16384:t25:c32:philox",
          "c4_kernel": "treasurehunt:128:64:v4:n2048:t10:c16:philox",
          "cleanup_kernel": "cleanup:21:10:v5:n32768:t25:c8:philox"}
tpath = os.path.join(out_dir, "traffic.json")
t = json.load(open(tpath)) if os.path.exists(tpath) else {"step_kernel_dram_bytes_per_launch": {}, "source": {}}
for name, key in SHAPES.items():
    dd, dst = raw(name)
    if not dd:
        continue
    lines(name)

    def num(k):
        v, u = dd[k]
        return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
    t["step_kernel_dram_bytes_per_launch"][key] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    t["source"][key] = os.path.relpath(dst, ROOT)
    print(name, key, t["step_kernel_dram_bytes_per_launch"][key])
json.dump(t, open(tpath, "w"), indent=1)
