"""Key metrics of an .ncu-rep (one captured kernel):  python tools/ncu_metrics.py REPORT [REPORT ...]"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
STALLS = ["math_pipe_throttle", "short_scoreboard", "long_scoreboard", "wait", "no_instruction", "not_selected", "barrier",
          "branch_resolving", "dispatch_stall", "mio_throttle", "lg_throttle", "membar", "imc_miss"]
def metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    out = [("kernel", d.get("Kernel Name", ("?",))[0], "")]
    out += [(k, d[k][0], d[k][1]) for k in KEYS if k in d]
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in d:
            out.append((k, d[k][0], d[k][1]))
    return out
if __name__ == "__main__":
    for rep in sys.argv[1:]:
        print("#", rep)
        for k, v, u in metrics(rep):
            print(f"{k},{v},{u}")
