#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into profiles/<name>.json + .txt.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_<kernel>   [--launches launches.csv]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    name_col = head.index("Kernel Name")
    launches = []
    for r in data:
        rec = {"kernel": r[name_col]}
        for i, c in enumerate(head):
            if c in KEYS and r[i] != "":
                try:
                    rec[c] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    rec[c] = {"value": r[i], "unit": units[i]}
        launches.append(rec)
    summary = {"report": rep, "launches": launches}
    if "--launches" in sys.argv:
        path = sys.argv[sys.argv.index("--launches") + 1]
        per = {}
        with open(path) as f:
            lines = [l for l in f if l.startswith('"')]
        for r in csv.DictReader(lines):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            k = r["Kernel Name"]
            v = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") == "us":
                v *= 1e3
            elif r.get("Metric Unit") == "ms":
                v *= 1e6
            per.setdefault(k, []).append(v)
        tot = sum(sum(v) for v in per.values())
        summary["launch_list"] = {
            "source": path, "total_ns": tot,
            "kernels": [{"kernel": k, "launches": len(v), "mean_ns": sum(v) / len(v),
                         "share": sum(v) / tot} for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1]))]}
    with open(out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    with open(out + ".txt", "w") as f:
        for rec in launches:
            f.write(f"== {rec['kernel']}\n")
            for k in KEYS:
                if k in rec:
                    f.write(f"  {k:90s} {rec[k]['value']} {rec[k]['unit']}\n")
        if "launch_list" in summary:
            f.write("== launch list (ncu --metrics gpu__time_duration.sum; cold-cache, serialised)\n")
            for k in summary["launch_list"]["kernels"]:
                f.write(f"  {k['share']*100:6.2f}%  n={k['launches']:4d}  mean={k['mean_ns']/1e3:9.2f} us  {k['kernel'][:100]}\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
