#!/usr/bin/env python3
"""Summarise ncu outputs into small text files for profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv [F]        -> per-kernel launch list summary (F: skip the first F frames,
                                                                             a frame ends with its k_accumulate launch)
  python tools/ncu_summary.py rep gpurun_out/prof_extend.ncu-rep          -> key metrics of a --set full capture"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, skip_frames=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    frames = 0
    for row in csv.DictReader(lines):
        if frames < skip_frames:
            if row["Kernel Name"].startswith("k_accumulate") and row.get("Metric Name", "gpu__time_duration.sum") == "gpu__time_duration.sum":
                frames += 1
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-28s launches=%5d total=%9.3f ms avg=%9.1f us share=%.3f" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, v[1] / tot))


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("## kernel:", name.split("(")[0])
        for i, h in enumerate(hdr):
            if h in WANT or "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                print("%-80s %-16s %s" % (h, units[i], r[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        rep(sys.argv[2])
