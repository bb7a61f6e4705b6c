"""Turns an .ncu-rep (read with `ncu -i`) into the short per-kernel summary committed under profiles/."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}   (launch id {r[hdr.index('ID')]})")
        for k in KEYS:
            if k in hdr:
                print(f"   {k:95s} {r[hdr.index(k)]} {units[hdr.index(k)]}")


def traffic(path, pairs_per_launch, source):
    """profiles/traffic.json for bench.py's roofline.traffic: DRAM bytes (read + written) per pair and launch of the dominant kernel
    (vsweep_kernel, mean over its launches in the capture = the two passes of a wave)."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    tot, n = 0.0, 0
    for r in rows[2:]:
        if "vsweep_kernel" not in r[hdr.index("Kernel Name")]:
            continue
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(k)]]
            tot += float(r[hdr.index(k)]) * scale
        n += 1
    rec = {"kernel": "vsweep_kernel", "launches_in_capture": n, "pairs_per_launch": pairs_per_launch,
           "dram_bytes_per_pair_per_launch": tot / n / pairs_per_launch, "source": source}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(rec)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":      # summarize.py X.ncu-rep --traffic PAIRS_PER_LAUNCH "source text"
        traffic(sys.argv[1], float(sys.argv[3]), sys.argv[4])
    else:
        main(sys.argv[1])
