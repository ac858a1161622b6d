#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) into the small JSON summary kept under profiles/: per captured launch the
duration, DRAM bytes, L2 / issue statistics and the top stall reasons (ncu -i ... --page raw --csv)."""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_static": "smem_static",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k in hdr:
                v = r[hdr.index(k)]
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
                d[name] = v
                u = units[hdr.index(k)]
                if u and name in ("duration", "dram_bytes_read", "dram_bytes_write"):
                    d[name + "_unit"] = u
        launches.append(d)
    json.dump({"source": rep.split("/")[-1], "command": " ".join(sys.argv[2:]), "launches": launches}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
