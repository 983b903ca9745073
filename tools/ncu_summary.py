#!/usr/bin/env python
"""Summaries of an ncu capture for profiles/: (1) a metric x launch table from `--page raw`, (2) DRAM traffic per launch
grouped like bench.py's kernel brackets.   python tools/ncu_summary.py <report.ncu-rep> <table.csv> <traffic.json>"""
import csv, json, subprocess, sys

METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__cycles_active.avg", "smsp__cycles_active.max"]
GROUPS = {"k_bp_count": "k_broadphase", "k_bp_scan": "k_broadphase", "k_bp_emit": "k_broadphase", "k_island_sort": "k_islands",
          "k_slot_counts": "k_islands", "k_solve_team": "k_solve", "k_solve_lanes": "k_solve", "k_solve_quads": "k_solve",
          "k_solve_qteam": "k_solve", "k_joint_equations": "k_equations"}


def short(name):
    n = name.replace("ep::", "").replace("void ", "")
    return n.split("(")[0].replace("(int)", "")


def main():
    rep, table, traffic = sys.argv[1:4]
    worlds = sys.argv[4] if len(sys.argv) > 4 else "8192"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [short(r[ki]) for r in data]
    with open(table, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in data])
    def num(r, m):
        i = hdr.index(m)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    per = {}
    for r, n in zip(data, names):
        base = n.split("<")[0]
        g = GROUPS.get(base, base)
        per[g] = per.get(g, 0.0) + num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    json.dump({"source": f"ncu --set full --clock-control none (tools/profile_round.sh), one settled C3 step at {worlds} worlds, {table}; "
                         "dram__bytes_read.sum + dram__bytes_write.sum per launch",
               "note": "grouped like bench.py's brackets: k_solve = all solver launches of the step, k_broadphase = k_bp_*, "
                       "k_islands = k_slot_counts + k_islands + k_island_sort",
               "bytes_per_launch_%s_worlds" % worlds: per}, open(traffic, "w"), indent=1)
    print(json.dumps(per))


if __name__ == "__main__":
    main()
