#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel, --set full --import-source on) into markdown: headline metrics, stall
reasons per issued instruction, and the source lines with the most stall samples.  Runs where ncu is installed
(no GPU needed):   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_kernel.md"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
    h, units = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % rep)
    for ki, v in enumerate(rows[2:]):
        name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print("## launch %d: `%s`\n" % (ki, name))
        print("| metric | value | unit |\n|---|---|---|")
        for i, n in enumerate(h):
            if n in WANT:
                print("| %s | %s | %s |" % (n, v[i], units[i]))
        stalls = []
        for i, n in enumerate(h):
            if "issue_stalled" in n and n.endswith("_per_issue_active.ratio") and "not_issued" not in n:
                try:
                    stalls.append((float(v[i].replace(",", "")), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        tot = sum(s for s, _ in stalls)
        print("\nWarp cycles per issued instruction: %.2f — by stall reason:\n" % tot)
        print("| reason | cycles | share |\n|---|---|---|")
        for s, n in stalls[:8]:
            print("| %s | %.2f | %.0f %% |" % (n, s, 100 * s / tot))
        print()
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "cuda,sass"))))
    cur, hdr = None, None
    agg = collections.defaultdict(lambda: [0, 0, ""])
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0].isdigit() and hdr:
            try:
                a = agg[(cur, int(r[0]))]
                a[0] += int(r[4])
                a[1] += int(r[hdr.index("Instructions Executed")])
                a[2] = r[1].strip()[:100]
            except (ValueError, IndexError):
                pass
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print("## source lines by warp-stall samples (all launches in the report)\n")
    print("| file:line | stall samples | instructions | source |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        print("| %s:%d | %.1f %% | %.1f %% | `%s` |" % (k[0], k[1], 100 * a[0] / tot, 100 * a[1] / toti, a[2].replace("|", "\\|")))


if __name__ == "__main__":
    main()
