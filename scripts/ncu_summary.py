#!/usr/bin/env python
"""Summarise ncu output of a gpurun pass into small text files for profiles/ (read on the CPU box).

  python scripts/ncu_summary.py gpurun_out/<tag> profiles/<name> [--traffic]

Writes <name>_launches.txt (per-kernel launch count, total/avg device time and share from the
`--metrics gpu__time_duration.sum` launch list), <name>_kernels.txt (key metrics of every launch in
the `--set full` capture; one file per report when there are several) and, with --traffic (the capture
of bench.py's own workload), updates profiles/traffic.json (DRAM bytes per launch per kernel).
"""
import csv
import glob
import json
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    update_traffic = "--traffic" in sys.argv[3:]
    os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
    # ---- launch list
    lf = os.path.join(src, "launches.csv")
    if os.path.exists(lf):
        rows = [r for r in csv.reader(l for l in open(lf) if l.startswith('"'))]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = {}
        for r in rows[1:]:
            v = float(r[vi].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)      # -> us
            a = agg.setdefault(r[ki], [0, 0.0])
            a[0] += 1; a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(dst + "_launches.txt", "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare shares)\n")
            f.write("%-28s %8s %12s %10s %7s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
            for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write("%-28s %8d %12.1f %10.2f %6.1f%%\n" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
        print(open(dst + "_launches.txt").read())
    # ---- full capture
    reps = sorted(glob.glob(os.path.join(src, "*.ncu-rep")))
    traffic_path = os.path.join(os.path.dirname(dst) or ".", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        name_i = hdr.index("Kernel Name")
        per = {}
        tag = os.path.basename(rep)[:-len(".ncu-rep")].split("_")[-1]
        kfile = dst + ("_%s" % tag if len(reps) > 1 else "") + "_kernels.txt"
        with open(kfile, "w") as f:
            f.write("# ncu --set full --clock-control none, from %s\n" % os.path.basename(rep))
            for r in rows[2:]:
                f.write("== %s (id %s)\n" % (r[name_i], r[0]))
                stalls = []
                for i, h in enumerate(hdr):
                    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                        try:
                            stalls.append((float(r[i].replace(",", "")), h.split("stalled_")[1]))
                        except ValueError:
                            pass
                for k, label in KEYS:
                    if k in hdr:
                        i = hdr.index(k)
                        f.write("  %-22s %s %s\n" % (label, r[i], units[i]))
                if stalls:
                    t = sum(v for v, _ in stalls) or 1.0
                    f.write("  stalls: " + ", ".join("%s %.0f%%" % (n, 100 * v / t) for v, n in sorted(stalls, reverse=True)[:5]) + "\n")
                if "dram__bytes_read.sum" in hdr:
                    ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                    per.setdefault(r[name_i], []).append(to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]))
        for k, v in per.items():
            traffic[k] = sum(v) / len(v)
        print(open(kfile).read())
    if update_traffic:
        json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
