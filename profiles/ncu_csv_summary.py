#!/usr/bin/env python
"""Readable table from an `ncu -i X.ncu-rep --page raw --csv` export: one block per launch with the metrics the north star
asks for (duration, DRAM bytes and throughput, L1 / L2 / SM throughput, issue utilisation, occupancy, shared-memory
wavefronts and bank conflicts, top stall reasons).   python profiles/ncu_csv_summary.py raw.csv [kernel-substring ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
filt = sys.argv[2:]
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("launch__waves_per_multiprocessor", "waves / SM"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?")
    if filt and not any(f in name for f in filt):
        continue
    print("== " + name[:150])
    for k, label in KEYS:
        if k in d and d[k] != "":
            print(f"   {label:32s} {d[k]:>18s} {units[hdr.index(k)]}")
    stalls = [(float(d[k]), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for k in hdr
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d.get(k) not in ("", None)]
    stalls.sort(reverse=True)
    print("   stalled warps per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))
