#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU needed).

    python profiles/ncu_summary.py launches gpurun_out/launches.csv        # per-kernel totals of a launch list
    python profiles/ncu_summary.py raw gpurun_out/prof.ncu-rep              # key metrics of a --set full capture
    python profiles/ncu_summary.py source gpurun_out/prof.ncu-rep [N]       # top-N source lines by stall samples
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        agg[row["Kernel Name"]].append(float(row["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':80s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:80]:80s} {len(v):5d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.2f} {sum(v) / tot:6.3f}")


def raw(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("kernels:", [r[ki][:60] for r in rows[2:]])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {units[i]:12s} {[r[i] for r in rows[2:]]}")
    print("-- warp stall reasons (pct of warp-active, first launch) --")
    st = [(float(rows[2][i].replace(',', '')), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__warp_issue_stalled") and h.endswith("_per_warp_active.pct") and rows[2][i]]
    for v, h in sorted(st, reverse=True)[:10]:
        print(f"   {v:7.2f}  {h}")


def source(rep, top=40):
    rows = ncu_csv(rep, "source")
    hdr = rows[0]
    print(hdr)
    # find the sampling column
    cands = [i for i, h in enumerate(hdr) if "Sampling" in h and "All" in h] or [i for i, h in enumerate(hdr) if "Samples" in h]
    si = cands[0]
    src = hdr.index("Source") if "Source" in hdr else 1
    body = []
    for r in rows[1:]:
        try:
            body.append((float(r[si] or 0), r))
        except (ValueError, IndexError):
            pass
    tot = sum(v for v, _ in body) or 1
    for v, r in sorted(body, key=lambda t: -t[0])[:top]:
        print(f"{v / tot * 100:6.2f}%  {r[0]:>6s}  {r[src][:130]}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "raw":
        raw(sys.argv[2])
    else:
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
