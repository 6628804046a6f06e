#!/usr/bin/env python
"""profiles/rollout_traffic.json from ncu launch metrics of the rollout kernel at bench.py's configuration (K steps, episode
handling on, staggered step counters):  python profiles/tools/r2_traffic.py gpurun_out/traffic_c2.csv c2 20 [...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out_path = os.path.join(ROOT, "profiles", "rollout_traffic.json")
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
args = sys.argv[1:]
for i in range(0, len(args), 3):
    path, cfg, K = args[i], args[i + 1], int(args[i + 2])
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    name, metric, val, unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = {}
    for r in rows[1:]:
        if "rollout_kernel" not in r[name]:
            continue
        launches.setdefault(r[hdr.index("ID")], {})[r[metric]] = (float(r[val].replace(",", "")), r[unit])
    last = launches[sorted(launches, key=int)[-1]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = last["dram__bytes_read.sum"][0] * scale[last["dram__bytes_read.sum"][1]]
    wr = last["dram__bytes_write.sum"][0] * scale[last["dram__bytes_write.sum"][1]]
    dur = last["gpu__time_duration.sum"]
    out[cfg] = {"dram_bytes_per_step": (rd + wr) / K, "dram_read_bytes_per_step": rd / K, "dram_write_bytes_per_step": wr / K,
                "steps_per_launch": K, "ncu_duration": f"{dur[0]} {dur[1]}",
                "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none of rollout_kernel, "
                          f"{cfg}, one launch of {K} steps with episode handling (profiles/tools/r2_ncu_rollout_target.py --reset 1)"}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
