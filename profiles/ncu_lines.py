#!/usr/bin/env python
"""Attribute the SASS-level samples of an ncu capture to CUDA source lines (run here, no GPU).

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep mapf_rl_b200/libmapf_b200.so <kernel-substring> [top]

ncu's source page is per SASS instruction; nvdisasm -g gives the line of every instruction of the same
function in the same order, so the two are zipped by instruction index.
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    out = []
    for cub in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
        cur, line, active, inl = None, None, False, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                if active and out:
                    return out
                active = kernel_sub in m.group(1)
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                # "inlined at" chains: keep the outermost line inside the kernel body when present
                m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                inl = (os.path.basename(m2[-1][0]), int(m2[-1][1])) if m2 else None
                continue
            if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
                ins = re.sub(r"/\*[0-9a-f]+\*/", "", l).strip()
                out.append((inl or line, line, ins))
        if active and out:
            return out
    return out


def main():
    rep, so, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # first kernel section only
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] == "Kernel Name":
            break
        body.append(r)
    sl = sass_lines(so, ksub)
    print(f"ncu instructions: {len(body)}   nvdisasm instructions: {len(sl)}")
    n = min(len(body), len(sl))
    si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = tot_e = 0
    reasons = collections.Counter()
    for k in range(n):
        r = body[k]
        key = sl[k][0]
        s, e = int(r[si] or 0), int(r[ei] or 0)
        per_line[key][0] += s
        per_line[key][1] += e
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                per_line[key][2][hdr[c]] += v
                reasons[hdr[c]] += v
        tot_s += s
        tot_e += e
    print(f"total samples {tot_s}, warp instructions {tot_e}")
    print("stall reasons:", ", ".join(f"{k}={v / max(tot_s, 1) * 100:.1f}%" for k, v in reasons.most_common(8)))
    src = {}
    print(f"{'samples%':>8s} {'inst%':>7s}  line  top-stalls")
    for key, (s, e, cnt) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        f, ln = key if key else ("?", 0)
        if f not in src:
            cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(so)), "**", f), recursive=True)
            src[f] = open(cands[0]).read().splitlines() if cands else []
        text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
        st = " ".join(f"{k[6:]}:{v}" for k, v in cnt.most_common(3))
        print(f"{s / max(tot_s, 1) * 100:8.2f} {e / max(tot_e, 1) * 100:7.2f}  {f}:{ln:<4d} [{st}]  {text}")


if __name__ == "__main__":
    main()
