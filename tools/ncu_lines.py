#!/usr/bin/env python
"""Per-CUDA-source-line totals (instructions, lane efficiency, stall samples + the dominant stall reasons) from an ncu report
captured with --import-source on.  usage: ncu_lines.py report.ncu-rep [kernel-regex] [top]"""
import csv
import subprocess
import sys


def main(path, kernel=None, top=40):
    cmd = ["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if kernel:
        cmd += ["--kernel-name", "regex:" + kernel, "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    res = {}
    for r in rows:
        if len(r) > 8 and r[0] == "Line No":
            hdr = {n: k for k, n in enumerate(r)}   # two "Source" columns: the first is the CUDA line text
            stalls = [(n, k) for k, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n]
            continue
        if r and r[0] == "Function Name":
            print("==", r[1][:100])
        if hdr is None or len(r) < 10 or not r[0]:
            continue
        try:
            ie = int(r[hdr["Instructions Executed"]]); te = int(r[hdr["Thread Instructions Executed"]]); sm = int(r[hdr["# Samples"]])
        except (ValueError, KeyError):
            continue
        e = res.setdefault(r[0], [0, 0, 0, r[1].strip()[:110], {}])
        e[0] += ie; e[1] += te; e[2] += sm
        for n, k in stalls:
            try:
                v = int(r[k])
            except ValueError:
                v = 0
            if v:
                e[4][n] = e[4].get(n, 0) + v
    tot = sum(x[0] for x in res.values()) or 1
    tots = sum(x[2] for x in res.values()) or 1
    print("total warp-instructions %d, samples %d" % (tot, tots))
    for ln, (ie, te, sm, src, st) in sorted(res.items(), key=lambda kv: -kv[1][2])[:top]:
        why = " ".join("%s:%d" % (n[6:], v) for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  L%-5s %-110s | %s" % (100.0 * ie / tot, 100.0 * sm / tots, te / max(1, ie), ln, src, why))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
