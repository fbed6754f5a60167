#!/usr/bin/env python
"""Per-CUDA-source-line totals (instructions, lane efficiency, stall samples) from an ncu report captured with
--import-source on; uses `ncu --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    res = []
    for r in rows:
        if len(r) > 8 and r[0] == "Line No":
            hdr = {n: k for k, n in enumerate(r)}
            # two "Source" columns: first = CUDA line text
            continue
        if r and r[0] == "Function Name":
            print("==", r[1][:100])
        if hdr is None or len(r) < 10 or not r[0]:
            continue
        try:
            ie = int(r[hdr["Instructions Executed"]]); te = int(r[hdr["Thread Instructions Executed"]]); sm = int(r[hdr["# Samples"]])
        except (ValueError, KeyError):
            continue
        res.append((ie, te, sm, r[0], r[1].strip()[:120]))
    tot = sum(x[0] for x in res) or 1
    tots = sum(x[2] for x in res) or 1
    print("total warp-instructions %d, samples %d" % (tot, tots))
    for ie, te, sm, ln, src in sorted(res, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  L%-5s %s" % (100.0 * ie / tot, 100.0 * sm / tots, te / max(1, ie), ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
