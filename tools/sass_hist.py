#!/usr/bin/env python
"""Resource usage (cuobjdump --dump-resource-usage) and a SASS opcode histogram for the hot kernels of the built objects
(build/zb_capi.o, build/zb_prims.o).  No GPU needed.  Output: profiles/r02_sass_hot_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ["zb_parse_dp_k", "zb_mf_scan_k", "zb_mf_text_k", "rs_scatter_k", "rs_hist_k", "unit_dist_k", "zb_parse_fix_k", "zb_sweep_k", "zb_hop_k", "zb_stitch_k",
       "zb_cand_k", "zb_split_eval_k", "zb_sub_tables_k"]


def main():
    out = []
    for obj in ("build/zb_capi.o", "build/zb_prims.o"):
        p = os.path.join(ROOT, obj)
        res = subprocess.run(["cuobjdump", "--dump-resource-usage", p], capture_output=True, text=True).stdout.splitlines()
        usage = {}
        for i, line in enumerate(res):
            m = re.match(r"\s*Function (\S+):", line)
            if m and i + 1 < len(res):
                usage[m.group(1)] = res[i + 1].strip()
        sass = subprocess.run(["cuobjdump", "-sass", p], capture_output=True, text=True).stdout
        cur, hist = None, {}
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1); hist[cur] = collections.Counter(); continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m and cur:
                hist[cur][m.group(1)] += 1
        for fn in sorted(hist):
            if not any(h in fn for h in HOT):
                continue
            dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
            h = hist[fn]
            tot = sum(h.values())
            out.append("== %s  (%s)" % (dem[:150], obj))
            out.append("   %s" % usage.get(fn, "?"))
            out.append("   %d SASS instructions; top opcodes: %s" % (tot, ", ".join("%s %d" % kv for kv in h.most_common(14))))
            special = {k: v for k, v in h.items() if k.split(".")[0] in ("UBLKCP", "SYNCS", "LDGSTS", "MATCH", "REDUX", "ATOMS", "ATOMG", "RED", "VOTE", "SHFL", "LDGDEPBAR", "VIMNMX", "VIMNMX3", "LOP3", "SHF")}
            base = collections.Counter()
            for k, v in special.items():
                base[k.split(".")[0]] += v
            out.append("   of note: %s" % ", ".join("%s %d" % kv for kv in sorted(base.items())))
            out.append("")
    dst = os.path.join(ROOT, "profiles", "r02_sass_hot_kernels.txt")
    with open(dst, "w") as f:
        f.write("SASS opcode histograms and resource usage of the hot kernels (sm_100a, nvcc 12.9; tools/sass_hist.py)\n"
                "UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, LDGSTS = cp.async, MATCH/REDUX/VOTE/SHFL = warp collectives\n\n")
        f.write("\n".join(out) + "\n")
    print("wrote", dst, len(out), "lines")


if __name__ == "__main__":
    sys.exit(main())
