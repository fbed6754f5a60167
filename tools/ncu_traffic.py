#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch for the kernels of `ncu --set full` reports -> profiles/ncu_traffic.json
(bench.py's roofline.traffic reads it).  usage: ncu_traffic.py out.json report.ncu-rep [...]"""
import csv
import json
import subprocess
import sys

TAGS = {"zb_parse_dp_k": "parse_dp", "zb_mf_scan_k": "mf_scan", "zb_mf_text_k": "mf_text", "rs_scatter_k": "rs_scatter", "rs_hist_k": "rs_hist",
        "zb_parse_fix_k": "parse_repair", "tile_filter_k": "mf_tile_filter", "zb_sweep_k": "path_sweep", "unit_dist_k": "mf_unit_dist", "lcp_pack": "lcp_pack"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(out, reports):
    acc = {}
    for path in reports:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        ir, iw, ik, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
        for r in rows[2:]:
            name = r[ik].split("(")[0].split("<")[0].replace("void ", "")
            tag = TAGS.get(name)
            if not tag:
                continue
            b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
            a = acc.setdefault(tag, [0.0, 0])
            a[0] += b
            a[1] += 1
    json.dump({k: round(v[0] / v[1]) for k, v in acc.items()}, open(out, "w"), indent=1, sort_keys=True)
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
