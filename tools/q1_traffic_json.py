#!/usr/bin/env python3
"""tools/q1_traffic_json.py SUMMARY.json -> the `roofline.traffic` record of bench.py (profiles/q1_kernel_traffic.json):
dram__bytes_read.sum + dram__bytes_write.sum of one k_scan_agg_small launch out of an `ncu --set full` summary
(tools/ncu_summary.py), stamped with the sha256 of the kernel source it was measured on - bench.py reports the figure only
while cloudberry_b200/csrc/scan_agg.cu still has that hash, null otherwise."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    d = json.load(open(sys.argv[1]))
    ls = [l for l in d["launches"] if "k_scan_agg_small" in l["kernel"]]
    if not ls:
        sys.exit("no k_scan_agg_small launch in %s" % sys.argv[1])
    l = ls[-1]
    total = l["dram_bytes_read"] * UNIT[l["dram_bytes_read_unit"]] + l["dram_bytes_write"] * UNIT[l["dram_bytes_write_unit"]]
    sha = hashlib.sha256(open(os.path.join(ROOT, "cloudberry_b200", "csrc", "scan_agg.cu"), "rb").read()).hexdigest()
    # ncu prints the instantiation as "void k_scan_agg_small<4, 55, 1, 1>(SmallAggParams)"; bench.py names it "<4,0x37,true,1>"
    import re
    m = re.search(r"k_scan_agg_small<\s*(\d+),\s*(\d+),\s*(\w+),\s*(\d+)>", l["kernel"])
    name = "k_scan_agg_small<%s,0x%x,%s,%s>" % (m.group(1), int(m.group(2)), "true" if m.group(3) in ("1", "true") else "false", m.group(4)) if m else l["kernel"]
    print(json.dumps({"kernel": name, "dram_bytes_per_launch": int(total), "source_sha256": sha,
                      "source": "ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum), %s" % os.path.basename(sys.argv[1])}))


if __name__ == "__main__":
    main()
