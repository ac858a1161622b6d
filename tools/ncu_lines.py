#!/usr/bin/env python3
"""Attribute an ncu capture's per-instruction samples to source lines (ncu's own CUDA source view needs the sources at
their build path; this works from the SASS page + nvdisasm's line table of the same binary).

    ncu -i REP --page source --csv --kernel-name K --launch-skip N --launch-count 1 > sass.csv
    python tools/ncu_lines.py sass.csv cloudberry_b200/libcbgpu.so _Z13k_probe_chain8PcParams [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(so, func):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if (".text.%s:" % func) not in out:
            continue
        tab, cur, inside = {}, None, False
        for l in out.split("\n"):
            if l.startswith(".text."):
                inside = l.startswith(".text.%s:" % func)
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
            if m:
                tab[int(m.group(1), 16)] = cur
        return tab
    return {}


def main():
    sass, so, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tab = line_table(so, func)
    rows = list(csv.reader(open(sass)))
    hdr = next(r for r in rows if "# Samples" in r)
    col = {n: hdr.index(n) for n in ("# Samples", "Instructions Executed", "L2 Theoretical Sectors Global", "stall_long_sb", "stall_barrier")}
    agg, tot, a0 = {}, [0] * 5, None
    for r in rows[rows.index(hdr) + 1:]:
        try:
            addr = int(r[0], 16)
        except ValueError:
            continue
        a0 = addr if a0 is None else a0
        ln = tab.get(addr - a0)
        vals = [int(r[c] or 0) for c in col.values()]
        a = agg.setdefault(ln, [0] * 5)
        for i, v in enumerate(vals):
            a[i] += v
            tot[i] += v
    src = {}
    print("%d instructions mapped; totals: samples %d, warp instructions %d, L2 sectors %d" % (len(tab), tot[0], tot[1], tot[2]))
    print(" samples   instr  l2sect  longsb barrier | line")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if ln:
            path = os.path.join(ROOT, "cloudberry_b200", "csrc", ln[0])
            if ln[0] not in src and os.path.exists(path):
                src[ln[0]] = open(path).read().split("\n")
            if ln[0] in src and ln[1] <= len(src[ln[0]]):
                text = src[ln[0]][ln[1] - 1].strip()[:100]
        print("%6.1f%% %6.1f%% %6.1f%% %6.1f%% %6.1f%% | %s:%s %s" % tuple([100.0 * a[i] / max(tot[i], 1) for i in range(5)] + [ln[0] if ln else "?", ln[1] if ln else 0, text]))


if __name__ == "__main__":
    main()
