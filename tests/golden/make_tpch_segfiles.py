#!/usr/bin/env python3
"""Write tests/golden/rpt_tpch_segfiles.npz: the reference regression's TPC-H tables (tests/golden/rpt_tpch.npz, from
src/test/regress/data/*.csv) as AOCS column files, written by the REFERENCE's own block writer (oracle/_ref) the way the
rpt_tpch AOCO DDL stores them: compresstype=none, blocksize=32768, checksums on (input/rpt_tpch.source:245-263).
lineitem goes into two segment files (segno 1 and 2), as two concurrent loaders would leave it; the other tables into
segno 1.  Key "<table>/<column>/<segno>" -> the column's file bytes.

Runs only where /root/reference exists (`make -C oracle` first)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import aocs_format as A  # noqa: E402
from cloudberry_b200 import plan as P  # noqa: E402
from cloudberry_b200 import tpch  # noqa: E402

# column -> (writer type, width of character(n))
CHARS = {"c_mktsegment": 10, "n_name": 25, "r_name": 25}


def main():
    if A.ref_lib() is None:
        sys.exit("oracle/_ref/libaocs_ref.so missing")
    d = np.load(os.path.join(ROOT, "tests", "golden", "rpt_tpch.npz"))
    dicts = json.load(open(os.path.join(ROOT, "tests", "golden", "rpt_tpch_expected.json")))["dict"]
    out = {}
    total = 0
    for table, cols in tpch.SCHEMA.items():
        n = len(d[cols[0][0]])
        parts = [(1, 0, n)] if table != "lineitem" else [(1, 0, 33000), (2, 33000, n)]
        for name, typ in cols:
            v = d[name]
            for segno, lo, hi in parts:
                x = v[lo:hi]
                if name in CHARS:
                    texts = dicts[name + "_dict"]
                    raw, nb = A.ref_write_column("bpchars", [texts[c].ljust(CHARS[name]) for c in x], None, True, 32768)
                elif typ == P.NUMERIC:
                    raw, nb = A.ref_write_column("numeric", x, None, True, 32768, dscale=2)
                elif typ == P.BPCHAR1:
                    raw, nb = A.ref_write_column("bpchar", [chr(c) for c in x], None, True, 32768)
                elif typ == P.DATE:
                    raw, nb = A.ref_write_column("date", x, None, True, 32768)
                elif typ == P.INT8:
                    raw, nb = A.ref_write_column("int8", x, None, True, 32768)
                else:
                    raw, nb = A.ref_write_column("int4", x, None, True, 32768)
                out["%s/%s/%d" % (table, name, segno)] = np.frombuffer(raw, dtype=np.uint8)
                total += len(raw)
                print("  %-10s %-16s segno %d  rows %6d  blocks %3d  %8d bytes" % (table, name, segno, hi - lo, nb, len(raw)))
    path = os.path.join(ROOT, "tests", "golden", "rpt_tpch_segfiles.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes for", total, "bytes of column files")


if __name__ == "__main__":
    main()
