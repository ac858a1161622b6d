#!/usr/bin/env python3
"""Build the committed golden fixtures from the reference's own regression data.

Reads (read-only) the TPC-H CSVs the reference's `rpt_tpch` regression test loads
(/root/reference/src/test/regress/input/rpt_tpch.source:98-107) and the expected result
rows of Q1/Q3/Q5 from /root/reference/src/test/regress/output/rpt_tpch.source
(:334-340, :465-477, :536-543), and writes

  tests/golden/rpt_tpch.npz          - only the columns the scan->join->agg path touches,
                                       in the device column encoding of DESIGN.md
                                       (numeric(15,2) -> int64 x100, date -> int32 days since
                                       2000-01-01, char(1) -> uint8, text -> dictionary code)
  tests/golden/rpt_tpch_expected.json - the expected result rows, as text, exactly as the
                                       reference prints them.

This script needs /root/reference and therefore only runs in the build container; the two
output files are committed so the GPU box (which has no /root/reference) can use them.
"""
import datetime
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference/src/test/regress"
HERE = os.path.dirname(os.path.abspath(__file__))
EPOCH = datetime.date(2000, 1, 1)


def rows(*names):
    for n in names:
        with open(os.path.join(REF, "data", n), "r", encoding="latin-1") as f:
            for line in f:
                line = line.rstrip("\n")
                if line:
                    yield line.split("|")


def dec2(s):
    """numeric(15,2) text -> exact int64 scaled by 100."""
    neg = s.startswith("-")
    if neg:
        s = s[1:]
    if "." in s:
        a, b = s.split(".")
    else:
        a, b = s, ""
    b = (b + "00")[:2]
    v = int(a) * 100 + int(b)
    return -v if neg else v


def date(s):
    y, m, d = map(int, s.split("-"))
    return (datetime.date(y, m, d) - EPOCH).days


def dictionary(values):
    names = sorted(set(values))
    code = {n: i for i, n in enumerate(names)}
    return np.array([code[v] for v in values], dtype=np.uint8), names


def main():
    out = {}
    meta = {}
    # load order follows the \copy order of the regression test (small file first)
    li = list(rows("lineitem_small.csv", "lineitem.csv"))
    out["l_orderkey"] = np.array([int(r[0]) for r in li], dtype=np.int64)
    out["l_suppkey"] = np.array([int(r[2]) for r in li], dtype=np.int32)
    out["l_quantity"] = np.array([dec2(r[4]) for r in li], dtype=np.int64)
    out["l_extendedprice"] = np.array([dec2(r[5]) for r in li], dtype=np.int64)
    out["l_discount"] = np.array([dec2(r[6]) for r in li], dtype=np.int64)
    out["l_tax"] = np.array([dec2(r[7]) for r in li], dtype=np.int64)
    out["l_returnflag"] = np.array([ord(r[8]) for r in li], dtype=np.uint8)
    out["l_linestatus"] = np.array([ord(r[9]) for r in li], dtype=np.uint8)
    out["l_shipdate"] = np.array([date(r[10]) for r in li], dtype=np.int32)

    od = list(rows("order_small.csv", "order.csv"))
    out["o_orderkey"] = np.array([int(r[0]) for r in od], dtype=np.int64)
    out["o_custkey"] = np.array([int(r[1]) for r in od], dtype=np.int32)
    out["o_orderdate"] = np.array([date(r[4]) for r in od], dtype=np.int32)
    out["o_shippriority"] = np.array([int(r[7]) for r in od], dtype=np.int32)

    cu = list(rows("customer.csv"))
    out["c_custkey"] = np.array([int(r[0]) for r in cu], dtype=np.int32)
    out["c_nationkey"] = np.array([int(r[3]) for r in cu], dtype=np.int32)
    out["c_mktsegment"], meta["c_mktsegment_dict"] = dictionary([r[6] for r in cu])

    su = list(rows("supplier.csv"))
    out["s_suppkey"] = np.array([int(r[0]) for r in su], dtype=np.int32)
    out["s_nationkey"] = np.array([int(r[3]) for r in su], dtype=np.int32)

    na = list(rows("nation.csv"))
    out["n_nationkey"] = np.array([int(r[0]) for r in na], dtype=np.int32)
    out["n_regionkey"] = np.array([int(r[2]) for r in na], dtype=np.int32)
    out["n_name"], meta["n_name_dict"] = dictionary([r[1] for r in na])

    re_ = list(rows("region.csv"))
    out["r_regionkey"] = np.array([int(r[0]) for r in re_], dtype=np.int32)
    out["r_name"], meta["r_name_dict"] = dictionary([r[1] for r in re_])

    np.savez_compressed(os.path.join(HERE, "rpt_tpch.npz"), **out)

    # expected rows: parse the psql table that follows each query in the expected output
    exp = open(os.path.join(REF, "output", "rpt_tpch.source"), encoding="latin-1").read().split("\n")

    def table_after(marker, start=0):
        """rows of the first psql result table whose rows begin with `marker`."""
        got = []
        for i in range(start, len(exp)):
            if exp[i].startswith(" " + marker + " "):
                got.append([c.strip() for c in exp[i].split("|")][1:])
            elif got:
                break
        return got

    expected = {
        "source": "src/test/regress/output/rpt_tpch.source (heap copy; AO and AOCO copies are identical)",
        "q1": table_after("mpph1"),
        "q3": table_after("mpph3"),
        "q5": table_after("mpph5"),
        "q1_shipdate_le": "1998-08-15",  # date '1998-12-01' - interval '108 day'
        "dict": meta,
        "rows": {k: int(v.shape[0]) for k, v in out.items()},
    }
    assert len(expected["q1"]) == 4 and len(expected["q3"]) == 10 and len(expected["q5"]) == 5
    with open(os.path.join(HERE, "rpt_tpch_expected.json"), "w") as f:
        json.dump(expected, f, indent=1)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
