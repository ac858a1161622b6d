"""Helper of tests/test_shim_plans.py and tests/test_gpu_shim_plans.py: Plan trees built by the REFERENCE's node constructors
(oracle/ref_plan.c, compiled with nodes/makefuncs.c, list.c ... where they lie) and translated by the f1 shim's translate_plan
(integration/cbgpu_shim.c), plus the range tables such a translated plan reads: per scan, the projected columns of the
reference's table in attribute order - what cbgpu_shim_load_relation would put on the device.  The test plays the loader."""
import ctypes as C
import os

import numpy as np

from cloudberry_b200 import plan as P
from cloudberry_b200 import tpch
from cloudberry_b200.relation import NP_DTYPE, HostRelation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("CB_PLAN_REF_LIB") or os.path.join(ROOT, "oracle", "_ref", "libplan_ref.so")

# the reference's regression schema (src/test/regress/sql/rpt_tpch.sql): range-table index -> table, attribute number -> column
TABLES = {1: "lineitem", 2: "orders", 3: "customer", 4: "supplier", 5: "nation", 6: "region"}
ATTNO = {
    "lineitem": {1: "l_orderkey", 3: "l_suppkey", 5: "l_quantity", 6: "l_extendedprice", 7: "l_discount", 8: "l_tax", 9: "l_returnflag",
                 10: "l_linestatus", 11: "l_shipdate"},
    "orders": {1: "o_orderkey", 2: "o_custkey", 5: "o_orderdate", 8: "o_shippriority"},
    "customer": {1: "c_custkey", 4: "c_nationkey", 7: "c_mktsegment"},
    "supplier": {1: "s_suppkey", 4: "s_nationkey"},
    "nation": {1: "n_nationkey", 2: "n_name", 3: "n_regionkey"},
    "region": {1: "r_regionkey", 2: "r_name"},
}


def lib():
    L = C.CDLL(LIB)
    L.ref_plan_translate.restype = C.c_void_p
    L.ref_plan_translate.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.ref_plan_scan.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    L.ref_plan_pending.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]
    L.ref_plan_resolve_pending.argtypes = [C.c_int, C.c_int64]
    L.ref_plan_error.restype = C.c_char_p
    return L


class Translated:
    """a CbPlan made by the shim out of a reference Plan, and what its scans read"""

    def __init__(self, L, which, text=None, a=0, b=0, c=0):
        addr = L.ref_plan_translate(which.encode(), text.encode() if text else None, a, b, c)
        assert addr, "translate_plan refused the reference's plan: %s" % (L.ref_plan_error() or b"").decode()
        self.plan = P.CbPlan.from_address(addr)
        self.scans = []
        i = 0
        while True:
            rti, att = C.c_int(), (C.c_int * 32)()
            n = L.ref_plan_scan(i, C.byref(rti), att, 32)
            if n < 0:
                break
            self.scans.append((TABLES[rti.value], [att[k] for k in range(n)]))
            i += 1
        self.pending = []
        i = 0
        while True:
            rti, attno, buf = C.c_int(), C.c_int(), C.create_string_buffer(256)
            n = L.ref_plan_pending(i, C.byref(rti), C.byref(attno), buf, 256)
            if n < 0:
                break
            self.pending.append((TABLES[rti.value], attno.value, buf.value.decode()))
            i += 1
        self.L = L

    def range_table(self, rels, hashfn):
        """rels: the golden HostRelations (tpch.RT order).  One relation per scan of the translated plan: the scan's projected
        attributes in ascending attribute order; string columns as CB_DICT32 codes (the shim loader's type for them).  Also
        turns the plan's pending string literals into codes of those dictionaries (blank-trimmed comparison: bpchareq)."""
        by_name = {r.name: r for r in rels}
        out = []
        for table, attnos in self.scans:
            src = by_name[table]
            names = [ATTNO[table][a] for a in attnos]
            cols, types, nulls, texts = [], [], [], []
            for nme in names:
                k = src.names.index(nme)
                t = src.types[k]
                if t == P.DICT8:
                    t = P.DICT32
                types.append(t)
                cols.append(np.ascontiguousarray(src.columns[k]).astype(NP_DTYPE[t]))
                nulls.append(src.nulls[k])
                texts.append(src.dict_texts[k])
            r = HostRelation(table, names, types, cols, [src.dscales[src.names.index(nme)] for nme in names], nulls, None, texts)
            out.append(r.set_dict_hashes(hashfn))
        for i, (table, attno, text) in enumerate(self.pending):
            texts = by_name[table].dict_texts[by_name[table].names.index(ATTNO[table][attno])]
            want = text.rstrip(" ")
            code = next((k for k, s in enumerate(texts) if s.rstrip(" ") == want), -1)
            self.L.ref_plan_resolve_pending(i, code)
        return out


def q3_top10(rows):
    """the Sort / Limit the shim leaves to the CPU executor: revenue descending, o_orderdate; then rpt_tpch's column order"""
    from decimal import Decimal
    rs = sorted(rows, key=lambda r: (-Decimal(r[3]), r[1]))[:10]
    return tpch.format_q3([(r[0], r[3], r[1], r[2]) for r in rs])
