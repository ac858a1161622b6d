"""Pin the CPU oracle against the reference's own regression fixture (not gpu).

The expected rows are the reference's src/test/regress/output/rpt_tpch.source (Q1 :334-340,
Q3 :465-477, Q5 :536-543) over src/test/regress/data/*.csv, committed as tests/golden/ by
tests/golden/make_golden.py.
"""
import pytest

from cloudberry_b200 import plan as P
from cloudberry_b200 import tpch


def test_q1_matches_reference_expected(oracle, golden):
    rels, exp = golden
    res = oracle.execute(tpch.q1_plan(1), [rels])
    assert tpch.format_q1(res.rows) == exp["q1"]


def test_q3_matches_reference_expected(oracle, golden):
    rels, exp = golden
    seg = exp["dict"]["c_mktsegment_dict"].index("MACHINERY")
    res = oracle.execute(tpch.q3_plan(seg, 1), [rels])
    assert tpch.format_q3(res.rows) == exp["q3"]


def test_q5_matches_reference_expected(oracle, golden):
    rels, exp = golden
    reg = exp["dict"]["r_name_dict"].index("AMERICA")
    res = oracle.execute(tpch.q5_plan(reg, 1), [rels])
    assert tpch.format_q5(res.rows, exp["dict"]["n_name_dict"]) == exp["q5"]


def _shard(oracle, rels, nsegs):
    """Distribute the range table as the reference does: lineitem/orders by cdbhash(orderkey),
    the rest replicated (rpt_tpch.source DISTRIBUTED BY / REPLICATED)."""
    import ctypes as C
    import numpy as np
    L = oracle.lib()
    segs = [[] for _ in range(nsegs)]
    for rel in rels:
        key = tpch.DIST_KEY[rel.name]
        if key is None:
            for s in range(nsegs):
                segs[s].append(rel)
            continue
        col = rel.columns[rel.attno(key) - 1]
        t = (C.c_int32 * 1)(rel.types[rel.attno(key) - 1])
        dest = np.array([L.ora_cdbhash_segment(t, (C.c_int64 * 1)(int(v)), None, 1, nsegs) for v in col])
        for s in range(nsegs):
            segs[s].append(rel.take(np.nonzero(dest == s)[0]))
    return segs


@pytest.mark.parametrize("nsegs", [2, 3])
def test_two_stage_q1_on_segments(oracle, golden, nsegs):
    """BASELINE config 0: Q1 on CPU segments - partial agg, Redistribute Motion, final agg, gather."""
    rels, exp = golden
    segs = _shard(oracle, rels, nsegs)
    res = oracle.execute(tpch.q1_plan(nsegs), segs, nthreads=nsegs)
    assert tpch.format_q1(res.rows) == exp["q1"]
    assert set(res.segments) == {0}          # gathered on one receiver


@pytest.mark.parametrize("replicated", [True, False])
def test_q3_q5_on_segments(oracle, golden, replicated):
    rels, exp = golden
    nsegs = 3
    segs = _shard(oracle, rels, nsegs)
    if not replicated:
        # customer by c_custkey, supplier by s_suppkey
        import ctypes as C
        import numpy as np
        L = oracle.lib()
        for name, key in (("customer", "c_custkey"), ("supplier", "s_suppkey")):
            i = tpch.RT.index(name)
            rel = rels[i]
            col = rel.columns[rel.attno(key) - 1]
            t = (C.c_int32 * 1)(P.INT4)
            dest = np.array([L.ora_cdbhash_segment(t, (C.c_int64 * 1)(int(v)), None, 1, nsegs) for v in col])
            for s in range(nsegs):
                segs[s][i] = rel.take(np.nonzero(dest == s)[0])
    seg = exp["dict"]["c_mktsegment_dict"].index("MACHINERY")
    res = oracle.execute(tpch.q3_plan(seg, nsegs, customer_replicated=replicated), segs, nthreads=2)
    assert tpch.format_q3(res.rows) == exp["q3"]
    reg = exp["dict"]["r_name_dict"].index("AMERICA")
    res = oracle.execute(tpch.q5_plan(reg, nsegs, replicated=replicated), segs, nthreads=2)
    assert tpch.format_q5(res.rows, exp["dict"]["n_name_dict"]) == exp["q5"]
