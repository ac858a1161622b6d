"""The f1 shim EXERCISED: integration/cbgpu_shim.c's translate_plan on Plan trees built by the reference's own
node constructors in the shapes its planner emits for TPC-H Q1 (one stage, and Gather <- Finalize <- Redistribute <- Partial on
three segments), Q3 and Q5 (oracle/ref_plan.c); the translated CbPlans are executed by the CPU oracle here - and by the CUDA
executor in tests/test_gpu_shim_plans.py - and must return the reference's own expected rows (rpt_tpch)."""
import os

import pytest

from cloudberry_b200 import plan as P
from cloudberry_b200 import tpch
from gpu_util import shard
import shim_plans as SP

pytestmark = pytest.mark.skipif(not os.path.exists(SP.LIB), reason="oracle/_ref/libplan_ref.so is built from /root/reference (make -C oracle ref)")


def test_q1_single_stage(oracle, golden):
    rels, exp = golden
    t = SP.Translated(SP.lib(), "q1", a=1, b=tpch.Q1_CUTOFF)
    assert t.scans == [("lineitem", [5, 6, 7, 8, 9, 10, 11])]          # the projected attributes, ascending
    rt = t.range_table(rels, oracle.hashbpchar)
    assert tpch.format_q1(oracle.execute(t.plan, [rt]).rows) == exp["q1"]


def test_q1_two_stage_with_partial_states(oracle, golden):
    """mark_partial_aggref's leftovers: partial Aggrefs typed bytea, bytea Vars in the Motion's target list and as the Finalize
    Aggrefs' arguments - the shim resolves them to the (N, sum) state types of the Aggrefs they come from"""
    rels, exp = golden
    nsegs = 3
    t = SP.Translated(SP.lib(), "q1", a=nsegs, b=tpch.Q1_CUTOFF)
    top = t.plan
    assert top.type == P.T_Motion
    final = top.lefttree.contents
    redist = final.lefttree.contents
    assert (final.type, redist.type, redist.lefttree.contents.type) == (P.T_Agg, P.T_Motion, P.T_Agg)
    # sum(numeric) states travel as NUMERIC with the input's display scale; count's as int8
    tl = [final.targetlist[i].expr.contents for i in range(final.ntargets)]
    assert [(e.restype, e.dscale) for e in tl[2:]] == [(P.NUMERIC, 2), (P.NUMERIC, 2), (P.NUMERIC, 4), (P.NUMERIC, 6), (P.NUMERIC, 2),
                                                      (P.NUMERIC, 2), (P.NUMERIC, 2), (P.INT8, 0)]
    assert all(e.args[0].contents.restype != 0 for e in tl[2:])
    segs = [t.range_table(s, oracle.hashbpchar) for s in shard(oracle, rels, nsegs)]
    assert tpch.format_q1(oracle.execute(t.plan, segs).rows) == exp["q1"]


def test_q3_with_a_string_literal(oracle, golden):
    rels, exp = golden
    t = SP.Translated(SP.lib(), "q3", text="MACHINERY", b=tpch.date_to_days(1995, 3, 15))
    assert [s[0] for s in t.scans] == ["lineitem", "orders", "customer"]
    assert t.pending == [("customer", 7, "MACHINERY")]
    rt = t.range_table(rels, oracle.hashbpchar)
    assert SP.q3_top10(oracle.execute(t.plan, [rt]).rows) == exp["q3"]


def test_q5_grouped_by_a_string_column(oracle, golden):
    rels, exp = golden
    t = SP.Translated(SP.lib(), "q5", text="AMERICA", b=tpch.date_to_days(1997, 1, 1), c=tpch.date_to_days(1998, 1, 1))
    assert sorted(s[0] for s in t.scans) == ["customer", "lineitem", "nation", "orders", "region", "supplier"]
    rt = t.range_table(rels, oracle.hashbpchar)
    assert tpch.format_q5(oracle.execute(t.plan, [rt]).rows, exp["dict"]["n_name_dict"]) == exp["q5"]


def test_customscan_route_wraps_the_same_plans():
    """route 1 of SURVEY.md 8b: the planner-hook side (wrap_subtrees) on the reference-built plans - one CustomScan on top, the
    original sub-tree in custom_plans, custom_scan_tlist = its target list, INDEX_VAR references of the right types; and the wrapped
    sub-tree is still what translate_plan takes (what BeginCustomScan does).  exprType / exprTypmod / exprCollation are the
    reference's own (nodes/nodeFuncs.c compiled where it lies)."""
    L = SP.lib()
    L.ref_plan_wrap.argtypes = L.ref_plan_translate.argtypes
    assert L.ref_plan_wrap(b"q1", None, 1, tpch.Q1_CUTOFF, 0) == 10
    assert L.ref_plan_wrap(b"q1", None, 3, tpch.Q1_CUTOFF, 0) == 10
    assert L.ref_plan_wrap(b"q3", b"MACHINERY", 0, tpch.date_to_days(1995, 3, 15), 0) == 4
    assert L.ref_plan_wrap(b"q5", b"AMERICA", 0, tpch.date_to_days(1997, 1, 1), tpch.date_to_days(1998, 1, 1)) == 2
