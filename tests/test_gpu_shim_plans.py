"""The f1 shim's output on the CUDA path (-m gpu): the CbPlans integration/cbgpu_shim.c's translate_plan makes out of
reference-built Plan trees (oracle/ref_plan.c; tests/shim_plans.py) run through cb_ExecInitNode / cb_ExecProcNode on the device
- both kernel routes, and the two-stage Q1 over three in-process segments with its Motions - and return the reference's expected
rows (rpt_tpch).  CPU twin: tests/test_shim_plans.py (same plans through the oracle)."""
import os

import pytest

from cloudberry_b200 import capi, tpch
from gpu_util import shard, to_device
import shim_plans as SP

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(SP.LIB), reason="oracle/_ref/libplan_ref.so travels prebuilt")]


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _run(ctx, t, rels, generic):
    rt = t.range_table(rels, capi.hashbpchar)
    dev = to_device(ctx, rt)
    ex = capi.Executor(ctx, dev, force_generic=generic)
    try:
        return ex.run(t.plan).rows
    finally:
        ex.close()
        for d in dev:
            d.free()


@pytest.mark.parametrize("generic", [False, True])
def test_translated_q1_q3_q5(ctx, golden, generic):
    rels, exp = golden
    L = SP.lib()
    t = SP.Translated(L, "q1", a=1, b=tpch.Q1_CUTOFF)
    assert tpch.format_q1(_run(ctx, t, rels, generic)) == exp["q1"]
    t = SP.Translated(L, "q3", text="MACHINERY", b=tpch.date_to_days(1995, 3, 15))
    assert SP.q3_top10(_run(ctx, t, rels, generic)) == exp["q3"]
    t = SP.Translated(L, "q5", text="AMERICA", b=tpch.date_to_days(1997, 1, 1), c=tpch.date_to_days(1998, 1, 1))
    assert tpch.format_q5(_run(ctx, t, rels, generic), exp["dict"]["n_name_dict"]) == exp["q5"]


def test_translated_two_stage_q1_on_segments(ctx, oracle, golden):
    rels, exp = golden
    nsegs = 3
    t = SP.Translated(SP.lib(), "q1", a=nsegs, b=tpch.Q1_CUTOFF)
    dsegs = [to_device(ctx, t.range_table(s, capi.hashbpchar)) for s in shard(oracle, rels, nsegs)]
    cl = capi.Cluster(ctx, dsegs)
    res = cl.run(t.plan)
    assert tpch.format_q1(res.rows) == exp["q1"]
    cl.close()
    for d in dsegs:
        for r in d:
            r.free()
