"""Capacity semantics: what does not fit the operator's memory runs in passes instead of failing -
multi-batch hybrid hash join (nodeHash.c:980-990, 1133, 2223-2242) and partitioned hash aggregation (nodeAgg.c:2149, 3215) -
and LEFT joins against a build side with duplicate keys (every match, or one NULL-extended row).  CUDA path vs the CPU oracle."""
import numpy as np
import pytest

from cloudberry_b200 import capi, tpch
from cloudberry_b200 import plan as P
from gpu_util import canon, to_device
from test_gpu_edge import agg_over, dim, fact, make, run_both, scan

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("mem_kb", [16, 1024])
def test_q3_q5_with_a_tiny_operator_memory(ctx, oracle, mem_kb):
    """SF0.05: the orders / customer build sides and Q3's ~570 groups exceed a 16 KB budget many times over -> batches and
    partitions; the rows are the oracle's all the same"""
    rels_o = tpch.gen_tables(0.05, oracle.hashbpchar)
    rels_p = tpch.gen_tables(0.05, capi.hashbpchar)
    dev = to_device(ctx, rels_p)
    ex = capi.Executor(ctx, dev, operator_mem_kb=mem_kb)
    ex1 = capi.Executor(ctx, dev)
    q3 = tpch.q3_plan(tpch.SEGMENTS.index("MACHINERY"), 1)
    q5 = tpch.q5_plan(tpch.REGIONS.index("AMERICA"), 1)
    for plan, fmt in ((q3, tpch.format_q3), (q5, lambda r: tpch.format_q5(r, tpch.NATIONS))):
        got = ex.run(plan)
        want = oracle.execute(plan, [rels_o])
        assert fmt(got.rows) == fmt(want.rows) == fmt(ex1.run(plan).rows)
        nb = [v["hashjoin_nbatch"] for v in got.instrument.values()]
        npart = [v["agg_npartitions"] for v in got.instrument.values()]
        if mem_kb == 16:
            assert max(nb) > 1, nb                      # some build side was split
            if plan is q3:
                assert max(npart) > 1, npart            # and the aggregate ran in partitions
    assert (ex.estate.contents.es_hashjoin_batches_run > 0) == (mem_kb == 16)
    ex.close()
    ex1.close()
    for d in dev:
        d.free()


@pytest.mark.parametrize("jointype", [P.JOIN_INNER, P.JOIN_LEFT, P.JOIN_SEMI, P.JOIN_ANTI])
def test_multi_batch_join_types(ctx, oracle, jointype):
    """every N:1 join type through a build side split into batches (unmatched LEFT / ANTI rows must come out exactly once: in
    their own batch's pass)"""
    fo, fp = make(fact, 30011, seed=11, null_frac=0.05, kmax=6000)
    do, dp = make(dim, 4000, seed=12, kmax=6000)
    sf = scan(1, fo, ["k", "amt", "g"])
    sd = scan(2, do, ["dk", "w", "c"])
    h = P.Hash(sd, [P.out_var(sd, 1)])
    targets = [("k", P.out_var(sf, 1)), ("amt", P.out_var(sf, 2)), ("g", P.out_var(sf, 3))]
    if jointype in (P.JOIN_INNER, P.JOIN_LEFT):
        targets += [("w", P.InnerVar(2, P.INT8)), ("c", P.InnerVar(3, P.DICT8))]
    j = P.HashJoin(jointype, sf, h, [P.out_var(sf, 1)], targets)
    names = [t[0] for t in targets]
    aggs = [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None)] + ([("sw", P.AGG_SUM, "w"), ("cw", P.AGG_COUNT, "w")] if "w" in names else [])
    plan = agg_over(j, names, ["g"] + (["c"] if "c" in names else []), aggs)
    want = oracle.execute(plan, [[fo, do]]).rows
    dev = to_device(ctx, [fp, dp])
    ex = capi.Executor(ctx, dev, operator_mem_kb=16)
    got = ex.run(plan)
    assert canon(got.rows) == canon(want)
    assert max(v["hashjoin_nbatch"] for v in got.instrument.values()) >= 4
    ex.close()
    for d in dev:
        d.free()


@pytest.mark.parametrize("generic", [False, True])
def test_left_join_with_duplicate_build_keys(ctx, oracle, generic):
    """N:M LEFT join: an outer row appears once per partner, or once NULL-extended when it has none (a NULL key has none)"""
    fo, fp = make(fact, 20011, seed=21, null_frac=0.1)
    do, dp = make(dim, 90, seed=22, null_frac=0.1, dup=3)
    sf = scan(1, fo, ["k", "v", "amt", "g"])
    sd = scan(2, do, ["dk", "w", "c"])
    h = P.Hash(sd, [P.out_var(sd, 1)])
    j = P.HashJoin(P.JOIN_LEFT, sf, h, [P.out_var(sf, 1)],
                   [("g", P.out_var(sf, 4)), ("v", P.out_var(sf, 2)), ("amt", P.out_var(sf, 3)), ("w", P.InnerVar(2, P.INT8)),
                    ("c", P.InnerVar(3, P.DICT8))])
    plan = agg_over(j, ["g", "v", "amt", "w", "c"], ["g", "c"],
                    [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None), ("cw", P.AGG_COUNT, "w"), ("sw", P.AGG_SUM, "w")])
    got, want = run_both(ctx, oracle, plan, [fo, do], [fp, dp], generic)
    assert canon(got) == canon(want)
    # NULL-extended rows exist (group key c is NULL for them) and so do multiplied ones
    assert any(r[1] is None for r in want)
    inner = P.HashJoin(P.JOIN_INNER, sf, h, [P.out_var(sf, 1)], [("g", P.out_var(sf, 4)), ("amt", P.out_var(sf, 3))])
    n_inner = sum(int(r[-1]) for r in run_both(ctx, oracle, agg_over(inner, ["g", "amt"], ["g"], [("n", P.AGG_COUNT_STAR, None)]), [fo, do], [fp, dp], generic)[1])
    assert sum(int(r[3]) for r in want) > n_inner


@pytest.mark.parametrize("generic", [False, True])
def test_having(ctx, oracle, generic):
    """HAVING over exact aggregate states (TPC-H Q18's `having sum(l_quantity) > 300` shape, plus avg / count / a grouping column
    under AND / OR / NOT): the groups the oracle's un-filtered aggregation yields, filtered with Python Decimals"""
    from decimal import Decimal
    rels_o = tpch.gen_tables(0.02, oracle.hashbpchar)
    rels_p = tpch.gen_tables(0.02, capi.hashbpchar)
    li = rels_o[0]
    sc = P.SeqScan(1, [(n, P.Var(1, li.attno(n), *li.var(n)[1:])) for n in ("l_orderkey", "l_quantity", "l_extendedprice")])
    v = tpch._child_var(sc)
    targets = [("l_orderkey", v("l_orderkey")), ("s", P.Aggref(P.AGG_SUM, v("l_quantity"))), ("a", P.Aggref(P.AGG_AVG, v("l_extendedprice"))),
               ("n", P.Aggref(P.AGG_COUNT_STAR))]
    base = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], targets, num_groups=40000)
    allrows = oracle.execute(base, [rels_o]).rows
    cases = [
        ([P.OpExpr(P.OP_GT, P.Aggref(P.AGG_SUM, v("l_quantity")), P.NumericConst("250"))], lambda r: Decimal(r[1]) > 250),
        ([P.OpExpr(P.OP_LE, P.Aggref(P.AGG_AVG, v("l_extendedprice")), P.NumericConst("20000.50")),
          P.OpExpr(P.OP_GE, P.Aggref(P.AGG_COUNT_STAR), P.Const(P.INT8, 5))], lambda r: Decimal(r[2]) <= Decimal("20000.50") and r[3] >= 5),
        ([P.BoolExpr(P.OR_EXPR, P.OpExpr(P.OP_EQ, P.Aggref(P.AGG_COUNT_STAR), P.Const(P.INT8, 7)),
                     P.BoolExpr(P.NOT_EXPR, P.OpExpr(P.OP_GT, v("l_orderkey"), P.Const(P.INT8, 1000))))],
         lambda r: r[3] == 7 or not (r[0] > 1000)),
    ]
    dev = to_device(ctx, rels_p)
    ex = capi.Executor(ctx, dev, force_generic=generic)
    for quals, keep in cases:
        plan = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], targets, num_groups=40000, quals=quals)
        got = ex.run(plan).rows
        want = [r for r in allrows if keep(r)]
        assert 0 < len(want) < len(allrows)
        assert canon(got) == canon(want)
    ex.close()
    for d in dev:
        d.free()


def _outer_join_plan(jointype, fo, do):
    sf = scan(1, fo, ["k", "v", "amt", "g"])
    sd = scan(2, do, ["dk", "w", "c"])
    h = P.Hash(sd, [P.out_var(sd, 1)])
    return P.HashJoin(jointype, sf, h, [P.out_var(sf, 1)],
                      [("k", P.out_var(sf, 1)), ("g", P.out_var(sf, 4)), ("amt", P.out_var(sf, 3)), ("dk", P.InnerVar(1, P.INT4)),
                       ("w", P.InnerVar(2, P.INT8)), ("c", P.InnerVar(3, P.DICT8))])


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("jointype", [P.JOIN_RIGHT, P.JOIN_FULL])
@pytest.mark.parametrize("nf,nd,dup", [(0, 40, 1), (5000, 0, 1), (300, 40, 1), (20011, 90, 3), (40, 45, 1)])
def test_right_and_full_joins(ctx, oracle, jointype, nf, nd, dup, generic):
    """RIGHT / FULL hash joins (HJ_FILL_INNER_TUPLES, nodeHashjoin.c:676-706): build rows nobody matched - NULL-keyed ones
    included - come back once, NULL-extended on the probe side; FULL also keeps the unmatched probe rows.  Both the joined rows
    themselves and an aggregate over them."""
    fo, fp = make(fact, nf, seed=31, null_frac=0.1, kmax=60)
    do, dp = make(dim, nd, seed=32, null_frac=0.1, dup=dup, kmax=60)
    j = _outer_join_plan(jointype, fo, do)
    got, want = run_both(ctx, oracle, j, [fo, do], [fp, dp], generic)
    assert canon(got) == canon(want)
    if nd > 1 and nf > 0:
        assert any(r[0] is None and r[2] is None for r in want)         # unmatched build rows exist
    if jointype == P.JOIN_FULL and nf:
        assert any(r[4] is None for r in want)                          # and unmatched probe rows
    names = ["k", "g", "amt", "dk", "w", "c"]
    plan = agg_over(j, names, ["g", "c"], [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None), ("ck", P.AGG_COUNT, "k"), ("sw", P.AGG_SUM, "w")])
    got, want = run_both(ctx, oracle, plan, [fo, do], [fp, dp], generic)
    assert canon(got) == canon(want)


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("nf,nd,fnull,dnull", [(5000, 40, 0.1, 0.0), (5000, 40, 0.1, 0.2), (5000, 0, 0.1, 0.0), (0, 40, 0.0, 0.0), (20011, 45, 0.0, 0.0)])
def test_not_in_join(ctx, oracle, nf, nd, fnull, dnull, generic):
    """LASJ_NOTIN (nodeHashjoin.c:371-390, 578-590): a NULL on the build side empties the result, a NULL probe key is dropped
    unless the build side is empty, the rest is an anti join; with NULL-free keys the compiled probe kernel takes it as ANTI"""
    from test_oracle_outer_joins import _notin_plan
    fo, fp = make(fact, nf, seed=43, null_frac=fnull, kmax=60)
    do, dp = make(dim, nd, seed=44, null_frac=dnull, kmax=60)
    j = _notin_plan(fo, do)
    got, want = run_both(ctx, oracle, j, [fo, do], [fp, dp], generic)
    assert canon(got) == canon(want)
    if dnull > 0:
        assert got == []
    plan = agg_over(j, ["k", "amt", "g"], ["g"], [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None)])
    got, want = run_both(ctx, oracle, plan, [fo, do], [fp, dp], generic)
    assert canon(got) == canon(want)
