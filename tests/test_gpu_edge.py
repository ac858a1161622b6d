"""Edge cases of the scan / join / aggregate path, CUDA executor vs the CPU oracle, exact (-m gpu).

What the reference's own regression tests exercise around these operators (src/test/regress/sql/join.sql,
aggregates.sql, appendonly / uao visibility tests): empty inputs on either side, ragged row counts around the
kernels' tile sizes, NULL join keys (never match: strict hash operators, nodeHash.c:2161) and NULL group keys
(group together, execGrouping.c:548), invisible rows (visimap), build-side duplicates (N:M), LEFT / SEMI / ANTI
joins, count/min/max/avg over NULLs, float8 sums (tolerance), and arithmetic that leaves 64 bits
(refused loudly, never wrapped)."""
import math

import numpy as np
import pytest

from cloudberry_b200 import capi
from cloudberry_b200 import plan as P
from cloudberry_b200.relation import HostRelation
from gpu_util import canon, to_device

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _nulls(rng, n, frac):
    return (rng.random(n) < frac).astype(np.uint8) if frac else None


def fact(n, seed=1, null_frac=0.0, visible_frac=1.0, kmax=50):
    rng = np.random.default_rng(seed)
    cols = [rng.integers(0, kmax, n), rng.integers(-10**6, 10**6, n), rng.integers(0, 10**9, n), rng.random(n) * 1000.0,
            rng.integers(0, 7, n), rng.integers(0, 3000, n)]
    nulls = [_nulls(rng, n, null_frac), _nulls(rng, n, null_frac), None, None, _nulls(rng, n, null_frac), None]
    vis = None
    if visible_frac < 1.0:
        bits = (rng.random(n) < visible_frac).astype(np.uint8)
        vis = np.packbits(bits, bitorder="little")
    return HostRelation("fact", ["k", "v", "amt", "x", "g", "d"], [P.INT4, P.INT8, P.NUMERIC, P.FLOAT8, P.DICT8, P.DATE], cols,
                        nulls=nulls, visimap=vis, dict_texts=[None, None, None, None, ["g%d" % i for i in range(7)], None])


def dim(n, seed=2, null_frac=0.0, dup=1, kmax=50):
    rng = np.random.default_rng(seed)
    keys = np.repeat(rng.permutation(kmax)[:max(n // dup, 0)], dup)[:n] if n else np.zeros(0, dtype=np.int64)
    n = len(keys)
    cols = [keys, rng.integers(0, 100, n), rng.integers(0, 5, n)]
    return HostRelation("dim", ["dk", "w", "c"], [P.INT4, P.INT8, P.DICT8], cols, nulls=[_nulls(rng, n, null_frac), None, None],
                        dict_texts=[None, None, ["c%d" % i for i in range(5)]])


def scan(relid, rel, names, quals=()):
    tl = []
    for nme in names:
        a, t, ds = rel.var(nme)
        tl.append((nme, P.Var(relid, a, t, ds)))
    return P.SeqScan(relid, tl, quals)


def run_both(ctx, oracle, plan, rels_o, rels_p, generic):
    dev = to_device(ctx, rels_p)
    ex = capi.Executor(ctx, dev, force_generic=generic)
    try:
        got = ex.run(plan)
    finally:
        ex.close()
        for d in dev:
            d.free()
    want = oracle.execute(plan, [rels_o])
    return got.rows, want.rows


def make(fn, *a, **k):
    """the same table twice: dictionary hashes through the oracle's and through the product's hashbpchar"""
    from oracle import oracle as O
    return fn(*a, **k).set_dict_hashes(O.hashbpchar), fn(*a, **k).set_dict_hashes(capi.hashbpchar)


def agg_over(child, names, keys, aggs):
    from cloudberry_b200.tpch import _child_var
    v = _child_var(child)
    targets = [(k, v(k)) for k in keys] + [(n, P.Aggref(op, None if arg is None else v(arg))) for n, op, arg in aggs]
    return P.Agg(child, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [names.index(k) + 1 for k in keys], targets, num_groups=64)


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("n", [0, 1, 31, 33, 255, 257, 2047, 2049, 4097, 100003])
def test_ragged_and_empty_scan_agg(ctx, oracle, n, generic):
    fo, fp = make(fact, n)
    names = ["g", "amt", "v", "d"]
    sc = scan(1, fo, names, [P.OpExpr(P.OP_LT, P.Var(1, fo.attno("d"), P.DATE), P.Const(P.DATE, 2000))])
    plan = agg_over(sc, names, ["g"], [("s", P.AGG_SUM, "amt"), ("c", P.AGG_COUNT_STAR, None), ("a", P.AGG_AVG, "amt"),
                                       ("mn", P.AGG_MIN, "v"), ("mx", P.AGG_MAX, "v")])
    got, want = run_both(ctx, oracle, plan, [fo], [fp], generic)
    assert canon(got) == canon(want)
    assert len(want) == 0 if n == 0 else True


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("jointype", [P.JOIN_INNER, P.JOIN_LEFT, P.JOIN_SEMI, P.JOIN_ANTI])
@pytest.mark.parametrize("nf,nd", [(0, 40), (5000, 0), (5000, 40), (2049, 1)])
def test_join_types_and_empty_sides(ctx, oracle, jointype, nf, nd, generic):
    fo, fp = make(fact, nf, seed=3)
    do, dp = make(dim, nd, seed=4)
    sf = scan(1, fo, ["k", "amt", "g"])
    sd = scan(2, do, ["dk", "w", "c"])
    h = P.Hash(sd, [P.out_var(sd, 1)])
    targets = [("k", P.out_var(sf, 1)), ("amt", P.out_var(sf, 2)), ("g", P.out_var(sf, 3))]
    if jointype in (P.JOIN_INNER, P.JOIN_LEFT):
        targets += [("w", P.InnerVar(2, P.INT8)), ("c", P.InnerVar(3, P.DICT8))]
    j = P.HashJoin(jointype, sf, h, [P.out_var(sf, 1)], targets)
    names = [t[0] for t in targets]
    aggs = [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None)]
    if "w" in names:
        aggs += [("sw", P.AGG_SUM, "w"), ("cw", P.AGG_COUNT, "w")]
    plan = agg_over(j, names, ["g"] + (["c"] if "c" in names else []), aggs)
    got, want = run_both(ctx, oracle, plan, [fo, do], [fp, dp], generic)
    assert canon(got) == canon(want)


@pytest.mark.parametrize("generic", [False, True])
def test_null_keys_visimap_and_duplicates(ctx, oracle, generic):
    """NULL join keys match nothing, NULL group keys form one group, invisible rows do not exist, a build side with
    duplicate keys multiplies the matches (N:M)."""
    fo, fp = make(fact, 20011, seed=5, null_frac=0.1, visible_frac=0.8)
    do, dp = make(dim, 90, seed=6, null_frac=0.1, dup=3)
    sf = scan(1, fo, ["k", "v", "amt", "g"])
    sd = scan(2, do, ["dk", "w", "c"])
    h = P.Hash(sd, [P.out_var(sd, 1)])
    j = P.HashJoin(P.JOIN_INNER, sf, h, [P.out_var(sf, 1)],
                   [("g", P.out_var(sf, 4)), ("v", P.out_var(sf, 2)), ("amt", P.out_var(sf, 3)), ("w", P.InnerVar(2, P.INT8)),
                    ("c", P.InnerVar(3, P.DICT8))])
    plan = agg_over(j, ["g", "v", "amt", "w", "c"], ["g", "c"],
                    [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None), ("cv", P.AGG_COUNT, "v"), ("av", P.AGG_AVG, "v"),
                     ("mn", P.AGG_MIN, "v"), ("mx", P.AGG_MAX, "w")])
    got, want = run_both(ctx, oracle, plan, [fo, do], [fp, dp], generic)
    assert len(want) > 5
    assert canon(got) == canon(want)
    # the NULL group exists on both sides
    assert any(r[0] is None for r in want) and any(r[0] is None for r in got)


@pytest.mark.parametrize("generic", [False, True])
def test_float8_sum_avg_tolerance(ctx, oracle, generic):
    """float8 SUM / AVG accumulate in a different order than float8pl's sequential sum: relative tolerance
    1e-12 * sqrt(N); group keys and counts stay exact."""
    n = 200003
    fo, fp = make(fact, n, seed=7)
    sc = scan(1, fo, ["g", "x"])
    plan = agg_over(sc, ["g", "x"], ["g"], [("sx", P.AGG_SUM, "x"), ("ax", P.AGG_AVG, "x"), ("n", P.AGG_COUNT_STAR, None)])
    got, want = run_both(ctx, oracle, plan, [fo], [fp], generic)
    g = {r[0]: r for r in got}
    w = {r[0]: r for r in want}
    assert set(g) == set(w)
    tol = 1e-12 * math.sqrt(n)
    for k in w:
        assert g[k][3] == w[k][3]
        assert abs(float(g[k][1]) - float(w[k][1])) <= tol * abs(float(w[k][1]))
        assert abs(float(g[k][2]) - float(w[k][2])) <= tol * abs(float(w[k][2]))


@pytest.mark.parametrize("generic", [False, True])
def test_overflow_is_refused(ctx, oracle, generic):
    """numeric_mul is exact in the reference; the 64-bit scaled form here must refuse a product that leaves
    64 bits instead of wrapping (CBGPU_ERR_OVERFLOW), on every kernel"""
    n = 5000
    fo, fp = make(fact, n, seed=8)
    fp.columns[2][:] = 4 * 10**18          # amt
    sc = scan(1, fp, ["g", "amt"])
    from cloudberry_b200.tpch import _child_var
    v = _child_var(sc)
    big = P.OpExpr(P.OP_MUL, v("amt"), P.OpExpr(P.OP_SUB, P.NumericConst("100.00"), v("amt")))
    plan = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], [("g", v("g")), ("s", P.Aggref(P.AGG_SUM, big))], num_groups=16)
    dev = to_device(ctx, [fp])
    ex = capi.Executor(ctx, dev, force_generic=generic)
    with pytest.raises(capi.CbgpuError):
        ex.run(plan)
    ex.close()
    for d in dev:
        d.free()


@pytest.mark.parametrize("typ,host_dtype", [(P.NUMERIC, np.int8), (P.NUMERIC, np.int16), (P.NUMERIC, np.int32), (P.INT8, np.int16),
                                            (P.INT4, np.int8), (P.DATE, np.int16)])
def test_narrow_host_columns_are_sign_extended(ctx, typ, host_dtype):
    """cbgpu_rel_load_column_narrow: a host column shipped in a narrower integer width arrives as the column's own width, negative
    values included; widths that cannot be widened into the column are refused"""
    n = 100003
    rng = np.random.default_rng(int(typ) * 100 + np.dtype(host_dtype).itemsize)
    info = np.iinfo(host_dtype)
    vals = rng.integers(info.min, info.max, n, endpoint=True).astype(host_dtype)
    vals[:4] = [info.min, info.max, -1, 0]
    rel = capi.DeviceRelation(ctx, n, [typ], name="narrow")
    rel.load_column_ptr(0, vals.ctypes.data, vals.dtype.itemsize)
    got, nulls = rel.read_column(0)
    assert np.array_equal(got.astype(np.int64), vals.astype(np.int64)) and not nulls.any()
    with pytest.raises(capi.CbgpuError):
        ctx.check(ctx.L.cbgpu_rel_load_column_narrow(rel.h, 0, vals.ctypes.data, 3))
    rel.free()
    f = capi.DeviceRelation(ctx, n, [P.FLOAT8], name="f8")
    with pytest.raises(capi.CbgpuError):
        ctx.check(ctx.L.cbgpu_rel_load_column_narrow(f.h, 0, vals.ctypes.data, 4))
    f.free()


def test_prefilter_pass_on_small_inputs(oracle):
    """the two-kernel plan of big selective scans (k_prefilter leaves row ids, k_probe_chain starts from them) forced onto small
    tables (CBGPU_PREFILTER_MIN_ROWS=1; knobs are read when a context is created): join types, quals, visimap, NULL-free keys,
    empty sides - the same rows as the oracle's; and a pipeline it cannot thin is remembered and left to the fused kernel"""
    import os
    from cloudberry_b200 import tpch
    os.environ["CBGPU_PREFILTER_MIN_ROWS"] = "1"
    os.environ["CBGPU_PREFILTER_KEEP_DIV"] = "1"         # use the survivors however many they are
    try:
        c = capi.Context(0)
    finally:
        del os.environ["CBGPU_PREFILTER_MIN_ROWS"]
        del os.environ["CBGPU_PREFILTER_KEEP_DIV"]
    try:
        for jointype in (P.JOIN_INNER, P.JOIN_SEMI, P.JOIN_ANTI):
            for nf, nd, vis in ((5000, 40, 1.0), (100003, 30, 0.7), (2049, 1, 1.0), (5000, 0, 1.0)):
                fo, fp = make(fact, nf, seed=3, visible_frac=vis)
                do, dp = make(dim, nd, seed=4)
                sf = scan(1, fo, ["k", "amt", "g"], [P.OpExpr(P.OP_LT, P.Var(1, fo.attno("d"), P.DATE), P.Const(P.DATE, 1500))])
                sd = scan(2, do, ["dk", "w", "c"])
                h = P.Hash(sd, [P.out_var(sd, 1)])
                targets = [("k", P.out_var(sf, 1)), ("amt", P.out_var(sf, 2)), ("g", P.out_var(sf, 3))]
                if jointype == P.JOIN_INNER:
                    targets += [("w", P.InnerVar(2, P.INT8)), ("c", P.InnerVar(3, P.DICT8))]
                j = P.HashJoin(jointype, sf, h, [P.out_var(sf, 1)], targets)
                names = [t[0] for t in targets]
                aggs = [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None)] + ([("sw", P.AGG_SUM, "w")] if "w" in names else [])
                plan = agg_over(j, names, ["g"] + (["c"] if "c" in names else []), aggs)
                for _ in range(2):                       # the second run meets the selectivity cache
                    got, want = run_both(c, oracle, plan, [fo, do], [fp, dp], False)
                    assert canon(got) == canon(want), (jointype, nf, nd, vis)
        # TPC-H Q3 / Q5 shapes
        rels_o = tpch.gen_tables(0.05, oracle.hashbpchar)
        rels_p = tpch.gen_tables(0.05, capi.hashbpchar)
        dev = to_device(c, rels_p)
        ex = capi.Executor(c, dev)
        for plan, fmt in ((tpch.q3_plan(tpch.SEGMENTS.index("MACHINERY"), 1), tpch.format_q3),
                          (tpch.q5_plan(tpch.REGIONS.index("AMERICA"), 1), lambda r: tpch.format_q5(r, tpch.NATIONS))):
            assert fmt(ex.run(plan).rows) == fmt(oracle.execute(plan, [rels_o]).rows)
        names = {v["node"] for v in ex.run(tpch.q3_plan(tpch.SEGMENTS.index("MACHINERY"), 1)).instrument.values()}
        assert names
        ex.close()
        for d in dev:
            d.free()
    finally:
        c.close()


@pytest.mark.parametrize("generic", [False, True])
def test_float8_min_max(ctx, oracle, generic):
    """min / max(float8) order NaN above +Infinity and -0 below +0 like float8_cmp_internal / float8smaller / float8larger
    (utils/adt/float.c): negative values, infinities, NaN, NULLs; one stage and two-stage over three segments"""
    n = 30011
    rng = np.random.default_rng(71)
    x = rng.normal(0, 1e6, n)
    x[rng.integers(0, n, 40)] = np.inf
    x[rng.integers(0, n, 40)] = -np.inf
    g = rng.integers(0, 9, n)
    x[(g == 3) & (rng.random(n) < 0.01)] = np.nan            # group 3 sees NaNs: its max is NaN, its min is not
    x[g == 5] = np.abs(x[g == 5])                            # an all-positive group
    x[g == 6] = -np.abs(x[g == 6])                           # an all-negative group
    nulls = (rng.random(n) < 0.1).astype(np.uint8)
    nulls[g == 8] = 1                                        # a group without a single value: NULL min / max

    def rel():
        return HostRelation("f", ["g", "x"], [P.INT4, P.FLOAT8], [g.astype(np.int32), x], nulls=[None, nulls])
    from oracle import oracle as O
    fo, fp = rel().set_dict_hashes(O.hashbpchar), rel().set_dict_hashes(capi.hashbpchar)
    sc = scan(1, fo, ["g", "x"])
    plan = agg_over(sc, ["g", "x"], ["g"], [("mn", P.AGG_MIN, "x"), ("mx", P.AGG_MAX, "x"), ("c", P.AGG_COUNT, "x")])
    got, want = run_both(ctx, oracle, plan, [fo], [fp], generic)

    def key(rows):
        out = []
        for r in sorted(rows, key=lambda r: r[0]):
            out.append(tuple("nan" if isinstance(v, float) and v != v else v for v in r))
        return out
    assert key(got) == key(want)
    by = {r[0]: r for r in got}
    assert by[3][2] != by[3][2] and by[3][1] == by[3][1]     # max is NaN, min is a number
    assert by[8][1] is None and by[8][2] is None and by[8][3] == 0
    assert by[5][1] >= 0 and by[6][2] <= 0
    # two-stage: partial min / max states cross a Motion and are merged
    from cloudberry_b200 import tpch
    cut = np.array_split(np.arange(n), 3)
    segs_o, segs_p = [[fo.take(c)] for c in cut], [[fp.take(c)] for c in cut]
    v = tpch._child_var(sc)
    aggs = [("mn", P.Aggref(P.AGG_MIN, v("x"))), ("mx", P.Aggref(P.AGG_MAX, v("x")))]
    partial = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_INITIAL_SERIAL, [1], [("g", v("g"))] + aggs, num_groups=16)
    red = P.Motion(partial, P.MOTIONTYPE_HASH, [P.out_var(partial, 1)], 3)
    finals = [(nme, P.Aggref(ar.op, P.OuterVar(2 + i, *P.out_type(red, 2 + i)), restype=ar.restype, dscale=ar.dscale)) for i, (nme, ar) in enumerate(aggs)]
    final = P.Agg(red, P.AGG_HASHED, P.AGGSPLIT_FINAL_DESERIAL, [1], [("g", P.out_var(red, 1))] + finals, num_groups=16)
    plan2 = P.Motion(final, P.MOTIONTYPE_GATHER)
    want2 = oracle.execute(plan2, segs_o).rows
    dsegs = [to_device(ctx, s) for s in segs_p]
    cl = capi.Cluster(ctx, dsegs)
    got2 = cl.run(plan2).rows
    cl.close()
    for d in dsegs:
        for r in d:
            r.free()
    assert key(got2) == key(want2) == [r[:3] for r in key(got)]
