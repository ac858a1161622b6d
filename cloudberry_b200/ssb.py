"""Star Schema Benchmark shaped tables and the Q4.x plans (BASELINE.json configs[4]: "SSB SF100 Q4.x ...
wide hash-agg, HBM-bound group-by").

The reference holds no SSB fixture (parity unpinned for SSB), so the tables are synthetic
(seeded, counter based like tpch.py) and parity is CUDA path == CPU oracle == an independent numpy
evaluation of the same SQL (tests/test_oracle_ssb.py, tests/test_gpu_ssb.py).

Plans are the star join the reference's planner emits for Q4.x with replicated dimensions
(HashAggregate <- Hash Join x4 <- Seq Scan lineorder, every dimension under its own Hash, dimension
quals pushed into the dimension scans): four N:1 probes from the fact table, group keys taken from the
inner sides, `sum(lo_revenue - lo_supplycost)`.
"""
import numpy as np

from . import plan as P
from .relation import HostRelation
from .tpch import _u, _child_var

RT = ["lineorder", "customer", "supplier", "part", "date"]
SCHEMA = {
    "lineorder": [("lo_custkey", P.INT4), ("lo_partkey", P.INT4), ("lo_suppkey", P.INT4), ("lo_orderdate", P.INT4),
                  ("lo_revenue", P.INT8), ("lo_supplycost", P.INT8)],
    "customer": [("c_custkey", P.INT4), ("c_city", P.DICT32), ("c_nation", P.DICT8), ("c_region", P.DICT8)],
    "supplier": [("s_suppkey", P.INT4), ("s_city", P.DICT32), ("s_nation", P.DICT8), ("s_region", P.DICT8)],
    "part": [("p_partkey", P.INT4), ("p_mfgr", P.DICT8), ("p_category", P.DICT8), ("p_brand1", P.DICT32)],
    "date": [("d_datekey", P.INT4), ("d_year", P.INT4)],
}
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]
NATIONS = ["ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE", "GERMANY", "INDIA", "INDONESIA",
           "IRAN", "IRAQ", "JAPAN", "JORDAN", "KENYA", "MOROCCO", "MOZAMBIQUE", "PERU", "CHINA", "ROMANIA",
           "SAUDI ARABIA", "VIETNAM", "RUSSIA", "UNITED KINGDOM", "UNITED STATES"]
NATION_REGION = [0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1]
CITIES = ["%-9.9s%d" % (n, d) for n in NATIONS for d in range(10)]          # 250 cities, 10 per nation
MFGRS = ["MFGR#%d" % m for m in range(1, 6)]
CATEGORIES = ["MFGR#%d%d" % (m, c) for m in range(1, 6) for c in range(1, 6)]
BRANDS = ["MFGR#%d%d%02d" % (m, c, b) for m in range(1, 6) for c in range(1, 6) for b in range(1, 41)]


def sizes(sf):
    return {"lineorder": int(6000000 * sf), "customer": max(int(30000 * sf), 50), "supplier": max(int(2000 * sf), 20),
            "part": max(int(200000 * (1 + np.log2(max(sf, 1)))) if sf >= 1 else int(200000 * sf), 100), "date": 2556}


def _dates():
    """d_datekey as yyyymmdd for 1992-01-01 .. 1998-12-31 (2556 days, leap years 1992 and 1996), d_year."""
    keys, years = [], []
    mdays = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    for y in range(1992, 1999):
        for m in range(12):
            n = mdays[m] + (1 if m == 1 and y % 4 == 0 else 0)
            for d in range(1, n + 1):
                keys.append(y * 10000 + (m + 1) * 100 + d)
                years.append(y)
    return np.array(keys[:2556], dtype=np.int32), np.array(years[:2556], dtype=np.int32)


def _rel(name, cols, dict_texts=None):
    names = [n for n, _ in SCHEMA[name]]
    types = [t for _, t in SCHEMA[name]]
    dt = [None] * len(names)
    for k, v in (dict_texts or {}).items():
        dt[names.index(k)] = v
    return HostRelation(name, names, types, [cols[n] for n in names], dscales=[0] * len(names), dict_texts=dt)


def gen_tables(sf, hashfn, seed=7, fact=True):
    """fact=False: the four dimensions only, lineorder as an empty relation (bench.py generates the fact table on the
    device, csrc/gen.cu cbgpu_gen_ssb_lineorder: the same formulas)"""
    sz = sizes(sf)
    if not fact:
        sz = dict(sz, lineorder=0)
    dk, dy = _dates()
    i = np.arange(sz["customer"], dtype=np.int64)
    c_city = (_u(seed, 21, i) % np.uint64(250)).astype(np.int32)
    cust = {"c_custkey": (i + 1).astype(np.int32), "c_city": c_city, "c_nation": (c_city // 10).astype(np.uint8),
            "c_region": np.array(NATION_REGION, dtype=np.uint8)[c_city // 10]}
    i = np.arange(sz["supplier"], dtype=np.int64)
    s_city = (_u(seed, 31, i) % np.uint64(250)).astype(np.int32)
    supp = {"s_suppkey": (i + 1).astype(np.int32), "s_city": s_city, "s_nation": (s_city // 10).astype(np.uint8),
            "s_region": np.array(NATION_REGION, dtype=np.uint8)[s_city // 10]}
    i = np.arange(sz["part"], dtype=np.int64)
    brand = (_u(seed, 41, i) % np.uint64(1000)).astype(np.int32)
    part = {"p_partkey": (i + 1).astype(np.int32), "p_brand1": brand, "p_category": (brand // 40).astype(np.uint8),
            "p_mfgr": (brand // 200).astype(np.uint8)}
    i = np.arange(sz["lineorder"], dtype=np.int64)
    lo = {
        "lo_custkey": (1 + _u(seed, 51, i) % np.uint64(sz["customer"])).astype(np.int32),
        "lo_partkey": (1 + _u(seed, 52, i) % np.uint64(sz["part"])).astype(np.int32),
        "lo_suppkey": (1 + _u(seed, 53, i) % np.uint64(sz["supplier"])).astype(np.int32),
        "lo_orderdate": dk[(_u(seed, 54, i) % np.uint64(2556)).astype(np.int64)],
        "lo_revenue": (100 + _u(seed, 55, i) % np.uint64(10000000)).astype(np.int64),
        "lo_supplycost": (50 + _u(seed, 56, i) % np.uint64(120000)).astype(np.int64),
    }
    rels = [
        _rel("lineorder", lo),
        _rel("customer", cust, {"c_city": CITIES, "c_nation": NATIONS, "c_region": REGIONS}),
        _rel("supplier", supp, {"s_city": CITIES, "s_nation": NATIONS, "s_region": REGIONS}),
        _rel("part", part, {"p_mfgr": MFGRS, "p_category": CATEGORIES, "p_brand1": BRANDS}),
        _rel("date", {"d_datekey": dk, "d_year": dy}),
    ]
    for r in rels:
        r.set_dict_hashes(hashfn)
    return rels


# ------------------------------------------------------------------------------------------------
# plans
# ------------------------------------------------------------------------------------------------
def _var(relname, col):
    relid = RT.index(relname) + 1
    names = [n for n, _ in SCHEMA[relname]]
    i = names.index(col)
    return P.Var(relid, i + 1, SCHEMA[relname][i][1], 0)


def _scan(relname, cols, quals=()):
    relid = RT.index(relname) + 1
    return P.SeqScan(relid, [(c, _var(relname, c)) for c in cols], quals)


def _eq(relname, col, code):
    t = dict(SCHEMA[relname])[col]
    return P.OpExpr(P.OP_EQ, _var(relname, col), P.Const(t, code))


def _in(relname, col, codes):
    """col IN (...) as the OR of equalities the planner would emit for a two-element list"""
    es = [_eq(relname, col, c) for c in codes]
    return es[0] if len(es) == 1 else P.BoolExpr(P.OR_EXPR, *es)


def _star(dims, group, nsegs=1, num_groups=1000):
    """lineorder probes the dimension hash tables in `dims` order.
    dims: [(relname, key column, fact column, projected columns, quals)]; group: output names of the group keys."""
    fact_cols = ["lo_custkey", "lo_partkey", "lo_suppkey", "lo_orderdate", "lo_revenue", "lo_supplycost"]
    cur = _scan("lineorder", fact_cols)
    carried = list(fact_cols)
    for relname, key, factcol, proj, quals in dims:
        d = _scan(relname, [key] + proj, quals)
        h = P.Hash(d, [P.out_var(d, 1)])
        v = _child_var(cur)
        targets = [(c, v(c)) for c in carried]
        for j, c in enumerate(proj):
            t = dict(SCHEMA[relname])[c]
            targets.append((c, P.InnerVar(2 + j, t)))
        cur = P.HashJoin(P.JOIN_INNER, cur, h, [v(factcol)], targets)
        carried += proj
    v = _child_var(cur)
    profit = P.OpExpr(P.OP_SUB, v("lo_revenue"), v("lo_supplycost"))
    targets = [(g, v(g)) for g in group] + [("profit", P.Aggref(P.AGG_SUM, profit))]
    # grpColIdx: the group keys' positions in the child's target list (plannodes.h:1342 Agg.grpColIdx)
    grp = [carried.index(g) + 1 for g in group]
    if nsegs == 1:
        return P.Agg(cur, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, grp, targets, num_groups=num_groups)
    # lineorder distributed, dimensions replicated: Gather Motion <- Finalize HashAggregate <- Redistribute Motion (group
    # keys) <- Partial HashAggregate <- the star join (the two-stage shape of expected/aggregates.out:3313-3328)
    partial = P.Agg(cur, P.AGG_HASHED, P.AGGSPLIT_INITIAL_SERIAL, grp, targets, num_groups=num_groups, streaming=True)
    nk = len(group)
    redist = P.Motion(partial, P.MOTIONTYPE_HASH, [P.out_var(partial, k + 1) for k in range(nk)], nsegs)
    t, ds = P.out_type(redist, nk + 1)
    final = P.Agg(redist, P.AGG_HASHED, P.AGGSPLIT_FINAL_DESERIAL, list(range(1, nk + 1)),
                  [(g, P.out_var(redist, k + 1)) for k, g in enumerate(group)] +
                  [("profit", P.Aggref(P.AGG_SUM, P.OuterVar(nk + 1, t, ds), restype=P.NUMERIC, dscale=0))], num_groups=num_groups)
    return P.Motion(final, P.MOTIONTYPE_GATHER)


def q4_1_plan(nsegs=1):
    """select d_year, c_nation, sum(lo_revenue - lo_supplycost) ... where c_region = 'AMERICA' and
    s_region = 'AMERICA' and (p_mfgr = 'MFGR#1' or p_mfgr = 'MFGR#2') group by d_year, c_nation"""
    am = REGIONS.index("AMERICA")
    return _star([
        ("supplier", "s_suppkey", "lo_suppkey", [], [_eq("supplier", "s_region", am)]),
        ("customer", "c_custkey", "lo_custkey", ["c_nation"], [_eq("customer", "c_region", am)]),
        ("part", "p_partkey", "lo_partkey", [], [_in("part", "p_mfgr", [0, 1])]),
        ("date", "d_datekey", "lo_orderdate", ["d_year"], []),
    ], ["d_year", "c_nation"], nsegs=nsegs, num_groups=200)


def q4_2_plan(nsegs=1):
    """... where c_region = 'AMERICA' and s_region = 'AMERICA' and (d_year = 1997 or d_year = 1998) and
    (p_mfgr = 'MFGR#1' or p_mfgr = 'MFGR#2') group by d_year, s_nation, p_category"""
    am = REGIONS.index("AMERICA")
    return _star([
        ("supplier", "s_suppkey", "lo_suppkey", ["s_nation"], [_eq("supplier", "s_region", am)]),
        ("customer", "c_custkey", "lo_custkey", [], [_eq("customer", "c_region", am)]),
        ("part", "p_partkey", "lo_partkey", ["p_category"], [_in("part", "p_mfgr", [0, 1])]),
        ("date", "d_datekey", "lo_orderdate", ["d_year"], [_in("date", "d_year", [1997, 1998])]),
    ], ["d_year", "s_nation", "p_category"], nsegs=nsegs, num_groups=500)


def q4_3_plan(nsegs=1):
    """... where s_nation = 'UNITED STATES' and (d_year = 1997 or d_year = 1998) and p_category = 'MFGR#14'
    group by d_year, s_city, p_brand1"""
    return _star([
        ("supplier", "s_suppkey", "lo_suppkey", ["s_city"], [_eq("supplier", "s_nation", NATIONS.index("UNITED STATES"))]),
        ("part", "p_partkey", "lo_partkey", ["p_brand1"], [_eq("part", "p_category", CATEGORIES.index("MFGR#14"))]),
        ("date", "d_datekey", "lo_orderdate", ["d_year"], [_in("date", "d_year", [1997, 1998])]),
        ("customer", "c_custkey", "lo_custkey", [], []),
    ], ["d_year", "s_city", "p_brand1"], nsegs=nsegs, num_groups=2000)


PLANS = {"q4.1": q4_1_plan, "q4.2": q4_2_plan, "q4.3": q4_3_plan}


# ------------------------------------------------------------------------------------------------
# independent evaluation of the same SQL with numpy (no executor, no hashing): the answer both the
# oracle and the CUDA path must give
# ------------------------------------------------------------------------------------------------
def numpy_answer(q, rels):
    lo, cust, supp, part, date = [dict(zip(r.names, r.columns)) for r in rels]
    am = REGIONS.index("AMERICA")
    year_of = dict(zip(date["d_datekey"].tolist(), date["d_year"].tolist()))
    d_year = np.array([year_of[k] for k in lo["lo_orderdate"].tolist()], dtype=np.int64)
    c = lo["lo_custkey"].astype(np.int64) - 1
    s = lo["lo_suppkey"].astype(np.int64) - 1
    p = lo["lo_partkey"].astype(np.int64) - 1
    profit = lo["lo_revenue"] - lo["lo_supplycost"]
    if q == "q4.1":
        m = (cust["c_region"][c] == am) & (supp["s_region"][s] == am) & (part["p_mfgr"][p] <= 1)
        keys = [d_year, cust["c_nation"][c].astype(np.int64)]
    elif q == "q4.2":
        m = ((cust["c_region"][c] == am) & (supp["s_region"][s] == am) & (part["p_mfgr"][p] <= 1) &
             ((d_year == 1997) | (d_year == 1998)))
        keys = [d_year, supp["s_nation"][s].astype(np.int64), part["p_category"][p].astype(np.int64)]
    else:
        m = ((supp["s_nation"][s] == NATIONS.index("UNITED STATES")) & ((d_year == 1997) | (d_year == 1998)) &
             (part["p_category"][p] == CATEGORIES.index("MFGR#14")))
        keys = [d_year, supp["s_city"][s].astype(np.int64), part["p_brand1"][p].astype(np.int64)]
    out = {}
    ks = [k[m] for k in keys]
    for row in zip(*[k.tolist() for k in ks], profit[m].tolist()):
        out[row[:-1]] = out.get(row[:-1], 0) + row[-1]
    return sorted([list(k) + [str(v)] for k, v in out.items()])


def canon(rows):
    return sorted([[int(x) if not isinstance(x, str) else x for x in r[:-1]] + [str(r[-1])] for r in rows])


# projected bytes per row of every table Q4.x scans: lineorder custkey 4 + suppkey 4 + partkey 4 + orderdate 4
# + revenue 8 + supplycost 8; the dimensions' key + attribute columns
QUERY_BYTES = {"lineorder": 32, "customer": 10, "supplier": 10, "part": 10, "date": 8}


def query_rows_bytes(sz):
    return sum(sz[t] for t in QUERY_BYTES), sum(sz[t] * w for t, w in QUERY_BYTES.items())


def device_tables(ctx, sf, hashfn, rank=0, world=1, seed=7):
    """The SSB database on this GPU-segment: the dimensions whole (DISTRIBUTED REPLICATED), rows [rank * n / world,
    (rank + 1) * n / world) of lineorder (any distribution is valid for a star join against replicated dimensions),
    generated on the device."""
    from . import capi
    sz = sizes(sf)
    host = gen_tables(sf, hashfn, seed=seed, fact=False)
    n = sz["lineorder"]
    lo, hi = rank * n // world, (rank + 1) * n // world
    fact = capi.DeviceRelation(ctx, hi - lo, [t for _, t in SCHEMA["lineorder"]], dscales=[0] * 6, name="lineorder")
    ctx.check(ctx.L.cbgpu_gen_ssb_lineorder(ctx.h, fact.h, seed, lo, sz["customer"], sz["part"], sz["supplier"]))
    dev = [fact] + [capi.DeviceRelation.from_host(ctx, r) for r in host[1:]]
    ctx.sync()
    return dev, sz
