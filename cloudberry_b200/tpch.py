"""TPC-H shaped relations and the plan trees the reference planner emits for Q1 / Q3 / Q5.

Host-side harness code (plans + synthetic data description); no compute on the product path.

Schemas keep only the columns the scan -> join -> agg path projects, in the
reference's column order (src/test/regress/input/rpt_tpch.source:5-95).  Plan shapes follow what
the reference planner produces (src/test/regress/expected/aggregates.out:3313-3328:
Gather Motion <- Finalize HashAggregate <- Redistribute Motion <- Partial HashAggregate <- Hash Join
<- Seq Scan / Hash <- Seq Scan).
"""
import datetime
import json
import os

import numpy as np

from . import plan as P
from .relation import HostRelation

EPOCH = datetime.date(2000, 1, 1)


def date_to_days(y, m, d):
    return (datetime.date(y, m, d) - EPOCH).days


def days_to_text(days):
    """psql DateStyle 'MDY' text as the regression output prints dates (mm-dd-yyyy)."""
    dt = EPOCH + datetime.timedelta(days=int(days))
    return "%02d-%02d-%04d" % (dt.month, dt.day, dt.year)


# range table order used by every plan below (scanrelid = index + 1)
RT = ["lineitem", "orders", "customer", "supplier", "nation", "region"]
SCHEMA = {
    "lineitem": [("l_orderkey", P.INT8), ("l_suppkey", P.INT4), ("l_quantity", P.NUMERIC), ("l_extendedprice", P.NUMERIC),
                 ("l_discount", P.NUMERIC), ("l_tax", P.NUMERIC), ("l_returnflag", P.BPCHAR1), ("l_linestatus", P.BPCHAR1),
                 ("l_shipdate", P.DATE)],
    "orders": [("o_orderkey", P.INT8), ("o_custkey", P.INT4), ("o_orderdate", P.DATE), ("o_shippriority", P.INT4)],
    "customer": [("c_custkey", P.INT4), ("c_nationkey", P.INT4), ("c_mktsegment", P.DICT8)],
    "supplier": [("s_suppkey", P.INT4), ("s_nationkey", P.INT4)],
    "nation": [("n_nationkey", P.INT4), ("n_regionkey", P.INT4), ("n_name", P.DICT8)],
    "region": [("r_regionkey", P.INT4), ("r_name", P.DICT8)],
}
# how the reference distributes them (rpt_tpch.source: DISTRIBUTED BY / REPLICATED)
DIST_KEY = {"lineitem": "l_orderkey", "orders": "o_orderkey", "customer": None, "supplier": None, "nation": None,
            "region": None}

NATIONS = ["ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE", "GERMANY", "INDIA", "INDONESIA",
           "IRAN", "IRAQ", "JAPAN", "JORDAN", "KENYA", "MOROCCO", "MOZAMBIQUE", "PERU", "CHINA", "ROMANIA",
           "SAUDI ARABIA", "VIETNAM", "RUSSIA", "UNITED KINGDOM", "UNITED STATES"]
NATION_REGION = [0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]
SEGMENTS = ["AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"]


def _rel(name, cols, dict_texts=None):
    names = [n for n, _ in SCHEMA[name]]
    types = [t for _, t in SCHEMA[name]]
    dt = [None] * len(names)
    for k, v in (dict_texts or {}).items():
        dt[names.index(k)] = v
    return HostRelation(name, names, types, [cols[n] for n in names], dict_texts=dt)


def load_golden(hashfn):
    """The reference regression fixture (tests/golden/rpt_tpch.npz, made by tests/golden/make_golden.py
    from src/test/regress/data/*.csv) as a range table, plus the expected rows."""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    d = np.load(os.path.join(here, "rpt_tpch.npz"))
    exp = json.load(open(os.path.join(here, "rpt_tpch_expected.json")))
    dicts = exp["dict"]
    rels = [
        _rel("lineitem", d),
        _rel("orders", d),
        _rel("customer", d, {"c_mktsegment": dicts["c_mktsegment_dict"]}),
        _rel("supplier", d),
        _rel("nation", d, {"n_name": dicts["n_name_dict"]}),
        _rel("region", d, {"r_name": dicts["r_name_dict"]}),
    ]
    for r in rels:
        r.set_dict_hashes(hashfn)
    return rels, exp


# ------------------------------------------------------------------------------------------------
# synthetic TPC-H shaped generator (seeded, counter based: every value is a pure function of
# (seed, column id, row index), so the CUDA generator in csrc/tpch_gen.cu and this numpy version
# produce identical tables and any row range can be regenerated independently).
# ------------------------------------------------------------------------------------------------
def _mix(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def _u(seed, col, idx):
    with np.errstate(over="ignore"):
        return _mix(np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(col) * np.uint64(0xD1B54A32D192ED03)
                    + np.asarray(idx, dtype=np.uint64))


STARTDATE = date_to_days(1992, 1, 1)
CURRENTDATE = date_to_days(1995, 6, 17)
ORDER_DATE_SPAN = 2406          # 1992-01-01 .. 1998-08-02
_LINE_OFF = np.repeat(np.arange(7), np.arange(1, 8))       # 28 lines -> order offset within a block of 7 orders
_LINE_NO = np.concatenate([np.arange(1, k + 1) for k in range(1, 8)])


def sizes(sf):
    """Row counts at scale factor sf (TPC-H: lineitem 6,001,215 x sf at sf=1; exact at the
    BASELINE configs)."""
    table = {1: 6001215, 100: 600037902, 300: 1799989091}
    li = table.get(sf, int(round(6000000 * sf)))
    return {"lineitem": li, "orders": int(1500000 * sf), "customer": int(150000 * sf), "supplier": int(10000 * sf),
            "nation": 25, "region": 5, "part": int(200000 * sf)}


def order_key(idx):
    """dbgen's sparse order keys: 8 of every 32."""
    idx = np.asarray(idx, dtype=np.int64)
    return (idx // 8) * 32 + (idx % 8) + 1


def gen_orders(seed, n_orders, n_cust, lo=0, hi=None):
    hi = n_orders if hi is None else hi
    i = np.arange(lo, hi, dtype=np.int64)
    return {
        "o_orderkey": order_key(i),
        "o_custkey": (1 + _u(seed, 11, i) % np.uint64(n_cust)).astype(np.int32),
        "o_orderdate": (STARTDATE + (_u(seed, 12, i) % np.uint64(ORDER_DATE_SPAN)).astype(np.int64)).astype(np.int32),
        "o_shippriority": np.zeros(hi - lo, dtype=np.int32),
    }


def gen_lineitem(seed, n_rows, n_supp, n_part, lo=0, hi=None):
    hi = n_rows if hi is None else hi
    j = np.arange(lo, hi, dtype=np.int64)
    oidx = (j // 28) * 7 + _LINE_OFF[j % 28]
    odate = STARTDATE + (_u(seed, 12, oidx) % np.uint64(ORDER_DATE_SPAN)).astype(np.int64)
    qty = 1 + (_u(seed, 21, j) % np.uint64(50)).astype(np.int64)
    pk = 1 + (_u(seed, 22, j) % np.uint64(n_part)).astype(np.int64)
    price = 90000 + (pk // 10) % 20001 + 100 * (pk % 1000)          # cents (dbgen retail price rule)
    ship = odate + 1 + (_u(seed, 25, j) % np.uint64(121)).astype(np.int64)
    receipt = ship + 1 + (_u(seed, 26, j) % np.uint64(30)).astype(np.int64)
    ra = np.where(_u(seed, 27, j) % np.uint64(2) == 0, ord("R"), ord("A"))
    return {
        "l_orderkey": order_key(oidx),
        "l_suppkey": (1 + _u(seed, 23, j) % np.uint64(n_supp)).astype(np.int32),
        "l_quantity": qty * 100,
        "l_extendedprice": qty * price,
        "l_discount": (_u(seed, 24, j) % np.uint64(11)).astype(np.int64),
        "l_tax": (_u(seed, 28, j) % np.uint64(9)).astype(np.int64),
        "l_returnflag": np.where(receipt <= CURRENTDATE, ra, ord("N")).astype(np.uint8),
        "l_linestatus": np.where(ship > CURRENTDATE, ord("O"), ord("F")).astype(np.uint8),
        "l_shipdate": ship.astype(np.int32),
    }


def gen_customer(seed, n):
    i = np.arange(n, dtype=np.int64)
    return {"c_custkey": (i + 1).astype(np.int32), "c_nationkey": (_u(seed, 31, i) % np.uint64(25)).astype(np.int32),
            "c_mktsegment": (_u(seed, 32, i) % np.uint64(5)).astype(np.uint8)}


def gen_supplier(seed, n):
    i = np.arange(n, dtype=np.int64)
    return {"s_suppkey": (i + 1).astype(np.int32), "s_nationkey": (_u(seed, 41, i) % np.uint64(25)).astype(np.int32)}


def gen_nation_region():
    nation = {"n_nationkey": np.arange(25, dtype=np.int32), "n_regionkey": np.array(NATION_REGION, dtype=np.int32),
              "n_name": np.arange(25, dtype=np.uint8)}
    region = {"r_regionkey": np.arange(5, dtype=np.int32), "r_name": np.arange(5, dtype=np.uint8)}
    return nation, region


def gen_tables(sf, hashfn, seed=42, li_rows=None):
    """Whole synthetic database on the host (small scale factors only)."""
    sz = sizes(sf)
    if li_rows is not None:
        sz["lineitem"] = li_rows
    nation, region = gen_nation_region()
    rels = [
        _rel("lineitem", gen_lineitem(seed, sz["lineitem"], sz["supplier"], sz["part"])),
        _rel("orders", gen_orders(seed, sz["orders"], sz["customer"])),
        _rel("customer", gen_customer(seed, sz["customer"]), {"c_mktsegment": SEGMENTS}),
        _rel("supplier", gen_supplier(seed, sz["supplier"])),
        _rel("nation", nation, {"n_name": NATIONS}),
        _rel("region", region, {"r_name": REGIONS}),
    ]
    for r in rels:
        r.set_dict_hashes(hashfn)
    return rels


# ------------------------------------------------------------------------------------------------
# plans
# ------------------------------------------------------------------------------------------------
def _scan(relname, cols, quals=()):
    """SeqScan over RT[relname] projecting `cols` (aoco_beginscan_extractcolumns,
    access/aocs/aocsam_handler.c:612: only referenced columns are read)."""
    relid = RT.index(relname) + 1
    schema = SCHEMA[relname]
    names = [n for n, _ in schema]
    tl = []
    for c in cols:
        i = names.index(c)
        t = schema[i][1]
        tl.append((c, P.Var(relid, i + 1, t, 2 if t == P.NUMERIC else 0)))
    return P.SeqScan(relid, tl, quals)


def _svar(relname, col):
    relid = RT.index(relname) + 1
    names = [n for n, _ in SCHEMA[relname]]
    i = names.index(col)
    t = SCHEMA[relname][i][1]
    return P.Var(relid, i + 1, t, 2 if t == P.NUMERIC else 0)


def _one():
    return P.NumericConst("1")


def q1_aggs(v):
    """The ten output expressions of Q1 over a child whose output attnos are given by v(name)."""
    disc_price = P.OpExpr(P.OP_MUL, v("l_extendedprice"), P.OpExpr(P.OP_SUB, _one(), v("l_discount")))
    charge = P.OpExpr(P.OP_MUL, P.OpExpr(P.OP_MUL, v("l_extendedprice"), P.OpExpr(P.OP_SUB, _one(), v("l_discount"))),
                      P.OpExpr(P.OP_ADD, _one(), v("l_tax")))
    return [
        ("sum_qty", P.Aggref(P.AGG_SUM, v("l_quantity"))),
        ("sum_base_price", P.Aggref(P.AGG_SUM, v("l_extendedprice"))),
        ("sum_disc_price", P.Aggref(P.AGG_SUM, disc_price)),
        ("sum_charge", P.Aggref(P.AGG_SUM, charge)),
        ("avg_qty", P.Aggref(P.AGG_AVG, v("l_quantity"))),
        ("avg_price", P.Aggref(P.AGG_AVG, v("l_extendedprice"))),
        ("avg_disc", P.Aggref(P.AGG_AVG, v("l_discount"))),
        ("count_order", P.Aggref(P.AGG_COUNT_STAR)),
    ]


Q1_CUTOFF = date_to_days(1998, 8, 15)      # date '1998-12-01' - interval '108 day'


def _child_var(child):
    names = [child.plan.targetlist[i].resname.decode() for i in range(child.plan.ntargets)]

    def v(name):
        i = names.index(name)
        t, ds = P.out_type(child, i + 1)
        return P.OuterVar(i + 1, t, ds)
    return v


def q1_plan(nsegs=1, cutoff=Q1_CUTOFF):
    """TPC-H Q1 (rpt_tpch.source:346-371).  One segment: HashAggregate <- Seq Scan.  Several:
    Gather Motion <- Finalize HashAggregate <- Redistribute Motion (l_returnflag, l_linestatus)
    <- Partial HashAggregate <- Seq Scan."""
    scan = _scan("lineitem", ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"],
                 [P.OpExpr(P.OP_LE, _svar("lineitem", "l_shipdate"), P.Const(P.DATE, cutoff))])
    v = _child_var(scan)
    keys = [("l_returnflag", v("l_returnflag")), ("l_linestatus", v("l_linestatus"))]
    if nsegs == 1:
        return P.Agg(scan, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1, 2], keys + q1_aggs(v), num_groups=6)
    partial = P.Agg(scan, P.AGG_HASHED, P.AGGSPLIT_INITIAL_SERIAL, [1, 2], keys + q1_aggs(v), num_groups=6,
                    streaming=True)
    redist = P.Motion(partial, P.MOTIONTYPE_HASH, [P.out_var(partial, 1), P.out_var(partial, 2)], nsegs)
    # the final stage's Aggrefs take the partial states (columns 3..10 of the Motion output)
    finals = []
    for i, (name, ar) in enumerate(q1_aggs(v)):
        t, ds = P.out_type(redist, 3 + i)
        finals.append((name, P.Aggref(ar.op, P.OuterVar(3 + i, t, ds), restype=ar.restype, dscale=ar.dscale)))
    final = P.Agg(redist, P.AGG_HASHED, P.AGGSPLIT_FINAL_DESERIAL, [1, 2],
                  [("l_returnflag", P.out_var(redist, 1)), ("l_linestatus", P.out_var(redist, 2))] + finals, num_groups=6)
    return P.Motion(final, P.MOTIONTYPE_GATHER)


def q3_plan(segment_code, nsegs=1, cutoff=None, limit=10, customer_replicated=True, merge_gather=False):
    """TPC-H Q3 (rpt_tpch.source:458-480).
    Limit/Sort <- [Gather] <- HashAggregate(l_orderkey, o_orderdate, o_shippriority)
       <- Hash Join (l_orderkey = o_orderkey)
            <- Seq Scan lineitem (l_shipdate > d)
            <- Hash <- Hash Join (o_custkey = c_custkey)
                         <- Seq Scan orders (o_orderdate < d)
                         <- Hash <- Seq Scan customer (c_mktsegment = seg)
    lineitem and orders are co-located on orderkey; customer is replicated (as in rpt_tpch) or, with
    customer_replicated=False, distributed by c_custkey, which puts a Redistribute Motion on
    o_custkey under the lower join and one on o_orderkey above it."""
    cutoff = date_to_days(1995, 3, 15) if cutoff is None else cutoff
    cust = _scan("customer", ["c_custkey"], [P.OpExpr(P.OP_EQ, _svar("customer", "c_mktsegment"), P.Const(P.DICT8, segment_code))])
    orders = _scan("orders", ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"],
                   [P.OpExpr(P.OP_LT, _svar("orders", "o_orderdate"), P.Const(P.DATE, cutoff))])
    outer = orders
    if not customer_replicated and nsegs > 1:
        outer = P.Motion(orders, P.MOTIONTYPE_HASH, [P.out_var(orders, 2)], nsegs)
    hc = P.Hash(cust, [P.out_var(cust, 1)])
    j1 = P.HashJoin(P.JOIN_INNER, outer, hc, [P.out_var(outer, 2)],
                    [("o_orderkey", P.out_var(outer, 1)), ("o_orderdate", P.out_var(outer, 3)),
                     ("o_shippriority", P.out_var(outer, 4))])
    inner = j1
    if not customer_replicated and nsegs > 1:
        inner = P.Motion(j1, P.MOTIONTYPE_HASH, [P.out_var(j1, 1)], nsegs)
    ho = P.Hash(inner, [P.out_var(inner, 1)])
    li = _scan("lineitem", ["l_orderkey", "l_extendedprice", "l_discount"],
               [P.OpExpr(P.OP_GT, _svar("lineitem", "l_shipdate"), P.Const(P.DATE, cutoff))])
    j2 = P.HashJoin(P.JOIN_INNER, li, ho, [P.out_var(li, 1)],
                    [("l_orderkey", P.out_var(li, 1)), ("o_orderdate", P.InnerVar(2, P.DATE)),
                     ("o_shippriority", P.InnerVar(3, P.INT4)),
                     ("l_extendedprice", P.out_var(li, 2)), ("l_discount", P.out_var(li, 3))])
    v = _child_var(j2)
    rev = P.OpExpr(P.OP_MUL, v("l_extendedprice"), P.OpExpr(P.OP_SUB, _one(), v("l_discount")))
    agg = P.Agg(j2, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1, 2, 3],
                [("l_orderkey", v("l_orderkey")), ("revenue", P.Aggref(P.AGG_SUM, rev)),
                 ("o_orderdate", v("o_orderdate")), ("o_shippriority", v("o_shippriority"))], num_groups=1000000)
    top = P.LimitSort(agg, [(2, True), (3, False)], limit)
    if nsegs == 1:
        return top
    # each segment keeps its local top-N, the gather receiver merges (Limit <- Gather Motion (merge)
    # <- Limit <- Sort in the reference's plan)
    g = P.Motion(top, P.MOTIONTYPE_GATHER, sort_keys=[(2, True), (3, False)] if merge_gather else ())
    return P.LimitSort(g, [(2, True), (3, False)], limit)


def q5_plan(region_code, nsegs=1, date_lo=None, date_hi=None, replicated=True):
    """TPC-H Q5 (rpt_tpch.source:512-535): six-table join chain, group by n_name.
    HashAggregate(n_name)
      <- Hash Join (s_nationkey = n_nationkey)          inner: Hash <- Hash Join nation x region(r_name = R)
      <- Hash Join (l_suppkey = s_suppkey AND c_nationkey = s_nationkey)   inner: Hash <- supplier
      <- Hash Join (o_custkey = c_custkey)              inner: Hash <- customer
      <- Hash Join (l_orderkey = o_orderkey)            inner: Hash <- orders (date range)
      <- Seq Scan lineitem
    With replicated=False (customer by c_custkey, supplier by s_suppkey) the chain gets a
    Redistribute Motion before the customer join and another before the supplier join."""
    date_lo = date_to_days(1997, 1, 1) if date_lo is None else date_lo
    date_hi = date_to_days(1998, 1, 1) if date_hi is None else date_hi
    multi = (not replicated) and nsegs > 1
    li = _scan("lineitem", ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"])
    orders = _scan("orders", ["o_orderkey", "o_custkey"],
                   [P.OpExpr(P.OP_GE, _svar("orders", "o_orderdate"), P.Const(P.DATE, date_lo)),
                    P.OpExpr(P.OP_LT, _svar("orders", "o_orderdate"), P.Const(P.DATE, date_hi))])
    ho = P.Hash(orders, [P.out_var(orders, 1)])
    j1 = P.HashJoin(P.JOIN_INNER, li, ho, [P.out_var(li, 1)],
                    [("o_custkey", P.InnerVar(2, P.INT4)), ("l_suppkey", P.out_var(li, 2)),
                     ("l_extendedprice", P.out_var(li, 3)), ("l_discount", P.out_var(li, 4))])
    o1 = P.Motion(j1, P.MOTIONTYPE_HASH, [P.out_var(j1, 1)], nsegs) if multi else j1
    cust = _scan("customer", ["c_custkey", "c_nationkey"])
    hc = P.Hash(cust, [P.out_var(cust, 1)])
    j2 = P.HashJoin(P.JOIN_INNER, o1, hc, [P.out_var(o1, 1)],
                    [("c_nationkey", P.InnerVar(2, P.INT4)), ("l_suppkey", P.out_var(o1, 2)),
                     ("l_extendedprice", P.out_var(o1, 3)), ("l_discount", P.out_var(o1, 4))])
    o2 = P.Motion(j2, P.MOTIONTYPE_HASH, [P.out_var(j2, 2)], nsegs) if multi else j2
    supp = _scan("supplier", ["s_suppkey", "s_nationkey"])
    hs = P.Hash(supp, [P.out_var(supp, 1), P.out_var(supp, 2)])
    j3 = P.HashJoin(P.JOIN_INNER, o2, hs, [P.out_var(o2, 2), P.out_var(o2, 1)],
                    [("s_nationkey", P.InnerVar(2, P.INT4)), ("l_extendedprice", P.out_var(o2, 3)),
                     ("l_discount", P.out_var(o2, 4))])
    region = _scan("region", ["r_regionkey"], [P.OpExpr(P.OP_EQ, _svar("region", "r_name"), P.Const(P.DICT8, region_code))])
    hr = P.Hash(region, [P.out_var(region, 1)])
    nation = _scan("nation", ["n_nationkey", "n_regionkey", "n_name"])
    jn = P.HashJoin(P.JOIN_INNER, nation, hr, [P.out_var(nation, 2)],
                    [("n_nationkey", P.out_var(nation, 1)), ("n_name", P.out_var(nation, 3))])
    hn = P.Hash(jn, [P.out_var(jn, 1)])
    j4 = P.HashJoin(P.JOIN_INNER, j3, hn, [P.out_var(j3, 1)],
                    [("n_name", P.InnerVar(2, P.DICT8)), ("l_extendedprice", P.out_var(j3, 2)),
                     ("l_discount", P.out_var(j3, 3))])
    v = _child_var(j4)
    rev = P.OpExpr(P.OP_MUL, v("l_extendedprice"), P.OpExpr(P.OP_SUB, _one(), v("l_discount")))
    if nsegs == 1:
        return P.Agg(j4, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], [("n_name", v("n_name")), ("revenue", P.Aggref(P.AGG_SUM, rev))],
                     num_groups=25)
    partial = P.Agg(j4, P.AGG_HASHED, P.AGGSPLIT_INITIAL_SERIAL, [1],
                    [("n_name", v("n_name")), ("revenue", P.Aggref(P.AGG_SUM, rev))], num_groups=25, streaming=True)
    redist = P.Motion(partial, P.MOTIONTYPE_HASH, [P.out_var(partial, 1)], nsegs)
    t, ds = P.out_type(redist, 2)
    final = P.Agg(redist, P.AGG_HASHED, P.AGGSPLIT_FINAL_DESERIAL, [1],
                  [("n_name", P.out_var(redist, 1)),
                   ("revenue", P.Aggref(P.AGG_SUM, P.OuterVar(2, t, ds), restype=P.NUMERIC, dscale=4))], num_groups=25)
    return P.Motion(final, P.MOTIONTYPE_GATHER)


# ------------------------------------------------------------------------------------------------
# result formatting, as the regression output prints the rows
# ------------------------------------------------------------------------------------------------
def format_q1(rows):
    out = []
    for r in sorted(rows, key=lambda r: (r[0], r[1])):
        out.append([chr(r[0]), chr(r[1])] + [str(x) for x in r[2:]])
    return out


def format_q3(rows):
    return [[str(r[0]), str(r[1]), days_to_text(r[2]), str(r[3])] for r in rows]


def format_q5(rows, nation_dict):
    from decimal import Decimal
    rs = sorted(rows, key=lambda r: -Decimal(r[1]))
    return [[nation_dict[r[0]], str(r[1])] for r in rs]
