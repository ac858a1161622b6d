"""BASELINE configs[0]: TPC-H SF1 Q1 on 2 CPU segments (AOCS lineitem scan + hash aggregate; plumbing, no GPU).

The oracle runs the two-stage plan (Partial HashAggregate -> Redistribute Motion -> Finalize HashAggregate -> Gather Motion,
src/test/regress/expected/aggregates.out:3313-3328) over the 6 001 215-row synthetic lineitem, sharded by cdbhash(l_orderkey)
onto two segments and executed by two threads; the answer is checked against an independent numpy evaluation of the SQL
(exact integers), with the numeric text of sum / avg taken from the reference's own numeric_div where oracle/_ref is built."""
import ctypes as C
import os

import numpy as np

from cloudberry_b200 import plan as P
from cloudberry_b200 import tpch

HERE = os.path.dirname(os.path.abspath(__file__))


def _rot(x, k):
    return ((x << np.uint32(k)) | (x >> np.uint32(32 - k))).astype(np.uint32)


def np_hash_uint32(k):
    """hash_bytes_uint32 (common/hashfn.c:627): Jenkins final() over a = seed + k"""
    with np.errstate(over="ignore"):
        a = b = c = np.full(k.shape, (0x9e3779b9 + 4 + 3923095) & 0xffffffff, dtype=np.uint32)
        a = a + k.astype(np.uint32)
        c = c ^ b; c = c - _rot(b, 14)
        a = a ^ c; a = a - _rot(c, 11)
        b = b ^ a; b = b - _rot(a, 25)
        c = c ^ b; c = c - _rot(b, 16)
        a = a ^ c; a = a - _rot(c, 4)
        b = b ^ a; b = b - _rot(a, 14)
        c = c ^ b; c = c - _rot(b, 24)
    return c


def np_hashint8(v):
    """hashint8 (access/hash/hashfunc.c:84-102): fold the high half into the low half, then hash_uint32"""
    lo = (v & 0xffffffff).astype(np.uint32)
    hi = ((v >> 32) & 0xffffffff).astype(np.uint32)
    return np_hash_uint32(np.where(v >= 0, lo ^ hi, lo ^ ~hi))


def np_jump(h, nseg):
    """jump_consistent_hash (cdb/cdbhash.c:530-541) on every element"""
    key = h.astype(np.uint64)
    b = np.full(h.shape, -1, dtype=np.int64)
    j = np.zeros(h.shape, dtype=np.int64)
    live = j < nseg
    while live.any():
        b = np.where(live, j, b)
        with np.errstate(over="ignore"):
            key = np.where(live, key * np.uint64(2862933555777941757) + np.uint64(1), key)
        nj = ((b + 1).astype(np.float64) * (float(1 << 31) / ((key >> np.uint64(33)).astype(np.float64) + 1.0))).astype(np.int64)
        j = np.where(live, nj, j)
        live = j < nseg
    return b


def _dec(v, ds):
    s = "-" if v < 0 else ""
    d = str(abs(v)).rjust(ds + 1, "0")
    return s + (d[:-ds] + "." + d[-ds:] if ds else d)


def test_sf1_q1_on_two_cpu_segments(oracle):
    sz = tpch.sizes(1)
    assert sz["lineitem"] == 6001215
    li = tpch._rel("lineitem", tpch.gen_lineitem(42, sz["lineitem"], sz["supplier"], sz["part"]))
    li.set_dict_hashes(oracle.hashbpchar)
    okey = li.columns[li.attno("l_orderkey") - 1]
    dest = np_jump(np_hashint8(okey.astype(np.int64)), 2)
    # the vectorised route is the oracle's route (sampled) - and, through tests/test_ref_exec.py, the reference's
    L = oracle.lib()
    t = (C.c_int32 * 1)(P.INT8)
    for i in np.random.default_rng(0).integers(0, len(okey), 2000):
        assert L.ora_cdbhash_segment(t, (C.c_int64 * 1)(int(okey[i])), None, 1, 2) == dest[i]
    shares = np.bincount(dest, minlength=2)
    assert shares.sum() == sz["lineitem"] and abs(int(shares[0]) - int(shares[1])) < 0.01 * sz["lineitem"]

    empty = [tpch._rel(n, c, d) for n, c, d in (
        ("orders", tpch.gen_orders(42, 0, 1), None), ("customer", tpch.gen_customer(42, 0), {"c_mktsegment": tpch.SEGMENTS}),
        ("supplier", tpch.gen_supplier(42, 0), None))]
    nation, region = tpch.gen_nation_region()
    rest = empty + [tpch._rel("nation", nation, {"n_name": tpch.NATIONS}), tpch._rel("region", region, {"r_name": tpch.REGIONS})]
    for r in rest:
        r.set_dict_hashes(oracle.hashbpchar)
    segs = [[li.take(np.nonzero(dest == s)[0])] + rest for s in range(2)]
    res = oracle.execute(tpch.q1_plan(2), segs, nthreads=2)
    got = tpch.format_q1(res.rows)
    assert set(res.segments) == {0}

    # independent evaluation: exact integer arithmetic on the scaled columns
    col = lambda n: li.columns[li.attno(n) - 1]
    keep = col("l_shipdate") <= tpch.Q1_CUTOFF
    rf, ls = col("l_returnflag")[keep], col("l_linestatus")[keep]
    qty, ext, disc, tax = (col(n)[keep].astype(np.int64) for n in ("l_quantity", "l_extendedprice", "l_discount", "l_tax"))
    so = os.path.join(HERE, "..", "oracle", "_ref", "libexec_ref.so")
    ref = C.CDLL(so) if os.path.exists(so) else None
    if ref is not None:
        ref.ref_numeric_binop.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(256)

    def avg_text(total, ds, n):
        if ref is None:
            return None
        assert ref.ref_numeric_binop(3, _dec(total, ds).encode(), str(n).encode(), buf, 256) == 0
        return buf.value.decode()

    want = []
    for key in sorted(set(zip(rf.tolist(), ls.tolist()))):
        m = (rf == key[0]) & (ls == key[1])
        n = int(m.sum())
        s_qty, s_ext, s_disc = int(qty[m].sum()), int(ext[m].sum()), int(disc[m].sum())
        dp = ext[m] * (100 - disc[m])
        s_dp = int(dp.sum())
        s_ch = sum(int(x) for x in np.add.reduceat(dp * (100 + tax[m]), np.arange(0, n, 1 << 16)))
        want.append([chr(key[0]), chr(key[1]), _dec(s_qty, 2), _dec(s_ext, 2), _dec(s_dp, 4), _dec(s_ch, 6),
                     avg_text(s_qty, 2, n), avg_text(s_ext, 2, n), avg_text(s_disc, 2, n), str(n)])
    assert len(got) == len(want) == 4
    for g, w in zip(got, want):
        for i, (a, b) in enumerate(zip(g, w)):
            if b is not None:
                assert a == b, (g[:2], i, a, b)


def test_sf1_q3_q5_against_pandas(oracle):
    """The oracle's join pipelines at SF1 (7.65 M rows scanned) against an independent evaluation of the same SQL with pandas
    merges and exact int64 sums: a size the reference has no fixture for."""
    import pandas as pd
    rels = tpch.gen_tables(1, oracle.hashbpchar)
    by = {r.name: r for r in rels}
    df = lambda name, cols: pd.DataFrame({c: by[name].columns[by[name].attno(c) - 1] for c in cols})
    li = df("lineitem", ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount", "l_shipdate"])
    od = df("orders", ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"])
    cu = df("customer", ["c_custkey", "c_nationkey", "c_mktsegment"])
    su = df("supplier", ["s_suppkey", "s_nationkey"])
    na = df("nation", ["n_nationkey", "n_regionkey", "n_name"])
    li["rev"] = li.l_extendedprice.astype(np.int64) * (100 - li.l_discount.astype(np.int64))       # scale 4

    # Q3 (rpt_tpch.source:458-480)
    seg = tpch.SEGMENTS.index("MACHINERY")
    cutoff = tpch.date_to_days(1995, 3, 15)
    j = od[od.o_orderdate < cutoff].merge(cu[cu.c_mktsegment == seg], left_on="o_custkey", right_on="c_custkey")
    j = li[li.l_shipdate > cutoff].merge(j, left_on="l_orderkey", right_on="o_orderkey")
    g = j.groupby(["l_orderkey", "o_orderdate", "o_shippriority"], as_index=False).rev.sum()
    g = g.sort_values(["rev", "o_orderdate"], ascending=[False, True], kind="stable").head(11)
    got = oracle.execute(tpch.q3_plan(seg, 1), [rels], nthreads=1).rows
    assert len(got) == 10
    want = [(int(r.l_orderkey), _dec(int(r.rev), 4), int(r.o_orderdate), int(r.o_shippriority)) for r in g.itertuples()]
    assert [(r[1], r[2]) for r in got] == [(w[1], w[2]) for w in want[:10]]
    if len({(w[1], w[2]) for w in want}) == len(want):          # no tie on the sort key at the cut: the rows are determined
        assert [tuple(r) for r in got] == want[:10]

    # Q5 (rpt_tpch.source:512-535)
    region = tpch.REGIONS.index("AMERICA")
    lo, hi = tpch.date_to_days(1997, 1, 1), tpch.date_to_days(1998, 1, 1)
    j = li.merge(od[(od.o_orderdate >= lo) & (od.o_orderdate < hi)], left_on="l_orderkey", right_on="o_orderkey")
    j = j.merge(cu, left_on="o_custkey", right_on="c_custkey")
    j = j.merge(su, left_on=["l_suppkey", "c_nationkey"], right_on=["s_suppkey", "s_nationkey"])
    j = j.merge(na[na.n_regionkey == region], left_on="s_nationkey", right_on="n_nationkey")
    want5 = {int(k): _dec(int(v), 4) for k, v in j.groupby("n_name").rev.sum().items()}
    got5 = {int(r[0]): r[1] for r in oracle.execute(tpch.q5_plan(region, 1), [rels], nthreads=1).rows}
    assert got5 == want5 and len(got5) == 5


def test_sf1_q3_q5_on_three_segments_with_motions(oracle):
    """the same database spread over 3 CPU segments as the DDL would (lineitem / orders by orderkey, customer by c_custkey,
    supplier by s_suppkey; nation / region replicated): the plans with Redistribute and Gather Motions return what one
    segment returns"""
    rels = tpch.gen_tables(1, oracle.hashbpchar)
    nsegs = 3
    segs = [[] for _ in range(nsegs)]
    for rel in rels:
        key = tpch.DIST_KEY[rel.name] if not (rel.name in ("customer", "supplier")) else {"customer": "c_custkey", "supplier": "s_suppkey"}[rel.name]
        if key is None:
            for s in range(nsegs):
                segs[s].append(rel)
            continue
        col = rel.columns[rel.attno(key) - 1]
        h = np_hashint8(col.astype(np.int64)) if col.dtype == np.int64 else np_hash_uint32(col.astype(np.int64).astype(np.uint32))
        dest = np_jump(h, nsegs)
        for s in range(nsegs):
            segs[s].append(rel.take(np.nonzero(dest == s)[0]))
    seg = tpch.SEGMENTS.index("MACHINERY")
    region = tpch.REGIONS.index("AMERICA")
    one3 = oracle.execute(tpch.q3_plan(seg, 1), [rels], nthreads=1).rows
    many3 = oracle.execute(tpch.q3_plan(seg, nsegs, customer_replicated=False), segs, nthreads=nsegs).rows
    assert [(r[1], r[2]) for r in many3] == [(r[1], r[2]) for r in one3] and len(many3) == 10
    one5 = oracle.execute(tpch.q5_plan(region, 1), [rels], nthreads=1).rows
    many5 = oracle.execute(tpch.q5_plan(region, nsegs, replicated=False), segs, nthreads=nsegs).rows
    assert sorted(map(tuple, many5)) == sorted(map(tuple, one5)) and len(many5) == 5
