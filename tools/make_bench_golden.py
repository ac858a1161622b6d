#!/usr/bin/env python3
"""tools/make_bench_golden.py - exact answers of bench.py's FULL-SIZE configurations, computed with numpy alone.

bench.py runs TPC-H Q1 / Q3 / Q5 and SSB Q4.x at BASELINE.json's sizes (SF100, SF300), where the CPU oracle cannot
follow (it takes seconds at SF1).  The synthetic tables are counter based (cloudberry_b200/tpch.py, ssb.py: every value
is a pure function of (seed, column, row index), and the device generator csrc/gen.cu computes the same formulas), so
the answers can be computed on the host without executor, hashing or joins: a lineitem row's order, customer, supplier
and nation follow from its row index by arithmetic.  This script evaluates the SQL that way, chunk by chunk on all host
cores, in exact integer arithmetic, and writes tests/golden/bench_golden.json.  bench.py compares every run's result
rows with it (`result_check`) at every N; tests/test_bench_golden.py re-derives a small slice.

Nothing here imports the oracle or the product libraries: it is the third, independent evaluation.

    python tools/make_bench_golden.py [--only q1,q3,q5,q5_300,ssb] [--procs 8]      # ~25 min of 8 cores for everything
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloudberry_b200 import ssb, tpch  # noqa: E402  (numpy generators and constants only)

CHUNK = 28 * 150000          # lineitem rows per task: whole blocks of 7 orders, so no order straddles two chunks
SEED = 42
OUT = os.path.join(ROOT, "tests", "golden", "bench_golden.json")


def _isum(a):
    """exact sum of an int64 array whose total may leave 64 bits: partial sums of 1 M elements, added as Python ints"""
    t = 0
    for i in range(0, len(a), 1 << 20):
        t += int(a[i:i + (1 << 20)].sum(dtype=np.int64))
    return t


# ---------------------------------------------------------------------------------------------------------------
# Q1: per (returnflag, linestatus): count, sum(qty), sum(ext), sum(ext * (1 - disc)), sum(ext * (1 - disc) * (1 + tax)),
# sum(disc); scaled integers (scale 2, 2, 4, 6, 2).  Rows [lo, hi) of the generator's row space.
# ---------------------------------------------------------------------------------------------------------------
def q1_chunk(args):
    lo, hi, sz = args
    c = tpch.gen_lineitem(SEED, sz["lineitem"], sz["supplier"], sz["part"], lo=lo, hi=hi)
    m = c["l_shipdate"] <= tpch.Q1_CUTOFF
    ext, disc, tax, qty = c["l_extendedprice"], c["l_discount"], c["l_tax"], c["l_quantity"]
    dp = ext * (100 - disc)
    ch = dp * (100 + tax)
    out = {}
    code = c["l_returnflag"].astype(np.int64) * 256 + c["l_linestatus"]
    for g in np.unique(code[m]).tolist():
        s = m & (code == g)
        out[g] = [int(s.sum()), _isum(qty[s]), _isum(ext[s]), _isum(dp[s]), _isum(ch[s]), _isum(disc[s])]
    return out


def run_q1(pool, sf, nranks):
    """shard r = generator rows [r * n, (r + 1) * n): what bench.py's rank r holds (weak scaling)"""
    sz = tpch.sizes(sf)
    n = sz["lineitem"]
    shards = []
    for r in range(nranks):
        tasks = [(lo, min(lo + CHUNK, (r + 1) * n), sz) for lo in range(r * n, (r + 1) * n, CHUNK)]
        acc = {}
        for part in pool.imap_unordered(q1_chunk, tasks, chunksize=4):
            for g, v in part.items():
                a = acc.setdefault(g, [0] * 6)
                for i in range(6):
                    a[i] += v[i]
        shards.append({"%c%c" % (g // 256, g % 256): [str(x) for x in v] for g, v in sorted(acc.items())})
        print("q1 shard %d done" % r, flush=True)
    return {"sf": sf, "rows_per_shard": n, "cutoff_days": tpch.Q1_CUTOFF,
            "state": ["count", "sum_qty(scale 2)", "sum_base_price(2)", "sum_disc_price(4)", "sum_charge(6)", "sum_disc(2)"],
            "shards": shards}


# ---------------------------------------------------------------------------------------------------------------
# Q3: revenue per order for orders of the segment before the cutoff, lines shipped after it; top 10 by revenue
# desc, o_orderdate.  A lineitem row's order index is arithmetic; rows of one order are adjacent.
# ---------------------------------------------------------------------------------------------------------------
def _line_order_index(j):
    return (j // 28) * 7 + tpch._LINE_OFF[j % 28]


def q3_chunk(args):
    lo, hi, sz, seg, cutoff, keep = args
    j = np.arange(lo, hi, dtype=np.int64)
    oidx = _line_order_index(j)
    odate = tpch.STARTDATE + (tpch._u(SEED, 12, oidx) % np.uint64(tpch.ORDER_DATE_SPAN)).astype(np.int64)
    ship = odate + 1 + (tpch._u(SEED, 25, j) % np.uint64(121)).astype(np.int64)
    cust = (tpch._u(SEED, 11, oidx) % np.uint64(sz["customer"])).astype(np.int64)         # c_custkey - 1
    cseg = (tpch._u(SEED, 32, cust) % np.uint64(5)).astype(np.int64)
    # the last few lineitem rows name orders beyond the orders table (sizes() rounds the two independently): no partner
    m = (odate < cutoff) & (ship > cutoff) & (cseg == seg) & (oidx < sz["orders"])
    if not m.any():
        return [], 0
    qty = 1 + (tpch._u(SEED, 21, j) % np.uint64(50)).astype(np.int64)
    pk = 1 + (tpch._u(SEED, 22, j) % np.uint64(sz["part"])).astype(np.int64)
    price = 90000 + (pk // 10) % 20001 + 100 * (pk % 1000)
    disc = (tpch._u(SEED, 24, j) % np.uint64(11)).astype(np.int64)
    rev = (qty * price * (100 - disc))[m]
    o = oidx[m]
    starts = np.flatnonzero(np.r_[True, o[1:] != o[:-1]])
    tot = np.add.reduceat(rev, starts)
    oo = o[starts]
    od = odate[m][starts]
    k = min(keep, len(tot))
    top = np.lexsort((od, -tot))[:k]
    return [(int(tot[i]), int(od[i]), int(tpch.order_key(oo[i]))) for i in top], int(len(tot))


def run_q3(pool, sf, limit=10):
    sz = tpch.sizes(sf)
    n = sz["lineitem"]
    seg = tpch.SEGMENTS.index("MACHINERY")
    cutoff = tpch.date_to_days(1995, 3, 15)
    keep = limit + 6
    tasks = [(lo, min(lo + CHUNK, n), sz, seg, cutoff, keep) for lo in range(0, n, CHUNK)]
    cand, groups = [], 0
    for res in pool.imap_unordered(q3_chunk, tasks, chunksize=4):
        cand += res[0]
        groups += res[1]
    cand.sort(key=lambda t: (-t[0], t[1], t[2]))
    top = cand[:keep]
    # the SQL orders by (revenue desc, o_orderdate): the answer is only defined if the cut does not fall inside a tie
    assert (top[limit - 1][0], top[limit - 1][1]) != (top[limit][0], top[limit][1]), "tie at the LIMIT boundary"
    for a, b in zip(top[:limit], top[1:limit]):
        assert (a[0], a[1]) != (b[0], b[1]), "tie inside the top rows: row order undefined"
    return {"sf": sf, "segment": "MACHINERY", "cutoff_days": cutoff, "groups": groups,
            "columns": ["l_orderkey", "revenue(scale 4)", "o_orderdate(days since 2000-01-01)", "o_shippriority"],
            "rows": [[str(k), str(rev), d, 0] for rev, d, k in top[:limit]]}


# ---------------------------------------------------------------------------------------------------------------
# Q5: revenue per nation of the region: lineitem -> order (date range) -> customer nation == supplier nation
# ---------------------------------------------------------------------------------------------------------------
def q5_chunk(args):
    lo, hi, sz, region, d_lo, d_hi = args
    j = np.arange(lo, hi, dtype=np.int64)
    oidx = _line_order_index(j)
    odate = tpch.STARTDATE + (tpch._u(SEED, 12, oidx) % np.uint64(tpch.ORDER_DATE_SPAN)).astype(np.int64)
    m = (odate >= d_lo) & (odate < d_hi) & (oidx < sz["orders"])
    j, oidx = j[m], oidx[m]
    cust = (tpch._u(SEED, 11, oidx) % np.uint64(sz["customer"])).astype(np.int64)
    cnat = (tpch._u(SEED, 31, cust) % np.uint64(25)).astype(np.int64)
    supp = (tpch._u(SEED, 23, j) % np.uint64(sz["supplier"])).astype(np.int64)
    snat = (tpch._u(SEED, 41, supp) % np.uint64(25)).astype(np.int64)
    m2 = (cnat == snat) & (np.array(tpch.NATION_REGION, dtype=np.int64)[snat] == region)
    j, snat = j[m2], snat[m2]
    qty = 1 + (tpch._u(SEED, 21, j) % np.uint64(50)).astype(np.int64)
    pk = 1 + (tpch._u(SEED, 22, j) % np.uint64(sz["part"])).astype(np.int64)
    price = 90000 + (pk // 10) % 20001 + 100 * (pk % 1000)
    disc = (tpch._u(SEED, 24, j) % np.uint64(11)).astype(np.int64)
    rev = qty * price * (100 - disc)
    return {int(g): _isum(rev[snat == g]) for g in np.unique(snat).tolist()}


def run_q5(pool, sf):
    sz = tpch.sizes(sf)
    n = sz["lineitem"]
    region = tpch.REGIONS.index("AMERICA")
    d_lo, d_hi = tpch.date_to_days(1997, 1, 1), tpch.date_to_days(1998, 1, 1)
    tasks = [(lo, min(lo + CHUNK, n), sz, region, d_lo, d_hi) for lo in range(0, n, CHUNK)]
    acc = {}
    for part in pool.imap_unordered(q5_chunk, tasks, chunksize=4):
        for g, v in part.items():
            acc[g] = acc.get(g, 0) + v
    return {"sf": sf, "region": "AMERICA", "columns": ["n_name", "revenue(scale 4)"],
            "rows": [[tpch.NATIONS[g], str(v)] for g, v in sorted(acc.items(), key=lambda kv: -kv[1])]}


# ---------------------------------------------------------------------------------------------------------------
# SSB Q4.1 - Q4.3: dimension attributes by arithmetic from the fact row's keys
# ---------------------------------------------------------------------------------------------------------------
_SSB = {}


def _ssb_dims(sf):
    if sf not in _SSB:
        sz = ssb.sizes(sf)
        dk, dy = ssb._dates()
        _SSB[sf] = (sz, dy.astype(np.int64), np.array(ssb.NATION_REGION, dtype=np.int64))
    return _SSB[sf]


def ssb_chunk(args):
    lo, hi, sf, seed = args
    sz, dyear, nat_reg = _ssb_dims(sf)
    i = np.arange(lo, hi, dtype=np.int64)
    c = (tpch._u(seed, 51, i) % np.uint64(sz["customer"])).astype(np.int64)
    p = (tpch._u(seed, 52, i) % np.uint64(sz["part"])).astype(np.int64)
    s = (tpch._u(seed, 53, i) % np.uint64(sz["supplier"])).astype(np.int64)
    year = dyear[(tpch._u(seed, 54, i) % np.uint64(2556)).astype(np.int64)]
    profit = (100 + (tpch._u(seed, 55, i) % np.uint64(10000000)).astype(np.int64)) - (50 + (tpch._u(seed, 56, i) % np.uint64(120000)).astype(np.int64))
    c_city = (tpch._u(seed, 21, c) % np.uint64(250)).astype(np.int64)
    s_city = (tpch._u(seed, 31, s) % np.uint64(250)).astype(np.int64)
    brand = (tpch._u(seed, 41, p) % np.uint64(1000)).astype(np.int64)
    c_nat, s_nat = c_city // 10, s_city // 10
    am = ssb.REGIONS.index("AMERICA")
    y78 = (year == 1997) | (year == 1998)
    out = {}
    sel = {
        "q4.1": ((nat_reg[c_nat] == am) & (nat_reg[s_nat] == am) & (brand // 200 <= 1), (year, c_nat)),
        "q4.2": ((nat_reg[c_nat] == am) & (nat_reg[s_nat] == am) & (brand // 200 <= 1) & y78, (year, s_nat, brand // 40)),
        "q4.3": ((s_nat == ssb.NATIONS.index("UNITED STATES")) & y78 & (brand // 40 == ssb.CATEGORIES.index("MFGR#14")), (year, s_city, brand)),
    }
    for q, (m, keys) in sel.items():
        code = np.zeros(int(m.sum()), dtype=np.int64)
        for k in keys:
            code = code * 10000 + k[m]
        u, inv = np.unique(code, return_inverse=True)
        tot = np.zeros(len(u), dtype=np.int64)
        np.add.at(tot, inv, profit[m])
        out[q] = dict(zip(u.tolist(), tot.tolist()))
    return out


def run_ssb(pool, sf, seed=7):
    n = ssb.sizes(sf)["lineorder"]
    step = 4000000
    tasks = [(lo, min(lo + step, n), sf, seed) for lo in range(0, n, step)]
    acc = {"q4.1": {}, "q4.2": {}, "q4.3": {}}
    for part in pool.imap_unordered(ssb_chunk, tasks, chunksize=2):
        for q, d in part.items():
            a = acc[q]
            for k, v in d.items():
                a[k] = a.get(k, 0) + v
    nk = {"q4.1": 2, "q4.2": 3, "q4.3": 3}
    res = {"sf": sf, "seed": seed, "rows_lineorder": n}
    for q, d in acc.items():
        rows = []
        for code, v in sorted(d.items()):
            ks = []
            for _ in range(nk[q]):
                ks.append(code % 10000)
                code //= 10000
            rows.append(ks[::-1] + [str(v)])
        res[q] = rows
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="q1,q3,q5,q5_300,ssb")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--q1-ranks", type=int, default=8)
    ap.add_argument("--out", default=OUT)
    args = ap.parse_args()
    want = set(args.only.split(","))
    gold = json.load(open(args.out)) if os.path.exists(args.out) else {}
    gold["how"] = ("tools/make_bench_golden.py: numpy evaluation of the SQL over the counter-based synthetic tables (seed 42; SSB seed 7), "
                   "exact integers as strings; independent of oracle/ and of the CUDA path")
    t0 = time.time()
    with mp.Pool(args.procs) as pool:
        if "q5" in want:
            gold["q5_sf100"] = run_q5(pool, 100)
            print("q5 sf100 %.0f s" % (time.time() - t0), flush=True)
        if "q3" in want:
            gold["q3_sf100"] = run_q3(pool, 100)
            print("q3 sf100 %.0f s" % (time.time() - t0), flush=True)
        if "ssb" in want:
            gold["ssb_sf100"] = run_ssb(pool, 100)
            print("ssb sf100 %.0f s" % (time.time() - t0), flush=True)
        if "q5_300" in want:
            gold["q5_sf300"] = run_q5(pool, 300)
            print("q5 sf300 %.0f s" % (time.time() - t0), flush=True)
        if "q1" in want:
            gold["q1_sf100"] = run_q1(pool, 100, args.q1_ranks)
            print("q1 sf100 %.0f s" % (time.time() - t0), flush=True)
        if "sf10" in want:
            # tests/test_gpu_sf.py: the same queries at a size between the oracle's reach and the bench's
            gold["q3_sf10"] = run_q3(pool, 10)
            gold["q5_sf10"] = run_q5(pool, 10)
            gold["q1_sf10"] = run_q1(pool, 10, 1)
            gold["ssb_sf10"] = run_ssb(pool, 10)
            print("sf10 %.0f s" % (time.time() - t0), flush=True)
        for k in sorted(want):
            if k.startswith("small"):
                # tests/test_bench_golden.py: the same code at a size the oracle can follow
                sf = float(k[5:] or 0.05)
                gold["small"] = {"sf": sf, "q1": run_q1(pool, sf, 2), "q3": run_q3(pool, sf), "q5": run_q5(pool, sf), "ssb": run_ssb(pool, sf)}
    json.dump(gold, open(args.out, "w"), indent=1, sort_keys=True)
    print("wrote %s in %.0f s" % (args.out, time.time() - t0))


if __name__ == "__main__":
    main()
