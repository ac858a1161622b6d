"""2-process CPU (gloo) run of the N>1 host path: two-stage Q1 across ranks.

Each rank: oracle partial aggregation over its cdbhash shard of the reference fixture; the partial
states are redistributed by cdbhash(group key) with harness.exchange_by_destination (the protocol
csrc/motion.cu implements over NCCL); the owner combines them and finalises with the PRODUCT's
host-side numeric code (libcbexec cb_numeric_*); rank 0 gathers and compares with the reference's
expected rows."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from cloudberry_b200 import capi, harness, tpch
    from cloudberry_b200 import plan as P
    from oracle import oracle as O
    from gpu_util import shard
    rank, _, world = harness.env_rank()
    dist = harness.init_dist("gloo")
    token = harness.broadcast_token(dist, lambda: os.urandom(128), rank)
    assert len(token) == 128
    rels, exp = tpch.load_golden(O.hashbpchar)
    mine = shard(O, rels, world)[rank]
    # partial stage on this segment: the plan below the Redistribute Motion
    full = tpch.q1_plan(world)                         # Gather <- Final <- Redistribute <- Partial <- Scan
    partial = full.plan.lefttree.contents.lefttree.contents.lefttree      # CbPlan* of the partial Agg
    import ctypes
    res = O.lib().ora_execute(partial, *(_segs(O, [mine])), 1)
    assert res, O.lib().ora_last_error()
    R = O.Result(res)
    keys = np.array([[r[0], r[1]] for r in R.rows], dtype=np.int64).reshape(-1, 2)
    L = O.lib()
    t = (C.c_int32 * 2)(P.BPCHAR1, P.BPCHAR1)
    dest = np.array([L.ora_cdbhash_segment(t, (C.c_int64 * 2)(int(k[0]), int(k[1])), None, 2, world) for k in keys], dtype=np.int64)
    naggs = len(R.rows[0]) - 2 if R.rows else 8
    cols = [keys[:, 0].copy(), keys[:, 1].copy()]
    for a in range(naggs):
        n = np.array([st[2 + a][0] for st in R.states], dtype=np.int64)
        s = np.array([st[2 + a][1] for st in R.states], dtype=object)
        lo = np.array([int(v) & (2 ** 64 - 1) for v in s], dtype=np.uint64).view(np.int64)
        hi = np.array([(int(v) >> 64) for v in s], dtype=np.int64)
        cols += [n, lo, hi]
    got, matrix = harness.exchange_by_destination(dist, rank, world, dest, cols)
    assert matrix.sum() == 4 * world or matrix.sum() <= 4 * world
    # final stage: combine states per group (int8_avg_combine / numeric_avg_combine), finalise on the host
    E = capi.ex()
    groups = {}
    for i in range(len(got[0])):
        k = (int(got[0][i]), int(got[1][i]))
        g = groups.setdefault(k, [[0, 0] for _ in range(naggs)])
        for a in range(naggs):
            n = int(got[2 + 3 * a][i])
            v = (int(got[4 + 3 * a][i]) << 64) | (int(got[3 + 3 * a][i]) & (2 ** 64 - 1))
            g[a][0] += n
            g[a][1] += v
    buf = C.create_string_buffer(128)
    rows = []
    dscales = [2, 2, 4, 6, 2, 2, 2, 0]
    kinds = ["sum", "sum", "sum", "sum", "avg", "avg", "avg", "count"]
    for k, g in groups.items():
        row = [k[0], k[1]]
        for a in range(naggs):
            n, v = g[a]
            lo = v & (2 ** 64 - 1)
            lo = lo - 2 ** 64 if lo >= 2 ** 63 else lo
            if kinds[a] == "count":
                row.append(n)
            elif kinds[a] == "sum":
                E.cb_numeric_sum_text(lo, v >> 64, dscales[a], buf, 128)
                row.append(buf.value.decode())
            else:
                E.cb_numeric_avg_text(lo, v >> 64, dscales[a], n, buf, 128)
                row.append(buf.value.decode())
        rows.append(row)
    allrows = [None] * world
    dist.all_gather_object(allrows, rows)              # Gather Motion to rank 0
    ok = True
    if rank == 0:
        flat = [r for part in allrows for r in part]
        ok = tpch.format_q1(flat) == exp["q1"]
        print("GLOO", "PASS" if ok else "FAIL", world)
    t_ms = harness.max_over_ranks(dist, [float(rank)])
    assert t_ms[0] == world - 1
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def _segs(O, segments):
    import ctypes as C
    keep = []
    segs = (O.OraSegment * len(segments))()
    for s, rels in enumerate(segments):
        arr = (C.POINTER(O.OraRel) * max(len(rels), 1))()
        for i, rel in enumerate(rels):
            r, k = O.make_rel(rel)
            keep += [r, k]
            arr[i] = C.pointer(r)
        segs[s].nrels = len(rels)
        segs[s].rels = arr
        keep.append(arr)
    _segs.keep = keep
    return segs, len(segments)


if __name__ == "__main__":
    main()
