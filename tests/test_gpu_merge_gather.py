"""Sorted Gather Motion (Motion.sendSorted, execMotionSortedReceiver nodeMotion.c:433): every sender's stream
is ordered by the Motion's sort keys and the receiver returns them merged.  The device merge (csrc/agg.cu: every row finds its
place by binary searches in the other senders' runs) against a stable Python merge and against the oracle's merge receive."""
import ctypes as C
import heapq

import numpy as np
import pytest

from cloudberry_b200 import capi, tpch
from cloudberry_b200 import plan as P
from cloudberry_b200.relation import HostRelation
from gpu_util import shard, to_device
from test_gpu_edge import agg_over, fact, scan

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _merged_order(ctx, dev, keycols, desc, max_runs):
    L = ctx.L
    n = len(keycols)
    order, nruns = C.c_void_p(), C.c_int32()
    ctx.check(L.cbgpu_merge_sorted_runs(ctx.h, dev.h, (C.c_int32 * n)(*keycols), (C.c_int32 * n)(*desc), None, n, max_runs,
                                        C.byref(order), C.byref(nruns)))
    if not order.value:
        return None, nruns.value
    host = (C.c_uint32 * dev.nrows)()
    ctx.check(L.cbgpu_read_u32(ctx.h, order, dev.nrows, host))
    L.cbgpu_dev_free(ctx.h, order)
    return list(host), nruns.value


@pytest.mark.parametrize("nruns,per_run", [(1, 100), (2, 1), (3, 1000), (8, 4097), (64, 50)])
@pytest.mark.parametrize("desc", [(0, 0), (1, 0), (0, 1)])
def test_device_merge_against_a_stable_merge(ctx, nruns, per_run, desc):
    """runs of ragged lengths (some empty), two keys with many ties, NULLs in the first key: the merged order is the one a
    stable merge of the runs gives - NULLs last ascending / first descending, equal keys in sender order"""
    rng = np.random.default_rng(nruns * 1000 + per_run + desc[0] * 7 + desc[1] * 13)
    runs = []
    for r in range(nruns):
        n = 0 if (r == 1 and nruns > 2) else int(rng.integers(max(per_run // 2, 1), per_run + 1))
        a = rng.integers(-5, 6, n)
        an = (rng.random(n) < 0.2).astype(np.uint8)
        b = rng.integers(0, 4, n)

        def k(i):
            x = (1, 0) if an[i] else (0, int(a[i]))              # NULL sorts above every value ...
            x = tuple(-v for v in x) if desc[0] else x           # ... and first when the key is descending
            return (x, -int(b[i]) if desc[1] else int(b[i]))
        idx = sorted(range(n), key=k)
        runs.append([(a[i], an[i], b[i], k(i)) for i in idx])
    rows = [row for run in runs for row in run]
    if not rows:
        pytest.skip("empty")
    rel = HostRelation("runs", ["a", "b", "tag"], [P.INT4, P.INT8, P.INT8],
                       [np.array([r[0] for r in rows]), np.array([r[2] for r in rows]), np.arange(len(rows))],
                       nulls=[np.array([r[1] for r in rows], dtype=np.uint8), None, None])
    dev = capi.DeviceRelation.from_host(ctx, rel)
    got, found = _merged_order(ctx, dev, [0, 1], list(desc), 64)
    # the stable merge: by key, then by position in the concatenation (= sender order, then order within the sender)
    want = [i for i, _ in sorted(enumerate(rows), key=lambda t: (t[1][3], t[0]))]
    assert found <= max(sum(1 for r in runs if r), 1)
    if got is None:
        assert want == list(range(len(rows)))
    else:
        assert got == want
    dev.free()


def test_an_unsorted_sender_is_reported(ctx):
    rel = HostRelation("runs", ["a"], [P.INT8], [np.array([1, 5, 2, 9, 3, 7, 0, 4])])
    dev = capi.DeviceRelation.from_host(ctx, rel)
    with pytest.raises(capi.CbgpuError) as e:
        _merged_order(ctx, dev, [0], [0], 2)
    assert "sort order" in str(e.value)
    dev.free()


@pytest.mark.parametrize("nsegs", [2, 3, 5])
def test_merge_gather_through_the_executor(ctx, oracle, nsegs):
    """per segment: aggregate, keep the local top 40 by (count desc, k) - then a Gather Motion that merges the segments'
    streams is the plan's top node, so its order IS the result's order; the oracle's merge receive gives the same rows in the
    same order (equal keys: lower segment first on both sides)"""
    from oracle import oracle as O
    n = 30011
    fo = fact(n, seed=51, kmax=300).set_dict_hashes(O.hashbpchar)
    fp = fact(n, seed=51, kmax=300).set_dict_hashes(capi.hashbpchar)
    cut = np.array_split(np.arange(n), nsegs)
    segs_o = [[fo.take(c)] for c in cut]
    segs_p = [[fp.take(c)] for c in cut]
    sc = scan(1, fo, ["k", "amt"])
    agg = agg_over(sc, ["k", "amt"], ["k"], [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None)])
    keys = [(3, True), (1, False)]
    top = P.LimitSort(agg, keys, 40)
    plan = P.Motion(top, P.MOTIONTYPE_GATHER, sort_keys=keys)
    want = oracle.execute(plan, segs_o).rows
    dsegs = [to_device(ctx, s) for s in segs_p]
    cl = capi.Cluster(ctx, dsegs)
    got = cl.run(plan)
    assert len(want) == 40 * nsegs
    assert [tuple(map(str, r)) for r in got.rows] == [tuple(map(str, r)) for r in want]
    # ... and the order is the Motion's: count descending, then k
    ks = [(-int(r[2]), int(r[0])) for r in got.rows]
    assert ks == sorted(ks)
    cl.close()
    for d in dsegs:
        for r in d:
            r.free()


def test_q3_with_the_reference_plan_shape(ctx, oracle, golden):
    """Limit <- Gather Motion (merge) <- Limit <- Sort, the plan the reference makes for TPC-H Q3 on several segments"""
    rels, exp = golden
    for r in rels:
        r.set_dict_hashes(capi.hashbpchar)
    segs = shard(oracle, rels, 3)
    dsegs = [to_device(ctx, s) for s in segs]
    cl = capi.Cluster(ctx, dsegs)
    seg = exp["dict"]["c_mktsegment_dict"].index("MACHINERY")
    res = cl.run(tpch.q3_plan(seg, 3, merge_gather=True))
    assert tpch.format_q3(res.rows) == exp["q3"]
    cl.close()
