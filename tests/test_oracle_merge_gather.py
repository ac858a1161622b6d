"""The oracle's merge receive (oracle/oracle.c materialize_motions: execMotionSortedReceiver, nodeMotion.c:433) against a stable
merge of the per-segment results done in Python - the checker checked, before tests/test_gpu_merge_gather.py relies on it."""
import numpy as np
import pytest

from cloudberry_b200 import plan as P
from test_gpu_edge import agg_over, fact, scan


@pytest.mark.parametrize("nsegs", [1, 2, 4])
def test_merge_receive_is_a_stable_merge_of_the_senders(nsegs):
    from oracle import oracle as O
    n = 20011
    fo = fact(n, seed=61, kmax=200).set_dict_hashes(O.hashbpchar)
    cut = np.array_split(np.arange(n), nsegs)
    segs = [[fo.take(c)] for c in cut]
    sc = scan(1, fo, ["k", "amt"])
    agg = agg_over(sc, ["k", "amt"], ["k"], [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None)])
    keys = [(3, True), (1, False)]
    top = P.LimitSort(agg, keys, 30)
    per_seg = [O.execute(top, [s]).rows for s in segs]
    assert all(len(r) == 30 for r in per_seg)
    tagged = [(-int(r[2]), int(r[0]), s, i, r) for s, rows in enumerate(per_seg) for i, r in enumerate(rows)]
    want = [t[4] for t in sorted(tagged, key=lambda t: t[:4])]
    got = O.execute(P.Motion(top, P.MOTIONTYPE_GATHER, sort_keys=keys), segs).rows
    assert [tuple(map(str, r)) for r in got] == [tuple(map(str, r)) for r in want]
    # without sort keys the streams arrive sender after sender
    plain = O.execute(P.Motion(top, P.MOTIONTYPE_GATHER), segs).rows
    assert [tuple(map(str, r)) for r in plain] == [tuple(map(str, r)) for rows in per_seg for r in rows]
